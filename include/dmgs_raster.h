/*
 * dmgs_raster.h -- C ABI of libdmgs_raster.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for DMGS's differentiable Gaussian-splatting rasteriser and the
 * mesh-face -> Gaussian binding that feeds it.  The reference binds this path through the
 * pybind11/torch module `diff_gaussian_rasterization._C` (imported at
 * /root/reference/gaussian_renderer/__init__.py:14; settings tuple at :36-49 / :129-142; call at
 * :86-94 / :178-186).  Here the same work is exposed as plain `extern "C"` functions taking raw
 * device pointers, sizes and a CUDA stream -- no torch types -- and dmgs_b200/rasterizer.py binds
 * them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library never allocates or frees device memory: state lives in three caller-owned
 *     byte buffers (geom / binning / image) whose sizes the *_bytes functions return -- the
 *     counterpart of the upstream module growing torch byte tensors through callbacks;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises
 *     unless params.debug != 0, in which case every stage is followed by a stream sync + check;
 *   - return value: 0 ok, < 0 argument error, > 0 a cudaError_t; dmgs_last_error() explains.
 *   - optional inputs are NULL when absent (exactly one of shs|colors_precomp and exactly one of
 *     (scales,rotations)|cov3D_precomp must be given, as the reference op demands).
 */
#ifndef DMGS_RASTER_H
#define DMGS_RASTER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors GaussianRasterizationSettings (gaussian_renderer/__init__.py:36-49) plus the two
 * switches needed to fold DMGS's Python SH path (gaussian_renderer/__init__.py:74-78, :166-170)
 * into preprocess. */
typedef struct dmgs_params {
    int32_t P;              /* number of Gaussians */
    int32_t sh_degree;      /* active SH degree (settings.sh_degree) */
    int32_t sh_coeffs;      /* stored coefficients per channel M (shs.shape[1] or features.shape[2]) */
    int32_t image_width;
    int32_t image_height;
    int32_t sh_layout;      /* 0: shs [P,M,3] (rasteriser layout); 1: features [P,3,M] (DMGS gs_info layout) */
    int32_t sh_activation;  /* 0: max(sh+0.5, 0) (rasteriser); 1: sigmoid(sh) (DMGS convert_SHs_python path) */
    int32_t debug;          /* settings.debug: sync + check after every stage */
    float tanfovx, tanfovy;
    float scale_modifier;
    float bg[3];
    float viewmatrix[16];   /* world_view_transform, flat row-major torch layout (= column-major W2C) */
    float projmatrix[16];   /* full_proj_transform, same layout */
    float campos[3];
} dmgs_params;

/* ---- buffer sizes (host-only, no CUDA calls) ------------------------------------------- */
size_t dmgs_geom_bytes(int32_t P);
size_t dmgs_binning_bytes(int32_t P, int64_t num_rendered, int32_t W, int32_t H);
size_t dmgs_image_bytes(int32_t W, int32_t H);

/* ---- forward, stage 1: replaces preprocessCUDA + InclusiveSum ----------
 * Per-Gaussian EWA projection / SH colour / tile rectangle and a stable depth sort of the
 * Gaussians (adaptive: only the digits in which the live depth keys differ are sorted).
 * Writes radii[P] (int32) and the number of (Gaussian, tile) instances to *num_rendered
 * (device uint32). */
int dmgs_preprocess_forward(const dmgs_params *prm, const float *means3D, const float *scales,
                            const float *rotations, const float *cov3D_precomp, const float *opacities,
                            const float *shs, const float *colors_precomp, int32_t *radii, void *geom,
                            uint32_t *num_rendered, void *stream);

/* ---- forward, stage 2: replaces duplicateWithKeys + SortPairs + identifyTileRanges (K3-K5) --
 * num_rendered is the value stage 1 produced (read back by the caller, as the upstream host
 * code does).  Produces the tile-major, depth-ordered instance list and per-tile ranges: by
 * direct placement from the depth-sorted Gaussian order for images of up to 16384 tiles, by a
 * stable radix partition of the emitted instances above that (or with DMGS_TILE_PARTITION=radix
 * in the environment). */
int dmgs_bin_forward(const dmgs_params *prm, const void *geom, int64_t num_rendered, void *binning, void *stream);
/* Same without the host read-back (no counterpart upstream): the binning buffer is sized for
 * `capacity` instances (dmgs_binning_bytes(P, capacity, W, H)) and `capacity` takes the place of
 * num_rendered in every later call on this state.  The real count stays on the device; if it
 * exceeds the capacity nothing is placed, every tile range is (0,0) (the frame renders as
 * background and its gradients are zero) and *overflow (device uint32) receives the count needed,
 * 0 otherwise.  The caller checks the flag at its next synchronisation point and repeats the frame
 * with a larger buffer.  Direct-placement path only (-9 for images above 16384 tiles). */
int dmgs_bin_forward_async(const dmgs_params *prm, const void *geom, int64_t capacity, void *binning,
                           uint32_t *overflow, void *stream);

/* ---- forward, stage 3: replaces renderCUDA forward (K6). out_color is [3,H,W]. */
int dmgs_blend_forward(const dmgs_params *prm, const void *geom, const void *binning, int64_t num_rendered,
                       float *out_color, void *image, void *stream);

/* ---- backward: replaces renderCUDA backward + computeCov2DCUDA + preprocessCUDA backward (K7-K9).
 * dL_dpix is [3,H,W].  Gradient outputs may be NULL when the corresponding input was not given;
 * dL_dmeans2D is [P,3] (x,y filled, z = 0), the contract scene/gaussian_model.py:405-407 reads. */
int dmgs_backward(const dmgs_params *prm, const float *means3D, const float *scales, const float *rotations,
                  const float *cov3D_precomp, const float *shs, const int32_t *radii, const void *geom,
                  const void *binning, const void *image, int64_t num_rendered, const float *dL_dpix,
                  float *dL_dmeans3D, float *dL_dmeans2D, float *dL_dopacity, float *dL_dcolors_precomp,
                  float *dL_dshs, float *dL_dscales, float *dL_drotations, float *dL_dcov3D, void *scratch,
                  void *stream);
size_t dmgs_backward_scratch_bytes(int32_t P);
/* The two halves of dmgs_backward, callable separately (bench.py brackets them with CUDA events):
 * K7 accumulates per-Gaussian blend gradients into `scratch` (P x 12 floats, followed by the square
 * counter of its persistent grid: always size it with dmgs_backward_scratch_bytes); K8+K9 consume it.  With
 * accumulate != 0 the gradients are ADDED to the output buffers (view-batched training:
 * one flat gradient buffer summed over the views of a step, no extra accumulation pass). */
int dmgs_blend_backward(const dmgs_params *prm, const void *geom, const void *binning, const void *image,
                        int64_t num_rendered, const float *dL_dpix, void *scratch, void *stream);
int dmgs_preprocess_backward(const dmgs_params *prm, const float *means3D, const float *scales,
                             const float *rotations, const float *cov3D_precomp, const float *shs,
                             const int32_t *radii, const void *geom, const void *scratch, float *dL_dmeans3D,
                             float *dL_dmeans2D, float *dL_dopacity, float *dL_dcolors_precomp, float *dL_dshs,
                             float *dL_dscales, float *dL_drotations, float *dL_dcov3D, int32_t accumulate,
                             void *stream);

/* ---- deferred SH gradient for view-batched training (dmgs_preprocess_backward with accumulate = 2).
 * A view's contribution to dL/dsh is rank one per channel, basis(dir) (x) g with g = dL/dcolour after the
 * activation, and its view-direction term of dL/dmean is linear in g as well.  In this mode
 * dmgs_preprocess_backward does NOT touch the SH rows (neither the coefficients nor their gradients): its
 * dL_dshs argument is this view's record array float[P][4] = {g.r, g.g, g.b, seen} (every other gradient is
 * accumulated as with accumulate = 1, the view-direction term of dL_dmeans3D excepted).  Once per step
 * dmgs_sh_grad_expand forms the gradient rows from the records of all V views (records + v * view_stride
 * floats; campos_host [V,3] HOST floats: the cameras' centres), stores (accumulate = 0) or adds
 * (accumulate = 1) them, and ADDS the view-direction term to dL_dmeans3D (shs = the coefficients):
 * V * 16 B + one coefficient row + one gradient row per Gaussian and step instead of V reads of the
 * coefficients and V read-modify-writes of the gradient row.                                      */
#define DMGS_MAX_STEP_VIEWS 32
int dmgs_sh_grad_expand(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t sh_layout, int32_t n_views,
                        const float *campos_host, const float *means3D, const float *shs, const float *records,
                        int64_t view_stride, float *dL_dshs, float *dL_dmeans3D, int32_t accumulate, void *stream);

/* ---- markVisible (K10): visible[i] = view-space z > 0.2 */
int dmgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                      uint8_t *visible, void *stream);

/* ---- mesh-face -> Gaussian binding (scene/gaussian_geo_model_mlp_flex.py:267-311, :370-385;
 *      stage-3 frame scene/gaussian_geo_model_finetune.py:414-421).
 * faces: int64 [F,3]; bc: [k,3]; g: device pointer to ONE float = tanh(scale_factor) * max_scale
 * (computed by the caller with two tiny torch ops, so no host read-back; NULL means g = 1, the
 * COLMAP variant scene/gaussian_geo_model_mlp_flex_colmap.py:500-504).
 * Outputs xyz [F*k,3], cov6 [F*k,6] (either may be NULL); rot_t2w [F,9] optional.           */
int dmgs_bind_forward(int64_t F, int32_t k, const float *verts, const int64_t *faces, const float *bc,
                      float rad_base, float thin_z, const float *g, int32_t adaptive, float *xyz, float *cov6,
                      float *rot_t2w, void *stream);
/* The gradient the reference's autograd produces (cov3D_L is a constant): accumulates into
 * dverts [V,3] and dg [1] (both must be zeroed by the caller). dL_dxyz / dL_dcov6 may be NULL;
 * dL_drot [F,9] is an optional extra gradient on the face frame (stage 3). */
int dmgs_bind_backward(int64_t F, int32_t k, const float *verts, const int64_t *faces, const float *bc,
                       float rad_base, float thin_z, const float *g, int32_t adaptive, const float *dL_dxyz,
                       const float *dL_dcov6, const float *dL_drot, float *dverts, float *dg, void *stream);

/* ---- binding FUSED INTO preprocess (stage-2 mode; SURVEY.md K1 "fused bind+preprocess", K8/K9 "scatter to verts"):
 *      scene/gaussian_geo_model_mlp_flex.py:267-311 + :370-385 feeding gaussian_renderer/__init__.py:178-186 in ONE
 *      pass per direction.  Gaussian i = face i / k, barycentric row i % k (face-major, as the reference);
 *      prm->P must equal F * k.  The mean and Sigma are built in registers (same arithmetic, same bits as
 *      dmgs_bind_forward) and go straight into the EWA projection + SH colour: no xyz[P,3] / cov6[P,6] round trip.
 * forward: everything dmgs_preprocess_forward does (state in `geom`, radii, num_rendered; then dmgs_bin_forward /
 *      dmgs_blend_forward as usual); xyz_out [P,3] optional (the texture MLP consumes the means).
 * backward (after dmgs_blend_backward filled `scratch`): dL/dmean and dL/dSigma never reach memory: they pass through
 *      the binding adjoint (the reference's truncated gradient: cov3D_L constant) and are ADDED with reductions to
 *      dverts [V,3] and dg [1] (dg may be NULL; the caller zeroes both).  dL_dmeans2D [P,3], dL_dopacity [P],
 *      dL_dcolors_precomp [P,3] / dL_dshs (layout of shs) as dmgs_preprocess_backward (accumulate 0: store, 1: add). */
int dmgs_preprocess_forward_bound(const dmgs_params *prm, int64_t F, int32_t k, const float *verts, const int64_t *faces,
                                  const float *bc, float rad_base, float thin_z, const float *g, int32_t adaptive,
                                  const float *opacities, const float *shs, const float *colors_precomp, int32_t *radii,
                                  void *geom, uint32_t *num_rendered, float *xyz_out, void *stream);
int dmgs_preprocess_backward_bound(const dmgs_params *prm, int64_t F, int32_t k, const float *verts, const int64_t *faces,
                                   const float *bc, float rad_base, float thin_z, const float *g, int32_t adaptive,
                                   const float *shs, const int32_t *radii, const void *geom, const void *scratch,
                                   float *dverts, float *dg, float *dL_dmeans2D, float *dL_dopacity,
                                   float *dL_dcolors_precomp, float *dL_dshs, int32_t accumulate, void *stream);

/* ---- stage-3 binding, per-Gaussian part (scene/gaussian_geo_model_finetune.py:446-453 get_scaling,
 *      :456-463 get_rotation, :465-482 get_rot_matrix, :501-516 get_covariance).
 * rot_t2w [F,9]: the face frames (dmgs_bind_forward's rot_t2w output); rotation2d, scaling2d [F*k,2]:
 * the per-Gaussian parameters `_rotation`, `_scaling` (face-major order f*k + j); thin_z: the constant
 * third scale.  Outputs (any may be NULL): scales [F*k,3] = (exp s0, exp s1, thin_z); quats [F*k,4] =
 * normalize(matrix_to_quaternion(R)) (w,x,y,z), R = rot_t2w [[a,-b,0],[b,a,0],[0,0,1]], (a,b) =
 * normalize(rotation2d); cov6 [F*k,6] = strip_symmetric((R S)(R S)^T).
 * backward: upstream gradients dL_dscales / dL_dquats / dL_dcov6 (any may be NULL) -> dL_drot [F,9]
 * (overwritten; feed it to dmgs_bind_backward's dL_drot), dL_drotation2d, dL_dscaling2d [F*k,2]
 * (overwritten).                                                                             */
int dmgs_stage3_forward(int64_t F, int32_t k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                        float thin_z, float *scales, float *quats, float *cov6, void *stream);
int dmgs_stage3_backward(int64_t F, int32_t k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                         float thin_z, const float *dL_dscales, const float *dL_dquats, const float *dL_dcov6,
                         float *dL_drot, float *dL_drotation2d, float *dL_dscaling2d, void *stream);

/* ==== rows next to the path ========================================== */

/* ---- fused L1 + SSIM image loss (utils/loss_utils.py:17-18 l1_loss, :35-63 ssim/_ssim; caller
 *      train_geo_stage2.py:115-116).  img, gt: [planes,H,W] fp32 (planes = batch x channels);
 * window11_host: the 11 normalised Gaussian taps (loss_utils.py:23-25, sigma 1.5) as HOST floats.
 * forward: out_means[planes][2] = {mean |img-gt|, mean ssim_map} per plane; fills `scratch`
 * (dmgs_l1_ssim_scratch_bytes) with the derivative fields the backward convolves.
 * backward: upstream[planes][2] (DEVICE) = dLoss/d(l1 mean), dLoss/d(ssim mean) per plane, already
 * divided by H*W; grad [planes,H,W] is overwritten.                                          */
size_t dmgs_l1_ssim_scratch_bytes(int32_t planes, int32_t H, int32_t W);
int dmgs_l1_ssim_forward(int32_t planes, int32_t H, int32_t W, const float *window11_host, const float *img,
                         const float *gt, void *scratch, float *out_means, void *stream);
int dmgs_l1_ssim_backward(int32_t planes, int32_t H, int32_t W, const float *window11_host, const float *img,
                          const float *gt, const void *scratch, const float *upstream, float *grad, void *stream);

/* ---- view-frustum test + ordered compaction of the visible faces
 *      (in_frustum: scene/gaussian_geo_model_finetune.py:33-48, COLMAP variant
 *      scene/gaussian_geo_model_mlp_flex_colmap.py:32-76; use at finetune.py:405-409).
 * proj16_host: the [4,4] full_proj_transform, row-major, HOST floats (points are right-multiplied).
 * has_cube / cube_len / piece_id / n_piece: the COLMAP variant's arguments (has_cube = 0, piece_id = -1
 * for the stage-3 variant).  faces == NULL: `pts` [N,3] are the query points themselves; otherwise
 * the query points are the centroids verts[faces].mean(1) of the N int64 faces over pts = verts.
 * mask: uint8 [N] (torch.bool layout).  Compaction (optional, count_out != NULL): count_out (DEVICE
 * int32) = number of visible items, faces_out [count,3] = faces[mask] (needs faces), index_out
 * [count] = their indices (either may be NULL); scratch: dmgs_frustum_scratch_bytes(N).       */
size_t dmgs_frustum_scratch_bytes(int64_t N);
int dmgs_in_frustum(int64_t N, const float *proj16_host, float cube_len, int32_t has_cube, int32_t piece_id,
                    int32_t n_piece, const float *pts, const int64_t *faces, uint8_t *mask, int64_t *faces_out,
                    int32_t *index_out, int32_t *count_out, void *scratch, void *stream);

/* ---- fused multi-tensor Adam step (torch.optim.Adam(l, lr=0.0, eps=1e-15) with one group per tensor:
 *      scene/gaussian_geo_model_finetune.py:526-537, step at train_geo_stage3.py:164-166).
 * One launch updates up to DMGS_ADAM_MAX_SEGMENTS tensors; all arrays fp32, 16-byte aligned.
 * Elements with (i % period) >= split use lr_hi (period == 0: lr for every element) -- e.g. the
 * [P,16,3] SH tensor whose DC and rest coefficients have different learning rates.
 * grad is read as grad * grad_scale (view averaging) and set to zero afterwards when zero_grad != 0.
 * step counts from 1 (torch's state['step'] after the increment).                             */
#define DMGS_ADAM_MAX_SEGMENTS 8
typedef struct dmgs_adam_segment {
    float *param, *grad, *exp_avg, *exp_avg_sq;
    int64_t n;
    double lr, lr_hi;
    int32_t period, split;
} dmgs_adam_segment;
int dmgs_adam_step(int32_t nseg, const dmgs_adam_segment *segments_host, double beta1, double beta2, double eps,
                   int64_t step, float grad_scale, int32_t zero_grad, void *stream);

/* ---- stage-2 colour field: multiresolution hash-grid encoding + bias-free MLP (replaces
 *      geo/texture.py:99-111 `MLPTexture3D.sample_noact`, i.e. tinycudann's HashGrid encoder [16 levels x 2 features,
 *      2^19 entries per level, base resolution 16, finest 4096; geo/texture.py:50-72] followed by the reference's
 *      torch `_MLP` [Linear(32,32) ReLU Linear(32,32) ReLU Linear(32,channels), no bias; geo/texture.py:18-41];
 *      consumer scene/gaussian_geo_model_mlp_flex.py:313).
 * grid: dmgs_texture_grid_params() fp32 parameters (level-major, entry, feature), which the kernels read through an
 * fp16 copy made by dmgs_texture_cast_params (the encoder's parameters and output are fp16 in the reference).
 * forward: xyz [N,3] -> out [N,channels] (channels a multiple of 4, <= 64); enc_out (optional, N*32 halfs) keeps the
 * encoder output for the backward.  backward: dW0/dW1/dW2 and d_grid (optional) are ADDED to (the caller zeroes
 * them); d_grid is multiplied by grid_grad_scale (the net effect of the reference's two backward hooks is 128 on
 * the encoder parameters and 1 everywhere else, geo/texture.py:30-31, :69-71); d_xyz (optional, [N,3]) is written;
 * scratch: dmgs_texture_backward_scratch_bytes(N) bytes (dL/d(encoding) between the two backward kernels).
 * aabb6_host: HOST array {min x,y,z, max x,y,z}; coordinates are normalised and clamped to [0,1] (:102-103). */
int64_t dmgs_texture_grid_params(void);
int dmgs_texture_cast_params(int64_t n_params, const float *params, void *params_half, void *stream);
int dmgs_texture_forward(int64_t N, int32_t channels, const float *aabb6_host, const float *xyz, const void *grid_half,
                         const float *W0, const float *W1, const float *W2, float *out, void *enc_out, void *stream);
int dmgs_texture_backward(int64_t N, int32_t channels, const float *aabb6_host, const float *xyz, const void *grid_half,
                          const void *enc, const float *W0, const float *W1, const float *W2, const float *dL_dout,
                          float grid_grad_scale, float *d_grid, float *dW0, float *dW1, float *dW2, float *d_xyz,
                          void *scratch, void *stream);
size_t dmgs_texture_backward_scratch_bytes(int64_t N);

/* ---- all-reduce (sum) of a flat fp32 buffer over NVLink peer memory (view-partitioned training step:
 *      the per-Gaussian gradient exchange; replaces ncclAllReduce on one box).
 * Every rank passes the same-sized buffer of n floats (n % 4 == 0, 16-byte aligned), mapped into every
 * peer's address space (CUDA IPC / symmetric memory): peer_ptrs_host[k] = rank k's buffer as seen from
 * THIS process (HOST array of `world` device pointers, own buffer at index `rank`).  multicast_ptr, when
 * not NULL, is the NVSwitch multicast mapping of the same buffers: the reduction then happens in the
 * switch (multimem.ld_reduce / multimem.st).  Rank r reduces slice r of everybody's buffer and writes
 * the scaled sum back to slice r of everybody's buffer, in place.  The caller must place a device-side
 * barrier across the ranks before the call (all buffers complete) and after it (all slices delivered). */
#define DMGS_MAX_PEERS 8
int dmgs_allreduce_peer(int64_t n, int32_t world, int32_t rank, const void *const *peer_ptrs_host, void *multicast_ptr,
                        float scale, void *stream);

/* ---- the gradient exchange FUSED with the optimiser step (view-partitioned training; replaces
 *      dmgs_allreduce_peer followed by dmgs_adam_step on every rank -- i.e. ncclAllReduce + torch.optim.Adam.step of
 *      train_geo_stage3.py:164-166 run data parallel).  Rank r owns slice r (dmgs_adam_exchange_shard) of every
 *      parameter tensor: ONE kernel sums that slice of all ranks' gradients (NVSwitch multimem.ld_reduce when the
 *      multicast pointers are given, peer loads otherwise), applies Adam with the rank's SHARD of exp_avg / exp_avg_sq
 *      (arrays of 4 * (end4 - begin4) floats) and writes the new parameters into every rank's parameter buffer
 *      (multimem.st / peer stores).  Parameters and gradients live in two flat buffers mapped into every rank (as for
 *      dmgs_allreduce_peer); a segment = one tensor: its offsets (in floats, multiples of 4) in the two buffers and
 *      its n elements; the ceil(n/4)*4 - n padding floats of a field are updated too (they stay zero).  The caller
 *      brackets the call with two device-side barriers, as for dmgs_allreduce_peer.  Gradients are NOT cleared. */
typedef struct dmgs_adam_xsegment {
    int64_t grad_offset, param_offset, n;
    float *exp_avg_shard, *exp_avg_sq_shard;
    double lr, lr_hi;
    int32_t period, split;
} dmgs_adam_xsegment;
int dmgs_adam_exchange_shard(int64_t n, int32_t world, int32_t rank, int64_t *begin4, int64_t *end4);
int dmgs_adam_exchange_peer(int32_t world, int32_t rank, int32_t nseg, const dmgs_adam_xsegment *segments_host,
                            const void *const *grad_peer_ptrs_host, void *grad_multicast_ptr,
                            const void *const *param_peer_ptrs_host, void *param_multicast_ptr, double beta1, double beta2,
                            double eps, int64_t step, float grad_scale, void *stream);


/* ---- inspection (parity tests): byte offsets of the named arrays inside the state buffers.
 * geom:    [0] depths f32[P]  [1] rec f32[P][8]={x,y,conA,conB,conC,opacity,cut,_}  [2] rgb f32[P][4]
 *          [3] clamped u8[P] (bit ch)  [4] cov3D f32[P][6]  [5] tiles_touched u32[P]
 *          [6] rect u16[P][4]={x0,x1,y0,y1}  [7] sort buffer A of Gaussian indices u32[P] (the
 *          depth order after an even number of executed passes)  [8] instance offsets (exclusive
 *          scan in depth order; radix tile partition only) u32[P]
 * binning: [0] sorted tile ids u32[R] (materialised by dmgs_sorted_keys on the placement path)
 *          [1] sorted Gaussian indices u32[R]  [2] ranges u32[T][2]
 * image:   [0] final_T f32[H*W]  [1] n_contrib u32[H*W]  (the buffer ends with 256 bytes of launch
 *          state: the square counter of the persistent forward blend)                           */
int dmgs_geom_layout(int32_t P, int64_t *offsets9);
int dmgs_binning_layout(int32_t P, int64_t num_rendered, int32_t W, int32_t H, int64_t *offsets3);
int dmgs_image_layout(int32_t W, int32_t H, int64_t *offsets2);
/* Materialises the 64-bit sort keys (tile << 32 | depth bits) of the sorted instance list. */
int dmgs_sorted_keys(const void *geom, const void *binning, int32_t P, int64_t num_rendered, int32_t W, int32_t H,
                     uint64_t *keys_out, void *stream);
/* y[i] = the library's exp() (the arithmetic-contract exp used for alpha). */
int dmgs_exp_array(const float *x, float *y, int64_t n, void *stream);

const char *dmgs_last_error(void);
int dmgs_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (monotonic). */
uint64_t dmgs_launch_count(void);
/* Residency of the two blend kernels (K6, K7): their grids are persistent, at most `forward` / `backward` CTAs of
 * 128 threads per multiprocessor (1..8; 0 leaves a value unchanged; default 8 = every register of the SM, or the
 * environment variables DMGS_BLEND_FWD_RESIDENCY / DMGS_BLEND_BWD_RESIDENCY read at load time).  Lower values leave
 * registers and warp slots to kernels of OTHER streams (several views in flight); results do not depend on it.
 * No upstream counterpart (upstream launches one CTA per tile).  Returns 0, or -7 for values outside 0..8. */
int dmgs_set_blend_residency(int32_t forward, int32_t backward);
/* Shared memory per multiprocessor the tile-placement kernels (K3-K5: per-warp tile counters) run with, 64..200 KB
 * (default 200 = shortest walk for a frame rendered alone; env DMGS_PLACE_SMEM_KB at load time).  `ViewStreams` uses
 * 128 KB while several views are in flight: six blend CTAs of another view then fit beside a placement kernel.  Buffer
 * layouts do not depend on it (they are sized for 200 KB), so it may change between any two calls; the tile lists are
 * identical for every value.  Returns 0, or -7 outside 64..200. */
int dmgs_set_place_smem_kb(int32_t kb);

#ifdef __cplusplus
}
#endif
#endif /* DMGS_RASTER_H */
