// kernels.cuh -- launcher declarations shared by the translation units of libdmgs_raster.so.
#pragma once
#include "common.cuh"

namespace dmgs {

DevParams make_dev_params(const dmgs_params *p);

// preprocess.cu
struct BindSrc;  // bind_math.cuh: mesh source of the fused bind + preprocess path (NULL: plain per-Gaussian inputs)
int launch_preprocess_fwd(const dmgs_params *prm, const BindSrc *bind, float *xyz_out, const float *means3D,
                          const float *scales, const float *rotations,
                          const float *cov3D_precomp, const float *opacities, const float *shs,
                          const float *colors_precomp, int32_t *radii, void *geom, const GeomLayout &L,
                          uint32_t *total_instances, uint32_t *key_stat, cudaStream_t s);
int launch_preprocess_bwd(const dmgs_params *prm, const BindSrc *bind, const float *means3D, const float *scales,
                          const float *rotations,
                          const float *cov3D_precomp, const float *shs, const int32_t *radii, const void *geom,
                          const GeomLayout &L, const float *grad_blend, float *dL_dmeans3D, float *dL_dmeans2D,
                          float *dL_dopacity, float *dL_dcolprec, float *dL_dshs, float *dL_dscales, float *dL_drots,
                          float *dL_dcov3D, int accumulate, cudaStream_t s);
int launch_mark_visible(int P, const float *means3D, const float *view_dev, uint8_t *visible, cudaStream_t s);
int launch_exp_array(const float *x, float *y, int64_t n, cudaStream_t s);

// sort.cu
// Stable LSD radix pass number `pass` over (key,value) pairs on `bits` bits starting at `shift`; buffers
// alternate A -> B -> A ...  `stat` (device, optional): {OR of keys, OR of ~keys} enables the adaptive
// depth sort that skips digits which do not vary (sort.cu).
int radix_pass(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int pass, int shift,
               int bits, uint32_t *hist, const uint32_t *stat, cudaStream_t s);
// Stable depth sort of n (key, value) pairs, 8-bit digits, onesweep: one histogram kernel + one kernel per digit
// with decoupled look-back between the tiles (sort.cu).  ghist (4 x 256 words) and counters (4 words) must be zero
// on entry (stat != NULL: the caller clears them together with stat before the keys are produced; stat == NULL:
// cleared here).  Result in (keys_a, vals_a) after an even number of executed passes, else in (keys_b, vals_b);
// without stat all four passes run (result in A).
int depth_sort_onesweep(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                        uint32_t *ghist, uint32_t *counters, uint32_t *status, const uint32_t *stat, cudaStream_t s);
// out[i] = sum_{j<i} in[gather ? gather[j] : j]; *total = sum of all (may be NULL)
int exclusive_scan_u32(const uint32_t *in, const uint32_t *gather, uint32_t *out, int64_t n, uint32_t *total,
                       uint32_t *scan_tmp, cudaStream_t s);
int launch_emit_instances(int P, const uint32_t *order, const uint32_t *offsets, const uint32_t *tiles,
                          const uint2 *rect, int gx, uint32_t *inst_tile, uint32_t *inst_gidx, cudaStream_t s);
int launch_tile_ranges(int64_t R, const uint32_t *sorted_tiles, uint2 *ranges, int T, cudaStream_t s);
int launch_sorted_keys(int64_t R, const uint32_t *sorted_tiles, const uint32_t *sorted_gidx, const float *depths,
                       uint64_t *keys_out, cudaStream_t s);

// place.cu
int launch_tile_placement(const PlacePlan &pl, int P, int T, int gx, const uint32_t *order_a, const uint32_t *order_b,
                          const uint32_t *stat, const uint2 *rect,
                          uint4 *srec, uint32_t *table, uint32_t *gsum, uint32_t *tile_start, uint2 *ranges,
                          uint32_t *out_gidx, int64_t capacity, uint32_t *overflow, uint32_t *tile_order, cudaStream_t s);
int launch_fill_tiles(int T, const uint2 *ranges, uint32_t *sorted_tiles, cudaStream_t s);

// blend.cu
int launch_blend_fwd(const dmgs_params *prm, const void *geom, const GeomLayout &GL, const void *binning,
                     const BinLayout &BL, float *out_color, void *image, const ImgLayout &IL, cudaStream_t s);
int launch_blend_bwd(const dmgs_params *prm, const void *geom, const GeomLayout &GL, const void *binning,
                     const BinLayout &BL, const void *image, const ImgLayout &IL, const float *dL_dpix,
                     float *grad_blend, cudaStream_t s);

int launch_sh_grad_expand(int P, int sh_degree, int M, int layout, int V, const float *campos_host, const float *means3D,
                          const float *shs, const float *records, int64_t view_stride, float *dL_dshs,
                          float *dL_dmeans3D, int accumulate, cudaStream_t s);

// stage3.cu
int launch_stage3_fwd(int64_t F, int k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                      float thin_z, float *scales, float *quats, float *cov6, cudaStream_t s);
int launch_stage3_bwd(int64_t F, int k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                      float thin_z, const float *dL_dscales, const float *dL_dquats, const float *dL_dcov6, float *dL_drot,
                      float *dL_drotation2d, float *dL_dscaling2d, cudaStream_t s);

// loss.cu
size_t loss_scratch_bytes(int planes, int H, int W);
int launch_l1_ssim_fwd(int planes, int H, int W, const float *window11, const float *img, const float *gt,
                       void *scratch, float *out_means, cudaStream_t s);
int launch_l1_ssim_bwd(int planes, int H, int W, const float *window11, const float *img, const float *gt,
                       const void *scratch, const float *upstream, float *grad, cudaStream_t s);

// frustum.cu
size_t frustum_scratch_bytes(int64_t N);
int launch_frustum(int64_t N, const float *proj16_host, float cube_len, int has_cube, int piece_id, int n_piece,
                   const float *pts, const int64_t *faces, uint8_t *mask, int64_t *faces_out, int32_t *index_out,
                   int32_t *count_out, void *scratch, cudaStream_t s);

// adam.cu
int launch_adam(int nseg, const dmgs_adam_segment *segs, double beta1, double beta2, double eps, int64_t step,
                float grad_scale, int zero_grad, cudaStream_t s);

// peer.cu
int launch_allreduce_peer(int64_t n, int world, int rank, const void *const *peer_ptrs_host, void *multicast_ptr,
                          float scale, cudaStream_t s);

void adam_exchange_shard(int64_t n, int world, int rank, int64_t *begin4, int64_t *end4);
int launch_adam_exchange(int world, int rank, int nseg, const dmgs_adam_xsegment *segs, const void *const *grad_peers_host,
                         void *grad_multicast, const void *const *param_peers_host, void *param_multicast, double beta1,
                         double beta2, double eps, int64_t step, float grad_scale, cudaStream_t s);

// texture.cu
int64_t texture_grid_params();
int launch_texture_cast(int64_t n_params, const float *params, void *params_half, cudaStream_t s);
int launch_texture_fwd(int64_t N, int C, const float *aabb6_host, const float *xyz, const void *grid_half, const float *W0,
                       const float *W1, const float *W2, float *out, void *enc_out, cudaStream_t s);
int launch_texture_bwd(int64_t N, int C, const float *aabb6_host, const float *xyz, const void *grid_half, const void *enc,
                       const float *W0, const float *W1, const float *W2, const float *dL_dout, float grid_grad_scale,
                       float *d_grid, float *dW0, float *dW1, float *dW2, float *d_xyz, void *scratch, cudaStream_t s);
size_t texture_bwd_scratch_bytes(int64_t N);

// binding.cu
int launch_bind_fwd(int64_t F, int k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                    float thin_z, const float *g, int adaptive, float *xyz, float *cov6, float *rot_t2w, cudaStream_t s);
int launch_bind_bwd(int64_t F, int k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                    float thin_z, const float *g, int adaptive, const float *dL_dxyz, const float *dL_dcov6,
                    const float *dL_drot, float *dverts, float *dg, cudaStream_t s);

}  // namespace dmgs
