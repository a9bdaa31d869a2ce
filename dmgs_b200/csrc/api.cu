// api.cu -- the extern "C" surface declared in include/dmgs_raster.h.
#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"
#include "kernels.cuh"
#include "bind_math.cuh"

namespace dmgs {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((uint64_t)n); }

static int env_residency(const char *name)
{
    const char *v = getenv(name);
    const int k = v ? atoi(v) : 0;
    return k >= 1 && k <= 8 ? k : 8;
}
int g_blend_residency[2] = {env_residency("DMGS_BLEND_FWD_RESIDENCY"), env_residency("DMGS_BLEND_BWD_RESIDENCY")};

static int current_device()
{
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}
int num_sms()
{
    static std::atomic<int> cache[64];
    const int d = current_device();
    if (d < 0 || d >= 64) return DMGS_DEFAULT_SMS;
    int n = cache[d].load();
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = DMGS_DEFAULT_SMS;
        }
        cache[d].store(n);
    }
    return n;
}
bool once_per_device(int slot)
{
    static std::atomic<bool> seen[ONCE_SLOTS][64];
    const int d = current_device();
    if (d < 0 || d >= 64) return true;  // unknown device: set the attributes every time (cheap)
    return !seen[slot][d].exchange(true);
}

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_stage(const dmgs_params *prm, cudaStream_t s, const char *stage)
{
    if (!prm || !prm->debug) return 0;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("stage '%s' failed: %s", stage, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

static int tile_bits(int T)
{
    int b = 0;
    while ((1 << b) < T) ++b;
    return b;
}

static int validate(const dmgs_params *p)
{
    if (!p) { set_error("params is NULL"); return -1; }
    if (p->P < 0 || p->image_width <= 0 || p->image_height <= 0) { set_error("bad sizes P=%d W=%d H=%d", p->P, p->image_width, p->image_height); return -2; }
    if (p->image_width > 65535 * DMGS_TILE || p->image_height > 65535 * DMGS_TILE) { set_error("image too large"); return -2; }
    if (p->sh_degree < 0 || p->sh_degree > 3) { set_error("sh_degree %d not in 0..3", p->sh_degree); return -3; }
    return 0;
}

// the SH rows must hold at least the coefficients of the active degree
static int validate_sh(const dmgs_params *p, const float *shs)
{
    if (shs && p->sh_coeffs < (p->sh_degree + 1) * (p->sh_degree + 1)) {
        set_error("sh_coeffs %d < (sh_degree + 1)^2 = %d", p->sh_coeffs, (p->sh_degree + 1) * (p->sh_degree + 1));
        return -3;
    }
    return 0;
}

}  // namespace dmgs

using namespace dmgs;

extern "C" {

int dmgs_abi_version(void) { return 1; }
uint64_t dmgs_launch_count(void) { return g_launches.load(); }
int dmgs_set_place_smem_kb(int32_t kb)
{
    if (kb < 64 || kb > 200) { set_error("placement shared memory: 64..200 KB per SM"); return -7; }
    g_place_smem_kb = kb;
    return 0;
}
int dmgs_set_blend_residency(int32_t forward, int32_t backward)
{
    if (forward < 0 || forward > 8 || backward < 0 || backward > 8) { set_error("blend residency: 0..8 CTAs per SM"); return -7; }
    if (forward) g_blend_residency[0] = forward;
    if (backward) g_blend_residency[1] = backward;
    return 0;
}
const char *dmgs_last_error(void) { return g_err; }

size_t dmgs_geom_bytes(int32_t P) { return geom_layout(P).total; }
size_t dmgs_binning_bytes(int32_t P, int64_t R, int32_t W, int32_t H) { return bin_layout(P, R, W, H).total; }
size_t dmgs_image_bytes(int32_t W, int32_t H) { return img_layout(W, H).total; }
// P x 12 sums of the blend backward + the square counter of its persistent grid (zeroed together)
size_t dmgs_backward_scratch_bytes(int32_t P) { return align_up((size_t)(P > 0 ? P : 1) * 12 * sizeof(float) + 16); }

int dmgs_geom_layout(int32_t P, int64_t *o)
{
    const GeomLayout L = geom_layout(P);
    o[0] = L.depths; o[1] = L.rec; o[2] = L.rgb; o[3] = L.clamped; o[4] = L.cov3D; o[5] = L.tiles; o[6] = L.rect;
    o[7] = L.order; o[8] = L.offsets;
    return 0;
}
int dmgs_binning_layout(int32_t P, int64_t R, int32_t W, int32_t H, int64_t *o)
{
    const BinLayout L = bin_layout(P, R, W, H);
    o[0] = L.tiles; o[1] = L.gidx; o[2] = L.ranges;
    return 0;
}
int dmgs_image_layout(int32_t W, int32_t H, int64_t *o)
{
    const ImgLayout L = img_layout(W, H);
    o[0] = L.final_T; o[1] = L.n_contrib;
    return 0;
}

static int preprocess_forward_impl(const dmgs_params *prm, const BindSrc *bind, float *xyz_out, const float *means3D,
                                   const float *scales, const float *rotations,
                                   const float *cov3D_precomp, const float *opacities, const float *shs,
                                   const float *colors_precomp, int32_t *radii, void *geom, uint32_t *num_rendered,
                                   void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if ((shs == nullptr) == (colors_precomp == nullptr)) { set_error("provide exactly one of shs / colors_precomp"); return -4; }
    if (!bind && ((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr)) {
        set_error("provide exactly one of (scales, rotations) / cov3D_precomp");
        return -5;
    }
    if (!num_rendered) { set_error("num_rendered is NULL"); return -6; }
    if ((rc = validate_sh(prm, shs))) return rc;
    const int P = prm->P;
    if (P == 0) {
        DMGS_CUDA(cudaMemsetAsync(num_rendered, 0, sizeof(uint32_t), s));
        return 0;
    }
    if ((!bind && !means3D) || !opacities || !radii || !geom) { set_error("NULL required pointer"); return -6; }
    const GeomLayout L = geom_layout(P);
    // with direct tile placement (place.cu) only the TOTAL of tiles-touched is needed: preprocess reduces
    // it on the fly; the radix tile partition also needs the per-Gaussian offsets (scan below)
    const int T = ((prm->image_width + DMGS_TILE - 1) / DMGS_TILE) * ((prm->image_height + DMGS_TILE - 1) / DMGS_TILE);
    const bool placed = place_plan(P, T).ok != 0;
    uint32_t *stat = placed ? at<uint32_t>(geom, L.stat) : nullptr;
    if (placed) {
        DMGS_CUDA(cudaMemsetAsync(num_rendered, 0, sizeof(uint32_t), s));
        DMGS_CUDA(cudaMemsetAsync(stat, 0, L.os_status - L.stat, s));  // key range, tile counters, digit histograms
    }
    rc = launch_preprocess_fwd(prm, bind, xyz_out, means3D, scales, rotations, cov3D_precomp, opacities, shs,
                               colors_precomp, radii, geom, L, placed ? num_rendered : nullptr, stat, s);
    if (rc) return rc;
    if ((rc = check_stage(prm, s, "preprocess"))) return rc;
    // stable depth sort of the Gaussians: 8-bit LSD passes over (keys_a, order) <-> (keys_b, vals_b).
    // On the placement path the passes are adaptive (digits in which no two live keys differ are
    // skipped on the device, typically the top byte); otherwise all four run and the result is in A.
    uint32_t *ka = at<uint32_t>(geom, L.keys_a), *kb = at<uint32_t>(geom, L.keys_b);
    uint32_t *va = at<uint32_t>(geom, L.order), *vb = at<uint32_t>(geom, L.vals_b);
    uint32_t *tmp = at<uint32_t>(geom, L.scan_tmp);
    rc = depth_sort_onesweep(ka, va, kb, vb, P, at<uint32_t>(geom, L.os_ghist), at<uint32_t>(geom, L.stat) + 4,
                             at<uint32_t>(geom, L.os_status), stat, s);
    if (rc) return rc;
    if ((rc = check_stage(prm, s, "depth sort"))) return rc;
    if (placed) return 0;
    rc = exclusive_scan_u32(at<uint32_t>(geom, L.tiles), va, at<uint32_t>(geom, L.offsets), P, num_rendered, tmp, s);
    if (rc) return rc;
    return check_stage(prm, s, "tile scan");
}

int dmgs_preprocess_forward(const dmgs_params *prm, const float *means3D, const float *scales, const float *rotations,
                            const float *cov3D_precomp, const float *opacities, const float *shs,
                            const float *colors_precomp, int32_t *radii, void *geom, uint32_t *num_rendered,
                            void *stream)
{
    return preprocess_forward_impl(prm, nullptr, nullptr, means3D, scales, rotations, cov3D_precomp, opacities, shs,
                                   colors_precomp, radii, geom, num_rendered, stream);
}

static int make_bind_src(const dmgs_params *prm, int64_t F, int32_t k, const float *verts, const int64_t *faces,
                         const float *bc, float rad_base, float thin_z, const float *g, int32_t adaptive, BindSrc *bs)
{
    if (!prm) { set_error("params is NULL"); return -1; }
    if (F < 0 || k <= 0 || F * (int64_t)k != (int64_t)prm->P) {
        set_error("bound preprocess: P = %d must equal F * k = %lld * %d", prm->P, (long long)F, k);
        return -2;
    }
    if (F > 0 && (!verts || !faces || !bc)) { set_error("bound preprocess: NULL mesh pointer"); return -6; }
    bs->verts = verts; bs->faces = faces; bs->bc = bc; bs->g_ptr = g;
    bs->rad_base = rad_base; bs->thin_z = thin_z; bs->k = k; bs->adaptive = adaptive;
    return 0;
}

int dmgs_preprocess_forward_bound(const dmgs_params *prm, int64_t F, int32_t k, const float *verts, const int64_t *faces,
                                  const float *bc, float rad_base, float thin_z, const float *g, int32_t adaptive,
                                  const float *opacities, const float *shs, const float *colors_precomp, int32_t *radii,
                                  void *geom, uint32_t *num_rendered, float *xyz_out, void *stream)
{
    BindSrc bs;
    int rc = make_bind_src(prm, F, k, verts, faces, bc, rad_base, thin_z, g, adaptive, &bs);
    if (rc) return rc;
    return preprocess_forward_impl(prm, &bs, xyz_out, nullptr, nullptr, nullptr, nullptr, opacities, shs, colors_precomp,
                                   radii, geom, num_rendered, stream);
}

static int bin_forward_impl(const dmgs_params *prm, const void *geom, int64_t R, void *binning, uint32_t *overflow,
                            bool async, void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int P = prm->P, W = prm->image_width, H = prm->image_height;
    const int gx = (W + DMGS_TILE - 1) / DMGS_TILE, gy = (H + DMGS_TILE - 1) / DMGS_TILE, T = gx * gy;
    if (R < 0 || R >= ((int64_t)1 << 31)) { set_error("num_rendered %lld out of range", (long long)R); return -7; }
    if (!binning) { set_error("binning buffer is NULL"); return -6; }
    const GeomLayout GL = geom_layout(P);
    const BinLayout BL = bin_layout(P, R, W, H);
    const PlacePlan pl = place_plan_run(P, T);
    if (pl.ok) {
        // direct placement from the depth-sorted Gaussian order (place.cu); ranges fall out of the scan
        if (R > 0 && P > 0) {
            rc = launch_tile_placement(pl, P, T, gx, at<uint32_t>(geom, GL.order), at<uint32_t>(geom, GL.vals_b),
                                       at<uint32_t>(geom, GL.stat), at<uint2>(geom, GL.rect),
                                       at<uint4>(binning, BL.srec), at<uint32_t>(binning, BL.table),
                                       at<uint32_t>(binning, BL.gsum), at<uint32_t>(binning, BL.tile_start),
                                       at<uint2>(binning, BL.ranges), at<uint32_t>(binning, BL.gidx), R, overflow,
                                       at<uint32_t>(binning, BL.tile_order), s);
            if (rc) return rc;
        } else {
            DMGS_CUDA(cudaMemsetAsync(at<uint2>(binning, BL.ranges), 0, sizeof(uint2) * (size_t)T, s));
            if (overflow) DMGS_CUDA(cudaMemsetAsync(overflow, 0, sizeof(uint32_t), s));
        }
        return check_stage(prm, s, "tile placement");
    }
    if (async) {
        set_error("dmgs_bin_forward_async needs the direct-placement path (<= %d tiles); read num_rendered back "
                  "and call dmgs_bin_forward", PLACE_MAX_TILES);
        return -9;
    }
    uint32_t *ta = at<uint32_t>(binning, BL.tiles), *tb = at<uint32_t>(binning, BL.tiles_b);
    uint32_t *ga = at<uint32_t>(binning, BL.gidx), *gb = at<uint32_t>(binning, BL.gidx_b);
    uint32_t *hist = at<uint32_t>(binning, BL.hist), *tmp = at<uint32_t>(binning, BL.scan_tmp);
    if (R > 0 && P > 0) {
        const int tbits = tile_bits(T);
        int nbits[2] = {0, 0}, npass = 0;
        if (tbits > 8) { nbits[0] = (tbits + 1) / 2; nbits[1] = tbits - nbits[0]; npass = 2; }
        else if (tbits > 0) { nbits[0] = tbits; npass = 1; }
        // emit into the buffer that makes the last pass land in (tiles, gidx)
        uint32_t *ek = (npass & 1) ? tb : ta, *ev = (npass & 1) ? gb : ga;
        rc = launch_emit_instances(P, at<uint32_t>(geom, GL.order), at<uint32_t>(geom, GL.offsets),
                                   at<uint32_t>(geom, GL.tiles), at<uint2>(geom, GL.rect), gx, ek, ev, s);
        if (rc) return rc;
        if ((rc = check_stage(prm, s, "emit instances"))) return rc;
        // pass p reads A when p is even: make A the buffer the instances were emitted into
        uint32_t *A_k = ek, *A_v = ev, *B_k = (ek == ta) ? tb : ta, *B_v = (ev == ga) ? gb : ga;
        int shift = 0;
        for (int pass = 0; pass < npass; ++pass) {
            rc = radix_pass(A_k, A_v, B_k, B_v, R, pass, shift, nbits[pass], hist, nullptr, s);
            if (rc) return rc;
            shift += nbits[pass];
        }
        if ((rc = check_stage(prm, s, "tile partition"))) return rc;
    }
    rc = launch_tile_ranges(R, ta, at<uint2>(binning, BL.ranges), T, s);
    if (rc) return rc;
    return check_stage(prm, s, "tile ranges");
}

int dmgs_bin_forward(const dmgs_params *prm, const void *geom, int64_t R, void *binning, void *stream)
{
    return bin_forward_impl(prm, geom, R, binning, nullptr, false, stream);
}

int dmgs_bin_forward_async(const dmgs_params *prm, const void *geom, int64_t capacity, void *binning,
                           uint32_t *overflow, void *stream)
{
    if (!overflow) { set_error("overflow pointer is NULL"); return -6; }
    if (capacity < 1) { set_error("capacity must be positive"); return -7; }
    return bin_forward_impl(prm, geom, capacity, binning, overflow, true, stream);
}

int dmgs_blend_forward(const dmgs_params *prm, const void *geom, const void *binning, int64_t R, float *out_color,
                       void *image, void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    if (!out_color || !image || !binning) { set_error("NULL required pointer"); return -6; }
    cudaStream_t s = (cudaStream_t)stream;
    const GeomLayout GL = geom_layout(prm->P);
    const BinLayout BL = bin_layout(prm->P, R, prm->image_width, prm->image_height);
    const ImgLayout IL = img_layout(prm->image_width, prm->image_height);
    rc = launch_blend_fwd(prm, geom, GL, binning, BL, out_color, image, IL, s);
    if (rc) return rc;
    return check_stage(prm, s, "blend forward");
}

int dmgs_blend_backward(const dmgs_params *prm, const void *geom, const void *binning, const void *image, int64_t R,
                        const float *dL_dpix, void *scratch, void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int P = prm->P;
    if (P == 0) return 0;
    if (!geom || !binning || !image || !dL_dpix || !scratch) { set_error("NULL required pointer"); return -6; }
    const GeomLayout GL = geom_layout(P);
    const BinLayout BL = bin_layout(P, R, prm->image_width, prm->image_height);
    const ImgLayout IL = img_layout(prm->image_width, prm->image_height);
    float *grad_blend = (float *)scratch;
    DMGS_CUDA(cudaMemsetAsync(grad_blend, 0, (size_t)P * 12 * sizeof(float) + 16, s));
    if (R > 0) {
        rc = launch_blend_bwd(prm, geom, GL, binning, BL, image, IL, dL_dpix, grad_blend, s);
        if (rc) return rc;
    }
    return check_stage(prm, s, "blend backward");
}

int dmgs_preprocess_backward(const dmgs_params *prm, const float *means3D, const float *scales, const float *rotations,
                             const float *cov3D_precomp, const float *shs, const int32_t *radii, const void *geom,
                             const void *scratch, float *dL_dmeans3D, float *dL_dmeans2D, float *dL_dopacity,
                             float *dL_dcolors_precomp, float *dL_dshs, float *dL_dscales, float *dL_drotations,
                             float *dL_dcov3D, int32_t accumulate, void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int P = prm->P;
    if ((rc = validate_sh(prm, shs))) return rc;
    if (P == 0) return 0;
    if (!means3D || !radii || !geom || !dL_dmeans3D || !dL_dmeans2D || !dL_dopacity || !scratch) {
        set_error("NULL required pointer");
        return -6;
    }
    const GeomLayout GL = geom_layout(P);
    rc = launch_preprocess_bwd(prm, nullptr, means3D, scales, rotations, cov3D_precomp, shs, radii, geom, GL,
                               (const float *)scratch, dL_dmeans3D, dL_dmeans2D, dL_dopacity, dL_dcolors_precomp,
                               dL_dshs, dL_dscales, dL_drotations, dL_dcov3D, accumulate, s);
    if (rc) return rc;
    return check_stage(prm, s, "preprocess backward");
}

int dmgs_preprocess_backward_bound(const dmgs_params *prm, int64_t F, int32_t k, const float *verts, const int64_t *faces,
                                   const float *bc, float rad_base, float thin_z, const float *g, int32_t adaptive,
                                   const float *shs, const int32_t *radii, const void *geom, const void *scratch,
                                   float *dverts, float *dg, float *dL_dmeans2D, float *dL_dopacity,
                                   float *dL_dcolors_precomp, float *dL_dshs, int32_t accumulate, void *stream)
{
    int rc = validate(prm);
    if (rc) return rc;
    BindSrc bs;
    if ((rc = make_bind_src(prm, F, k, verts, faces, bc, rad_base, thin_z, g, adaptive, &bs))) return rc;
    if ((rc = validate_sh(prm, shs))) return rc;
    if (accumulate != 0 && accumulate != 1) { set_error("bound backward: accumulate must be 0 or 1"); return -6; }
    cudaStream_t s = (cudaStream_t)stream;
    const int P = prm->P;
    if (P == 0) return 0;
    if (!radii || !geom || !dverts || !dL_dmeans2D || !dL_dopacity || !scratch) { set_error("NULL required pointer"); return -6; }
    const GeomLayout GL = geom_layout(P);
    // in this mode the dL_dmeans3D / dL_dcov3D slots of the kernel carry dverts [V,3] / dg [1] (both accumulated)
    rc = launch_preprocess_bwd(prm, &bs, nullptr, nullptr, nullptr, nullptr, shs, radii, geom, GL, (const float *)scratch,
                               dverts, dL_dmeans2D, dL_dopacity, dL_dcolors_precomp, dL_dshs, nullptr, nullptr, dg,
                               accumulate, s);
    if (rc) return rc;
    return check_stage(prm, s, "bound preprocess backward");
}

int dmgs_backward(const dmgs_params *prm, const float *means3D, const float *scales, const float *rotations,
                  const float *cov3D_precomp, const float *shs, const int32_t *radii, const void *geom,
                  const void *binning, const void *image, int64_t R, const float *dL_dpix, float *dL_dmeans3D,
                  float *dL_dmeans2D, float *dL_dopacity, float *dL_dcolors_precomp, float *dL_dshs, float *dL_dscales,
                  float *dL_drotations, float *dL_dcov3D, void *scratch, void *stream)
{
    int rc = dmgs_blend_backward(prm, geom, binning, image, R, dL_dpix, scratch, stream);
    if (rc) return rc;
    return dmgs_preprocess_backward(prm, means3D, scales, rotations, cov3D_precomp, shs, radii, geom, scratch,
                                    dL_dmeans3D, dL_dmeans2D, dL_dopacity, dL_dcolors_precomp, dL_dshs, dL_dscales,
                                    dL_drotations, dL_dcov3D, 0, stream);
}

int dmgs_mark_visible(int32_t P, const float *means3D, const float *viewmatrix, const float *projmatrix,
                      uint8_t *visible, void *stream)
{
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !visible))) { set_error("bad arguments"); return -6; }
    return launch_mark_visible(P, means3D, viewmatrix, visible, (cudaStream_t)stream);
}

int dmgs_bind_forward(int64_t F, int32_t k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                      float thin_z, const float *g, int32_t adaptive, float *xyz, float *cov6, float *rot_t2w, void *stream)
{
    if (F < 0 || k <= 0 || (F > 0 && (!verts || !faces || !bc))) { set_error("bad arguments"); return -6; }
    return launch_bind_fwd(F, k, verts, faces, bc, rad_base, thin_z, g, adaptive, xyz, cov6, rot_t2w, (cudaStream_t)stream);
}

int dmgs_bind_backward(int64_t F, int32_t k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                       float thin_z, const float *g, int32_t adaptive, const float *dL_dxyz, const float *dL_dcov6,
                       const float *dL_drot, float *dverts, float *dg, void *stream)
{
    if (F < 0 || k <= 0 || (F > 0 && (!verts || !faces || !bc || !dverts))) { set_error("bad arguments"); return -6; }
    return launch_bind_bwd(F, k, verts, faces, bc, rad_base, thin_z, g, adaptive, dL_dxyz, dL_dcov6, dL_drot, dverts,
                           dg, (cudaStream_t)stream);
}

int dmgs_sh_grad_expand(int32_t P, int32_t sh_degree, int32_t sh_coeffs, int32_t sh_layout, int32_t n_views,
                        const float *campos_host, const float *means3D, const float *shs, const float *records,
                        int64_t view_stride, float *dL_dshs, float *dL_dmeans3D, int32_t accumulate, void *stream)
{
    if (P < 0 || sh_degree < 0 || sh_degree > 3 || sh_coeffs < (sh_degree + 1) * (sh_degree + 1)) { set_error("sh_grad_expand: bad shape"); return -14; }
    if (P > 0 && (!campos_host || !means3D || !shs || !records || !dL_dshs || !dL_dmeans3D)) { set_error("sh_grad_expand: NULL pointer"); return -6; }
    return launch_sh_grad_expand(P, sh_degree, sh_coeffs, sh_layout, n_views, campos_host, means3D, shs, records, view_stride,
                                 dL_dshs, dL_dmeans3D, accumulate, (cudaStream_t)stream);
}

int dmgs_stage3_forward(int64_t F, int32_t k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                        float thin_z, float *scales, float *quats, float *cov6, void *stream)
{
    if (F < 0 || k <= 0 || (F > 0 && (!rot_t2w || !rotation2d || !scaling2d))) { set_error("stage3: bad arguments"); return -6; }
    return launch_stage3_fwd(F, k, rot_t2w, rotation2d, scaling2d, thin_z, scales, quats, cov6, (cudaStream_t)stream);
}

int dmgs_stage3_backward(int64_t F, int32_t k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                         float thin_z, const float *dL_dscales, const float *dL_dquats, const float *dL_dcov6,
                         float *dL_drot, float *dL_drotation2d, float *dL_dscaling2d, void *stream)
{
    if (F < 0 || k <= 0 || (F > 0 && (!rot_t2w || !rotation2d || !scaling2d))) { set_error("stage3: bad arguments"); return -6; }
    return launch_stage3_bwd(F, k, rot_t2w, rotation2d, scaling2d, thin_z, dL_dscales, dL_dquats, dL_dcov6, dL_drot,
                             dL_drotation2d, dL_dscaling2d, (cudaStream_t)stream);
}

// ---- rows next to the path -------------------------------------------------
size_t dmgs_l1_ssim_scratch_bytes(int32_t planes, int32_t H, int32_t W)
{
    return loss_scratch_bytes(planes > 0 ? planes : 1, H > 0 ? H : 1, W > 0 ? W : 1);
}

int dmgs_l1_ssim_forward(int32_t planes, int32_t H, int32_t W, const float *window11_host, const float *img,
                         const float *gt, void *scratch, float *out_means, void *stream)
{
    if (planes <= 0 || H <= 0 || W <= 0 || planes > 65535) { set_error("l1_ssim: bad shape [%d,%d,%d]", planes, H, W); return -10; }
    if (!window11_host || !img || !gt || !scratch || !out_means) { set_error("l1_ssim: NULL pointer"); return -6; }
    return launch_l1_ssim_fwd(planes, H, W, window11_host, img, gt, scratch, out_means, (cudaStream_t)stream);
}

int dmgs_l1_ssim_backward(int32_t planes, int32_t H, int32_t W, const float *window11_host, const float *img,
                          const float *gt, const void *scratch, const float *upstream, float *grad, void *stream)
{
    if (planes <= 0 || H <= 0 || W <= 0 || planes > 65535) { set_error("l1_ssim: bad shape [%d,%d,%d]", planes, H, W); return -10; }
    if (!window11_host || !img || !gt || !scratch || !upstream || !grad) { set_error("l1_ssim: NULL pointer"); return -6; }
    return launch_l1_ssim_bwd(planes, H, W, window11_host, img, gt, scratch, upstream, grad, (cudaStream_t)stream);
}

size_t dmgs_frustum_scratch_bytes(int64_t N) { return frustum_scratch_bytes(N > 0 ? N : 1); }

int dmgs_in_frustum(int64_t N, const float *proj16_host, float cube_len, int32_t has_cube, int32_t piece_id,
                    int32_t n_piece, const float *pts, const int64_t *faces, uint8_t *mask, int64_t *faces_out,
                    int32_t *index_out, int32_t *count_out, void *scratch, void *stream)
{
    if (N < 0 || N >= ((int64_t)1 << 31)) { set_error("in_frustum: N out of range"); return -10; }
    if (!proj16_host || (N > 0 && (!pts || !mask))) { set_error("in_frustum: NULL pointer"); return -6; }
    if (count_out && !scratch) { set_error("in_frustum: compaction needs the scratch buffer"); return -6; }
    if (faces_out && !faces) { set_error("in_frustum: faces_out needs faces"); return -6; }
    return launch_frustum(N, proj16_host, cube_len, has_cube, piece_id, n_piece, pts, faces, mask, faces_out, index_out,
                          count_out, scratch, (cudaStream_t)stream);
}

int dmgs_adam_step(int32_t nseg, const dmgs_adam_segment *segments_host, double beta1, double beta2, double eps,
                   int64_t step, float grad_scale, int32_t zero_grad, void *stream)
{
    if (!segments_host) { set_error("adam: NULL segments"); return -6; }
    return launch_adam(nseg, segments_host, beta1, beta2, eps, step, grad_scale, zero_grad, (cudaStream_t)stream);
}

int64_t dmgs_texture_grid_params(void) { return texture_grid_params(); }

int dmgs_texture_cast_params(int64_t n_params, const float *params, void *params_half, void *stream)
{
    if (n_params != texture_grid_params()) { set_error("texture: the hash grid holds %lld parameters, got %lld", (long long)texture_grid_params(), (long long)n_params); return -14; }
    if (!params || !params_half) { set_error("texture: NULL required pointer"); return -6; }
    return launch_texture_cast(n_params, params, params_half, (cudaStream_t)stream);
}

int dmgs_texture_forward(int64_t N, int32_t channels, const float *aabb6_host, const float *xyz, const void *grid_half,
                         const float *W0, const float *W1, const float *W2, float *out, void *enc_out, void *stream)
{
    if (N > 0 && (!aabb6_host || !xyz || !grid_half || !W0 || !W1 || !W2 || !out)) { set_error("texture: NULL required pointer"); return -6; }
    if (!aabb6_host) { set_error("texture: NULL AABB"); return -6; }
    return launch_texture_fwd(N, channels, aabb6_host, xyz, grid_half, W0, W1, W2, out, enc_out, (cudaStream_t)stream);
}

int dmgs_texture_backward(int64_t N, int32_t channels, const float *aabb6_host, const float *xyz, const void *grid_half,
                          const void *enc, const float *W0, const float *W1, const float *W2, const float *dL_dout,
                          float grid_grad_scale, float *d_grid, float *dW0, float *dW1, float *dW2, float *d_xyz,
                          void *scratch, void *stream)
{
    if (!aabb6_host) { set_error("texture: NULL AABB"); return -6; }
    if (N > 0 && (!xyz || !grid_half || !enc || !W0 || !W1 || !W2 || !dL_dout || !dW0 || !dW1 || !dW2 || !scratch)) { set_error("texture: NULL required pointer"); return -6; }
    return launch_texture_bwd(N, channels, aabb6_host, xyz, grid_half, enc, W0, W1, W2, dL_dout, grid_grad_scale, d_grid,
                              dW0, dW1, dW2, d_xyz, scratch, (cudaStream_t)stream);
}

size_t dmgs_texture_backward_scratch_bytes(int64_t N) { return texture_bwd_scratch_bytes(N); }

int dmgs_adam_exchange_shard(int64_t n, int32_t world, int32_t rank, int64_t *begin4, int64_t *end4)
{
    if (!begin4 || !end4 || n < 0 || world < 1 || rank < 0 || rank >= world) { set_error("adam_exchange_shard: bad arguments"); return -12; }
    adam_exchange_shard(n, world, rank, begin4, end4);
    return 0;
}

int dmgs_adam_exchange_peer(int32_t world, int32_t rank, int32_t nseg, const dmgs_adam_xsegment *segments_host,
                            const void *const *grad_peer_ptrs_host, void *grad_multicast_ptr,
                            const void *const *param_peer_ptrs_host, void *param_multicast_ptr, double beta1, double beta2,
                            double eps, int64_t step, float grad_scale, void *stream)
{
    if (!segments_host || !grad_peer_ptrs_host || !param_peer_ptrs_host) { set_error("adam_exchange: NULL required pointer"); return -6; }
    return launch_adam_exchange(world, rank, nseg, segments_host, grad_peer_ptrs_host, grad_multicast_ptr, param_peer_ptrs_host,
                                param_multicast_ptr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

int dmgs_allreduce_peer(int64_t n, int32_t world, int32_t rank, const void *const *peer_ptrs_host, void *multicast_ptr,
                        float scale, void *stream)
{
    if (!peer_ptrs_host && !multicast_ptr) { set_error("allreduce_peer: no peer pointers"); return -6; }
    return launch_allreduce_peer(n, world, rank, peer_ptrs_host, multicast_ptr, scale, (cudaStream_t)stream);
}

int dmgs_sorted_keys(const void *geom, const void *binning, int32_t P, int64_t R, int32_t W, int32_t H,
                     uint64_t *keys_out, void *stream)
{
    const GeomLayout GL = geom_layout(P);
    const BinLayout BL = bin_layout(P, R, W, H);
    const int T = ((W + DMGS_TILE - 1) / DMGS_TILE) * ((H + DMGS_TILE - 1) / DMGS_TILE);
    if (place_plan(P, T).ok && R > 0) {  // the placement path never materialises the sorted tile ids
        int rc = launch_fill_tiles(T, at<uint2>(binning, BL.ranges), at<uint32_t>(const_cast<void *>(binning), BL.tiles),
                                   (cudaStream_t)stream);
        if (rc) return rc;
    }
    return launch_sorted_keys(R, at<uint32_t>(binning, BL.tiles), at<uint32_t>(binning, BL.gidx),
                              at<float>(geom, GL.depths), keys_out, (cudaStream_t)stream);
}

int dmgs_exp_array(const float *x, float *y, int64_t n, void *stream)
{
    return launch_exp_array(x, y, n, (cudaStream_t)stream);
}

}  // extern "C"
