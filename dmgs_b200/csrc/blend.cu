// blend.cu -- per-tile front-to-back alpha blending (forward) and its back-to-front adjoint.
// Replaces upstream renderCUDA forward/backward.
//
// One CTA per 16x16 tile, 8 warps; warp w owns the 8x4 pixel sub-rectangle (w & 1, w >> 1).
// The tile's depth-ordered entries are staged through shared memory in rounds of 256 (one
// 32-byte record + one 16-byte colour per Gaussian, both sector-aligned gathers that hit L2).
//
// Warp-level culling (the B200-first part): for every group of 32 staged entries the warp first
// runs ONE pass with lane <-> entry in which each lane bounds the entry's exponent over the
// warp's whole 8x4 rectangle (exact minimum of the conic's quadratic form over the rectangle);
// a ballot gives the entries that can reach alpha >= 1/255 somewhere in the rectangle, and only
// those are evaluated per pixel (lane <-> pixel).  The bound is conservative by construction
// (margins cover fp32 rounding), skipped entries are exactly the ones the per-pixel tests would
// skip for all 32 pixels, so images, final T and contributor counts stay bit-identical.
//
// Backward: per contributing (warp, entry) the nine partial gradients are reduced across the 32
// pixels with a transposed butterfly (12 shuffles instead of 45) that leaves each total in a
// different lane, so ONE predicated atomic instruction adds all nine to the Gaussian's record.
#include "blend_common.cuh"

namespace dmgs {

__global__ void __launch_bounds__(BLK, 4)
blend_fwd_kernel(const __grid_constant__ BlendArgs a, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 float *__restrict__ out_color, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib)
{
    __shared__ float4 s_ra[BLK];   // x, y, conA, conB
    __shared__ float4 s_rb[BLK];   // conC, opacity, cut, -
    __shared__ float4 s_rgb[BLK];
    __shared__ __align__(16) unsigned char s_cw[(BLK / 32) * CW_WARP_BYTES];  // per-warp compacted survivors

    const int lane = threadIdx.x & 31;
    int px0, py0;
    warp_rect(px0, py0);
    const int px = px0 + (lane & 7), py = py0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const f32x2 npx = pk2(-pxf, -pxf), npy = pk2(-pyf, -pyf);
    const float rx0 = (float)px0, rx1 = (float)(px0 + 7), ry0 = (float)py0, ry1 = (float)(py0 + 3);
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    const int total = (int)(rng.y - rng.x);
    const int rounds = (total + BLK - 1) / BLK;

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t last = 0;
    const uint32_t a_ra = smem_addr(s_ra), a_rb = smem_addr(s_rb), a_rgb = smem_addr(s_rgb);
    const uint32_t a_cw = smem_addr(s_cw) + (uint32_t)(threadIdx.x >> 5) * CW_WARP_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int r = 0; r < rounds; ++r) {
        if (__syncthreads_count(done) == BLK) break;
        const int idx = r * BLK + threadIdx.x;
        if (idx < total) {
            const uint32_t g = gidx[rng.x + idx];
            s_ra[threadIdx.x] = rec[2 * (size_t)g];
            s_rb[threadIdx.x] = rec[2 * (size_t)g + 1];
            s_rgb[threadIdx.x] = rgb4[g];
        }
        __syncthreads();
        const int nb = min(BLK, total - r * BLK);
        if (__all_sync(0xffffffffu, done)) continue;  // this warp's pixels are finished; keep staging
        for (int s0 = 0; s0 < nb; s0 += 32) {
            // lane <-> entry: which of these 32 entries can touch the warp's rectangle?  Survivors are compacted,
            // in list order, into the warp's structure-of-arrays buffer.
            const int e = s0 + lane;
            bool keep = false;
            float4 ra, rb;
            if (e < nb) {
                ra = lds128(a_ra + 16u * e);
                rb = lds128(a_rb + 16u * e);
                keep = !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, rx0, rx1, ry0, ry1);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (!m) continue;
            const int n = __popc(m);
            if (keep) {
                const uint32_t w = a_cw + 4u * (uint32_t)__popc(m & lt_mask);
                sts32(w, ra.x); sts32(w + CW_STRIDE, ra.y); sts32(w + 2 * CW_STRIDE, ra.z); sts32(w + 3 * CW_STRIDE, -ra.w);
                sts32(w + 4 * CW_STRIDE, rb.x); sts32(w + 5 * CW_STRIDE, rb.y); sts32u(w + 6 * CW_STRIDE, (uint32_t)e);
            }
            if (lane == 0 && (n & 1)) {  // sentinel pads an odd count: opacity 0 -> alpha 0 -> never a contributor
                const uint32_t w = a_cw + 4u * (uint32_t)n;
                sts32(w, 0.0f); sts32(w + CW_STRIDE, 0.0f); sts32(w + 2 * CW_STRIDE, 0.0f); sts32(w + 3 * CW_STRIDE, 0.0f);
                sts32(w + 4 * CW_STRIDE, 0.0f); sts32(w + 5 * CW_STRIDE, 0.0f); sts32u(w + 6 * CW_STRIDE, 0u);
            }
            __syncwarp();
            // lane <-> pixel over the survivors, two per iteration (packed FP32), in list order
            for (int t = 0; t < n; t += 2) {
                const uint32_t cw = a_cw + 4u * (uint32_t)t;
                f32x2 power, alpha, dx, dy, G;
                alpha_pair(cw, npx, npy, power, alpha, dx, dy, G);
                float p0, p1, a0, a1;
                upk2(power, p0, p1);
                upk2(alpha, a0, a1);
                const bool h0 = p0 <= 0.0f && a0 >= 1.0f / 255.0f, h1 = p1 <= 0.0f && a1 >= 1.0f / 255.0f;
                if (!done && h0) {
                    const float test_T = T * (1.0f - a0);
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const uint32_t j = lds32(cw + 6 * CW_STRIDE);
                        const float4 c = lds128(a_rgb + 16u * j);
                        C0 = fma_(c.x * a0, T, C0);
                        C1 = fma_(c.y * a0, T, C1);
                        C2 = fma_(c.z * a0, T, C2);
                        T = test_T;
                        last = (uint32_t)(r * BLK) + j + 1u;
                    }
                }
                if (!done && h1) {
                    const float test_T = T * (1.0f - a1);
                    if (test_T < 0.0001f) {
                        done = true;
                    } else {
                        const uint32_t j = lds32(cw + 6 * CW_STRIDE + 4u);
                        const float4 c = lds128(a_rgb + 16u * j);
                        C0 = fma_(c.x * a1, T, C0);
                        C1 = fma_(c.y * a1, T, C1);
                        C2 = fma_(c.z * a1, T, C2);
                        T = test_T;
                        last = (uint32_t)(r * BLK) + j + 1u;
                    }
                }
            }
            __syncwarp();  // the buffer is rewritten by the next group
            if (__all_sync(0xffffffffu, done)) break;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * a.W + px, HW = (size_t)a.H * a.W;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = fma_(T, a.bg[0], C0);
        out_color[HW + pix] = fma_(T, a.bg[1], C1);
        out_color[2 * HW + pix] = fma_(T, a.bg[2], C2);
    }
}

int launch_blend_fwd(const dmgs_params *prm, const void *geom, const GeomLayout &GL, const void *binning,
                     const BinLayout &BL, float *out_color, void *image, const ImgLayout &IL, cudaStream_t s)
{
    BlendArgs a;
    a.W = prm->image_width; a.H = prm->image_height;
    a.gx = (a.W + DMGS_TILE - 1) / DMGS_TILE; a.gy = (a.H + DMGS_TILE - 1) / DMGS_TILE;
    for (int i = 0; i < 3; ++i) a.bg[i] = prm->bg[i];
    if (a.W <= 0 || a.H <= 0) return 0;
    blend_fwd_kernel<<<dim3(a.gx, a.gy), BLK, 0, s>>>(a, at<uint2>(binning, BL.ranges), at<uint32_t>(binning, BL.gidx),
                                                      at<float4>(geom, GL.rec), at<float4>(geom, GL.rgb), out_color,
                                                      at<float>(image, IL.final_T), at<uint32_t>(image, IL.n_contrib));
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
