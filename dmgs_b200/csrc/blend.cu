// blend.cu -- per-tile front-to-back alpha blending (forward) and its back-to-front adjoint.
// Replaces upstream renderCUDA forward/backward.
//
// One CTA per 16x16 tile, 8 warps; warp w owns the 8x4 pixel sub-rectangle
// (w & 1, w >> 1) so that warp-level culling has a compact footprint.  Entries of the tile's
// depth-ordered list are staged through shared memory in rounds of 256 (one gather per thread:
// a 32-byte record + a 16-byte colour per Gaussian, both sector-aligned in L2).
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int BLK = 256;

struct BlendArgs {
    int W, H, gx, gy;
    float bg[3];
};

__device__ __forceinline__ void pixel_of_thread(int &px, int &py)
{
    // warp w -> 8x4 sub-rectangle; lane -> pixel inside it
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = blockIdx.x * DMGS_TILE + (w & 1) * 8 + (lane & 7);
    py = blockIdx.y * DMGS_TILE + (w >> 1) * 4 + (lane >> 3);
}

__global__ void __launch_bounds__(BLK)
blend_fwd_kernel(const __grid_constant__ BlendArgs a, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 float *__restrict__ out_color, float *__restrict__ final_T, uint32_t *__restrict__ n_contrib)
{
    __shared__ float4 s_ra[BLK];   // x, y, conA, conB
    __shared__ float2 s_rb[BLK];   // conC, opacity
    __shared__ float4 s_rgb[BLK];

    int px, py;
    pixel_of_thread(px, py);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    int todo = (int)(rng.y - rng.x);
    const int rounds = (todo + BLK - 1) / BLK;

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f;
    uint32_t contributor = 0, last = 0;

    for (int r = 0; r < rounds; ++r, todo -= BLK) {
        if (__syncthreads_count(done) == BLK) break;
        const int idx = r * BLK + threadIdx.x;
        if (idx < (int)(rng.y - rng.x)) {
            const uint32_t g = gidx[rng.x + idx];
            const float4 ra = rec[2 * (size_t)g], rb = rec[2 * (size_t)g + 1];
            s_ra[threadIdx.x] = ra;
            s_rb[threadIdx.x] = make_float2(rb.x, rb.y);
            s_rgb[threadIdx.x] = rgb4[g];
        }
        __syncthreads();
        const int nb = min(BLK, todo);
        for (int j = 0; !done && j < nb; ++j) {
            ++contributor;
            const float4 ra = s_ra[j];
            const float2 rb = s_rb[j];
            const float dx = ra.x - pxf, dy = ra.y - pyf;
            const float q = fma_(rb.x * dy, dy, (ra.z * dx) * dx);
            const float power = fma_(-(ra.w * dx), dy, -0.5f * q);
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, rb.y * dmgs_exp(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.0f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 c = s_rgb[j];
            C0 = fma_(c.x * alpha, T, C0);
            C1 = fma_(c.y * alpha, T, C1);
            C2 = fma_(c.z * alpha, T, C2);
            T = test_T;
            last = contributor;
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

// ------------------------------------------------------------------------------ backward
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(BLK)
blend_bwd_kernel(const __grid_constant__ BlendArgs a, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                 const float *__restrict__ dL_dpix, float *__restrict__ grad_blend)
{
    __shared__ float4 s_ra[BLK];
    __shared__ float2 s_rb[BLK];
    __shared__ float4 s_rgb[BLK];
    __shared__ uint32_t s_id[BLK];
    __shared__ int s_max;

    int px, py;
    pixel_of_thread(px, py);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    const size_t pix = (size_t)py * a.W + px, HW = (size_t)a.H * a.W;
    const int lane = threadIdx.x & 31;

    const float T_final = inside ? final_T[pix] : 0.0f;
    const int last = inside ? (int)n_contrib[pix] : 0;
    float dp0 = 0, dp1 = 0, dp2 = 0;
    if (inside) { dp0 = dL_dpix[pix]; dp1 = dL_dpix[HW + pix]; dp2 = dL_dpix[2 * HW + pix]; }
    const float bg_dot = dot3(a.bg[0], dp0, a.bg[1], dp1, a.bg[2], dp2);
    const float ddelx_dx = 0.5f * (float)a.W, ddely_dy = 0.5f * (float)a.H;

    // only the first max(n_contrib) entries of the list matter
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int count = s_max;
    if (count == 0) return;

    float T = T_final, acc0 = 0, acc1 = 0, acc2 = 0, lc0 = 0, lc1 = 0, lc2 = 0, last_alpha = 0;
    const int rounds = (count + BLK - 1) / BLK;
    for (int r = rounds - 1; r >= 0; --r) {
        __syncthreads();
        const int idx = r * BLK + threadIdx.x;
        if (idx < count) {
            const uint32_t g = gidx[rng.x + idx];
            const float4 ra = rec[2 * (size_t)g], rb = rec[2 * (size_t)g + 1];
            s_id[threadIdx.x] = g;
            s_ra[threadIdx.x] = ra;
            s_rb[threadIdx.x] = make_float2(rb.x, rb.y);
            s_rgb[threadIdx.x] = rgb4[g];
        }
        __syncthreads();
        const int nb = min(BLK, count - r * BLK);
        for (int j = nb - 1; j >= 0; --j) {
            const int pos = r * BLK + j;  // 0-based list position
            float g0 = 0, g1 = 0, g2 = 0, g3 = 0, g4 = 0, g5 = 0, g6 = 0, g7 = 0, g8 = 0;
            bool hit = false;
            if (pos < last) {
                const float4 ra = s_ra[j];
                const float2 rb = s_rb[j];
                const float dx = ra.x - pxf, dy = ra.y - pyf;
                const float q = fma_(rb.x * dy, dy, (ra.z * dx) * dx);
                const float power = fma_(-(ra.w * dx), dy, -0.5f * q);
                if (power <= 0.0f) {
                    const float G = dmgs_exp(power);
                    const float alpha = fminf(0.99f, rb.y * G);
                    if (alpha >= 1.0f / 255.0f) {
                        hit = true;
                        T = T / (1.0f - alpha);
                        const float w = alpha * T;
                        const float4 c = s_rgb[j];
                        float dL_dalpha;
                        acc0 = fma_(last_alpha, lc0, (1.0f - last_alpha) * acc0);
                        acc1 = fma_(last_alpha, lc1, (1.0f - last_alpha) * acc1);
                        acc2 = fma_(last_alpha, lc2, (1.0f - last_alpha) * acc2);
                        lc0 = c.x; lc1 = c.y; lc2 = c.z;
                        dL_dalpha = (c.x - acc0) * dp0;
                        dL_dalpha = fma_(c.y - acc1, dp1, dL_dalpha);
                        dL_dalpha = fma_(c.z - acc2, dp2, dL_dalpha);
                        g6 = w * dp0; g7 = w * dp1; g8 = w * dp2;
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha = fma_(-T_final / (1.0f - alpha), bg_dot, dL_dalpha);
                        const float dL_dG = rb.y * dL_dalpha;
                        const float gdx = G * dx, gdy = G * dy;
                        const float dG_ddelx = fma_(-gdy, ra.w, -gdx * ra.z);
                        const float dG_ddely = fma_(-gdx, ra.w, -gdy * rb.x);
                        g0 = (dL_dG * dG_ddelx) * ddelx_dx;
                        g1 = (dL_dG * dG_ddely) * ddely_dy;
                        g2 = (-0.5f * gdx) * dx * dL_dG;
                        g3 = (-0.5f * gdx) * dy * dL_dG;
                        g4 = (-0.5f * gdy) * dy * dL_dG;
                        g5 = G * dL_dalpha;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, hit)) continue;
            g0 = warp_sum(g0); g1 = warp_sum(g1); g2 = warp_sum(g2); g3 = warp_sum(g3); g4 = warp_sum(g4);
            g5 = warp_sum(g5); g6 = warp_sum(g6); g7 = warp_sum(g7); g8 = warp_sum(g8);
            if (lane == 0) {
                float *dst = grad_blend + 12 * (size_t)s_id[j];
                atomicAdd(dst + 0, g0); atomicAdd(dst + 1, g1); atomicAdd(dst + 2, g2); atomicAdd(dst + 3, g3);
                atomicAdd(dst + 4, g4); atomicAdd(dst + 5, g5); atomicAdd(dst + 6, g6); atomicAdd(dst + 7, g7);
                atomicAdd(dst + 8, g8);
            }
        }
    }
}

int launch_blend_bwd(const dmgs_params *prm, const void *geom, const GeomLayout &GL, const void *binning,
                     const BinLayout &BL, const void *image, const ImgLayout &IL, const float *dL_dpix,
                     float *grad_blend, cudaStream_t s)
{
    BlendArgs a;
    a.W = prm->image_width; a.H = prm->image_height;
    a.gx = (a.W + DMGS_TILE - 1) / DMGS_TILE; a.gy = (a.H + DMGS_TILE - 1) / DMGS_TILE;
    for (int i = 0; i < 3; ++i) a.bg[i] = prm->bg[i];
    if (a.W <= 0 || a.H <= 0) return 0;
    blend_bwd_kernel<<<dim3(a.gx, a.gy), BLK, 0, s>>>(a, at<uint2>(binning, BL.ranges), at<uint32_t>(binning, BL.gidx),
                                                      at<float4>(geom, GL.rec), at<float4>(geom, GL.rgb),
                                                      at<float>(image, IL.final_T), at<uint32_t>(image, IL.n_contrib),
                                                      dL_dpix, grad_blend);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
