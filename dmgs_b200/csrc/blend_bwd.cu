// blend_bwd.cu -- back-to-front adjoint of the per-tile alpha blend; see blend.cu for
// the tile / warp-rectangle layout and the culling scheme, which are shared.
//
// The backward is held to 1e-4 relative (BASELINE.json), not to bit parity, but everything that
// decides WHICH entries contribute (power, exp, alpha and the two tests) repeats the forward's
// arithmetic operation for operation (explicit __fmul_rn / __fmaf_rn), so the contributor set is
// bit-identical to the forward's.  (Compiling this unit with -fmad=true was measured: no gain, the
// gradient arithmetic is already written as explicit fused multiply-adds.)
#include "blend_common.cuh"

namespace dmgs {

// ------------------------------------------------------------------------------ backward
// Transposed butterfly: N per-lane values -> N totals over the warp.  At every step a lane keeps
// one half of its values and hands the other half to its partner, so the payload halves with the
// distance: 5+3+2+1+1 = 12 shuffles for N = 9.  The total of slot `tr_slot9(lane)` ends in v[0].
template <int N, int OFF>
__device__ __forceinline__ void tr_reduce(float *v, int lane)
{
    if constexpr (N == 1) {
#pragma unroll
        for (int o = OFF; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
    } else {
        constexpr int LO = (N + 1) / 2;
        const bool up = lane & OFF;
#pragma unroll
        for (int i = 0; i < LO; ++i) {
            const float hi = (LO + i < N) ? v[LO + i] : 0.0f;
            const float send = up ? v[i] : hi;
            const float keepv = up ? hi : v[i];
            v[i] = keepv + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        tr_reduce<LO, OFF / 2>(v, lane);
    }
}
// slot whose total lands in this lane's v[0] after tr_reduce<9,16> (-1: a padding slot)
__device__ __forceinline__ int tr_slot9(int lane)
{
    int n = 9, base = 0, cnt = 9;
#pragma unroll
    for (int off = 16; off >= 2; off >>= 1) {
        const int lo = (n + 1) / 2;
        if (lane & off) { base += lo; cnt -= lo; } else { cnt = min(cnt, lo); }
        n = lo;
    }
    return cnt >= 1 ? base : -1;
}

__global__ void __launch_bounds__(BLK, BWD_MIN_BLOCKS)
blend_bwd_kernel(const __grid_constant__ BlendArgs a, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                 const float *__restrict__ dL_dpix, float *__restrict__ grad_blend)
{
    __shared__ float4 s_ra[BLK];
    __shared__ float4 s_rb[BLK];
    __shared__ float4 s_rgb[BLK];
    __shared__ float4 s_acc[BLK * 3];  // per staged entry: the 9 (+3 pad) gradient sums of this tile
    __shared__ uint32_t s_id[BLK];
    __shared__ __align__(16) unsigned char s_cw[(BLK / 32) * CW_WARP_BYTES];  // per-warp compacted survivors
    __shared__ int s_max;

    const int lane = threadIdx.x & 31;
    int px0, py0;
    warp_rect(px0, py0);
    const int px = px0 + (lane & 7), py = py0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const f32x2 npx = pk2(-pxf, -pxf), npy = pk2(-pyf, -pyf);
    const float rx0 = (float)px0, rx1 = (float)(px0 + 7), ry0 = (float)py0, ry1 = (float)(py0 + 3);
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    const size_t pix = (size_t)py * a.W + px, HW = (size_t)a.H * a.W;

    const float T_final = inside ? final_T[pix] : 0.0f;
    const int last = inside ? (int)n_contrib[pix] : 0;
    float dp0 = 0, dp1 = 0, dp2 = 0;
    if (inside) { dp0 = dL_dpix[pix]; dp1 = dL_dpix[HW + pix]; dp2 = dL_dpix[2 * HW + pix]; }
    const float bgT = -T_final * dot3(a.bg[0], dp0, a.bg[1], dp1, a.bg[2], dp2);
    const float ddelx_dx = 0.5f * (float)a.W, ddely_dy = 0.5f * (float)a.H;
    const int slot = tr_slot9(lane);
    const bool owner = slot >= 0 && !(lane & 1);
    const uint32_t a_ra = smem_addr(s_ra), a_rb = smem_addr(s_rb), a_rgb = smem_addr(s_rgb);
    const uint32_t a_acc = smem_addr(s_acc) + 4u * (uint32_t)(slot < 0 ? 0 : slot);
    const uint32_t a_cw = smem_addr(s_cw) + (uint32_t)(threadIdx.x >> 5) * CW_WARP_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // only the first max(n_contrib) entries of the list matter: per tile for staging, per warp for work
    if (threadIdx.x == 0) s_max = 0;
    s_acc[threadIdx.x] = make_float4(0, 0, 0, 0);
    s_acc[BLK + threadIdx.x] = make_float4(0, 0, 0, 0);
    s_acc[2 * BLK + threadIdx.x] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int count = s_max;
    if (count == 0) return;

    float T = T_final, behind = 0, last_cd = 0, last_alpha = 0;
    const int rounds = (count + BLK - 1) / BLK;
    for (int r = rounds - 1; r >= 0; --r) {
        const int idx = r * BLK + threadIdx.x;
        if (idx < count) {
            const uint32_t g = gidx[rng.x + idx];
            s_id[threadIdx.x] = g;
            s_ra[threadIdx.x] = rec[2 * (size_t)g];
            s_rb[threadIdx.x] = rec[2 * (size_t)g + 1];
            s_rgb[threadIdx.x] = rgb4[g];
        }
        __syncthreads();
        if (r * BLK < wmax) {  // else nothing in this round is a contributor for this warp
            const int nb = min(BLK, min(count, wmax) - r * BLK);
            const int lim = last - r * BLK;  // staged entries j < lim lie before this pixel's last contributor
            for (int s0 = ((nb - 1) / 32) * 32; s0 >= 0; s0 -= 32) {
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
                if (keep) {  // compact the survivors in list order (see blend.cu)
                    const uint32_t w = a_cw + 4u * (uint32_t)__popc(m & lt_mask);
                    sts32(w, ra.x); sts32(w + CW_STRIDE, ra.y); sts32(w + 2 * CW_STRIDE, ra.z); sts32(w + 3 * CW_STRIDE, -ra.w);
                    sts32(w + 4 * CW_STRIDE, rb.x); sts32(w + 5 * CW_STRIDE, rb.y); sts32u(w + 6 * CW_STRIDE, (uint32_t)e);
                }
                if (lane == 0 && (n & 1)) {  // sentinel: opacity 0 -> alpha 0 -> never a contributor
                    const uint32_t w = a_cw + 4u * (uint32_t)n;
                    sts32(w, 0.0f); sts32(w + CW_STRIDE, 0.0f); sts32(w + 2 * CW_STRIDE, 0.0f); sts32(w + 3 * CW_STRIDE, 0.0f);
                    sts32(w + 4 * CW_STRIDE, 0.0f); sts32(w + 5 * CW_STRIDE, 0.0f); sts32u(w + 6 * CW_STRIDE, 0u);
                }
                __syncwarp();
                // back to front over the survivors, two per iteration: the alpha arithmetic (the forward's, so the
                // contributor set is identical) runs packed for both, the recurrences and reductions one by one
                for (int t = (n - 1) & ~1; t >= 0; t -= 2) {
                    const uint32_t cw = a_cw + 4u * (uint32_t)t;
                    f32x2 power2, alpha2, dx2, dy2, G2;
                    alpha_pair(cw, npx, npy, power2, alpha2, dx2, dy2, G2);
                    float pw[2], al[2], dxs[2], dys[2], Gs[2];
                    upk2(power2, pw[0], pw[1]);
                    upk2(alpha2, al[0], al[1]);
                    upk2(dx2, dxs[0], dxs[1]);
                    upk2(dy2, dys[0], dys[1]);
                    upk2(G2, Gs[0], Gs[1]);
                    const f32x2 jj = lds64(cw + 6 * CW_STRIDE);
                    const int js[2] = {(int)(uint32_t)(jj & 0xffffffffull), (int)(uint32_t)(jj >> 32)};
#pragma unroll
                    for (int h = 1; h >= 0; --h) {
                        const int j = js[h];
                        const float alpha = al[h], dx = dxs[h], dy = dys[h];
                        // per-pixel work stops at cg = G * dL/dalpha and w = alpha * T; lanes that do not
                        // contribute keep both at zero, so the products below need no other masking
                        float cg = 0.0f, w = 0.0f;
                        const bool hit = j < lim && pw[h] <= 0.0f && alpha >= 1.0f / 255.0f;
                        if (hit) {
                            // one refined reciprocal replaces two IEEE divisions by (1 - alpha) (no FCHK /
                            // slow-path branches; operands are in [0.01, 1] so no special cases exist)
                            const float oma = 1.0f - alpha;
                            const float inv = rcp_nr(oma);
                            const float t0 = T * inv;  // T / (1 - alpha), residual-corrected: the error must
                            T = fma_(fma_(-t0, oma, T), inv, t0);  // not accumulate along the list
                            w = alpha * T;
                            const float4 c = lds128(a_rgb + 16u * j);
                            // colour behind this entry enters only through its dot product with dL/dpixel
                            const float cd = fma_(c.z, dp2, fma_(c.y, dp1, c.x * dp0));
                            behind = fma_(last_alpha, last_cd, (1.0f - last_alpha) * behind);
                            last_cd = cd;
                            last_alpha = alpha;
                            cg = Gs[h] * fma_(bgT, inv, (cd - behind) * T);
                        }
                        if (!__any_sync(0xffffffffu, hit)) continue;
                        // moments of cg about the Gaussian's centre (the flush below turns the tile's sums into
                        // dL/dmean2D and dL/dconic) and the colour gradient
                        const float cgx = cg * dx, cgy = cg * dy;
                        float v[9] = {cgx, cgy, cgx * dx, cgx * dy, cgy * dy, cg, w * dp0, w * dp1, w * dp2};
                        tr_reduce<9, 16>(v, lane);
                        if (owner) reds_add(a_acc + 48u * j, v[0]);
                    }
                }
                __syncwarp();  // the buffer is rewritten by the next group
            }
        }
        __syncthreads();
        // flush the round's tile-level sums: three 16-byte vector reductions per touched Gaussian
        if (idx < count) {
            float4 g0 = s_acc[3 * threadIdx.x], g1 = s_acc[3 * threadIdx.x + 1];
            const float4 g2 = s_acc[3 * threadIdx.x + 2];
            const bool nz = g0.x != 0.0f || g0.y != 0.0f || g0.z != 0.0f || g0.w != 0.0f || g1.x != 0.0f || g1.y != 0.0f ||
                            g1.z != 0.0f || g1.w != 0.0f || g2.x != 0.0f;
            if (nz) {
                // moments -> gradients: dG/ddelta = -G (A dx + B dy, B dx + C dy), dG/dconic = -0.5 G (dx^2, dx dy, dy^2)
                const float4 ra = s_ra[threadIdx.x], rb = s_rb[threadIdx.x];
                const float opx = -rb.y * ddelx_dx, opy = -rb.y * ddely_dy, oph = -0.5f * rb.y;
                const float mx = g0.x, my = g0.y;
                g0.x = opx * fma_(ra.w, my, ra.z * mx);
                g0.y = opy * fma_(rb.x, my, ra.w * mx);
                g0.z *= oph; g0.w *= oph; g1.x *= oph;
                float *dst = grad_blend + 12 * (size_t)s_id[threadIdx.x];
                red_global_v4(dst, g0);
                red_global_v4(dst + 4, g1);
                red_global_v4(dst + 8, g2);
                s_acc[3 * threadIdx.x] = make_float4(0, 0, 0, 0);
                s_acc[3 * threadIdx.x + 1] = make_float4(0, 0, 0, 0);
                s_acc[3 * threadIdx.x + 2] = make_float4(0, 0, 0, 0);
            }
        }
        __syncthreads();
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
