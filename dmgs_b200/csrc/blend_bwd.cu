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
    __shared__ float4 s_ra[ROUND];
    __shared__ float4 s_rb[ROUND];
    __shared__ float4 s_rgb[ROUND];
    __shared__ float4 s_acc[ROUND * 3];  // per staged entry: the 9 (+3 pad) gradient sums of this tile
    __shared__ uint32_t s_id[ROUND];
    __shared__ __align__(16) unsigned char s_cw[(BLK / 32) * CW_WARP_BYTES];  // per-warp compacted survivors
    __shared__ int s_max;

    const int lane = threadIdx.x & 31;
    int px0, py0;
    warp_rect(px0, py0);
    const int px = px0 + (lane & 7), pya = py0 + (lane >> 3), pyb = pya + 4;
    const bool in_a = px < a.W && pya < a.H, in_b = px < a.W && pyb < a.H;
    const float pxf = (float)px;
    const f32x2 npx = pk2(-pxf, -pxf), npy = pk2(-(float)pya, -(float)pyb);
    const float rx0 = (float)px0, rx1 = (float)(px0 + 7), ry0 = (float)py0, ry1 = (float)(py0 + 7);
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    const size_t pix_a = (size_t)pya * a.W + px, pix_b = (size_t)pyb * a.W + px, HW = (size_t)a.H * a.W;

    // per pixel: T (transmittance in front of the entry being visited), `behind` = (colour blended behind that
    // entry) . dL/dpixel, normalised by the transmittance behind it, and the background's share bgT
    float Ta = in_a ? final_T[pix_a] : 0.0f, Tb = in_b ? final_T[pix_b] : 0.0f;
    const int last_a = in_a ? (int)n_contrib[pix_a] : 0, last_b = in_b ? (int)n_contrib[pix_b] : 0;
    float da0 = 0, da1 = 0, da2 = 0, db0 = 0, db1 = 0, db2 = 0;
    if (in_a) { da0 = dL_dpix[pix_a]; da1 = dL_dpix[HW + pix_a]; da2 = dL_dpix[2 * HW + pix_a]; }
    if (in_b) { db0 = dL_dpix[pix_b]; db1 = dL_dpix[HW + pix_b]; db2 = dL_dpix[2 * HW + pix_b]; }
    const float bgTa = -Ta * dot3(a.bg[0], da0, a.bg[1], da1, a.bg[2], da2);
    const float bgTb = -Tb * dot3(a.bg[0], db0, a.bg[1], db1, a.bg[2], db2);
    float bha = 0.0f, bhb = 0.0f;
    const f32x2 dp0 = pk2(da0, db0), dp1 = pk2(da1, db1), dp2 = pk2(da2, db2);
    const float ddelx_dx = 0.5f * (float)a.W, ddely_dy = 0.5f * (float)a.H;
    const int slot = tr_slot9(lane);
    const bool owner = slot >= 0 && !(lane & 1);
    const uint32_t a_ra = smem_addr(s_ra), a_rb = smem_addr(s_rb), a_rgb = smem_addr(s_rgb);
    const uint32_t a_acc = smem_addr(s_acc) + 4u * (uint32_t)(slot < 0 ? 0 : slot);
    const uint32_t a_cw = smem_addr(s_cw) + (uint32_t)(threadIdx.x >> 5) * CW_WARP_BYTES;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // only the first max(n_contrib) entries of the list matter: per tile for staging, per warp for work
    if (threadIdx.x == 0) s_max = 0;
#pragma unroll
    for (int i = threadIdx.x; i < ROUND * 3; i += BLK) s_acc[i] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, max(last_a, last_b));
    if (lane == 0) atomicMax(&s_max, wmax);
    __syncthreads();
    const int count = s_max;
    if (count == 0) return;

    const int rounds = (count + ROUND - 1) / ROUND;
    for (int r = rounds - 1; r >= 0; --r) {
#pragma unroll
        for (int h = 0; h < ROUND / BLK; ++h) {
            const int st = h * BLK + threadIdx.x, idx = r * ROUND + st;
            if (idx < count) {
                const uint32_t g = gidx[rng.x + idx];
                s_id[st] = g;
                s_ra[st] = rec[2 * (size_t)g];
                s_rb[st] = rec[2 * (size_t)g + 1];
                s_rgb[st] = rgb4[g];
            }
        }
        __syncthreads();
        if (r * ROUND < wmax) {  // else nothing in this round is a contributor for this warp
            const int nb = min(ROUND, min(count, wmax) - r * ROUND);
            const int lim_a = last_a - r * ROUND, lim_b = last_b - r * ROUND;  // staged j < lim: before the last contributor
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
                if (keep) cw_store(a_cw, __popc(m & lt_mask), ra, rb, e);  // compact the survivors in list order
                __syncwarp();
                // back to front over the survivors: the alpha arithmetic (the forward's, so the contributor set is
                // identical) runs packed for the lane's two pixels, the recurrences one by one
                for (int t = n - 1; t >= 0; --t) {
                    const uint32_t cw = a_cw + CW_REC * (uint32_t)t;
                    f32x2 power2, alpha2, dx2, dy2, G2;
                    alpha_two(cw, npx, npy, power2, alpha2, dx2, dy2, G2);
                    float pw0, pw1, al0, al1, G0, G1;
                    upk2(power2, pw0, pw1);
                    upk2(alpha2, al0, al1);
                    upk2(G2, G0, G1);
                    const int j = (int)lds32(cw + 48u);
                    const bool hit_a = j < lim_a && pw0 <= 0.0f && al0 >= 1.0f / 255.0f;
                    const bool hit_b = j < lim_b && pw1 <= 0.0f && al1 >= 1.0f / 255.0f;
                    if (!__any_sync(0xffffffffu, hit_a || hit_b)) continue;
                    // per-pixel work stops at cg = G * dL/dalpha and w = alpha * T; pixels that do not contribute keep
                    // both at zero, so the products below need no other masking
                    float cga = 0.0f, wa = 0.0f, cgb = 0.0f, wb = 0.0f;
                    const float4 c = lds128(a_rgb + 16u * j);
                    if (hit_a) {
                        // one refined reciprocal replaces the IEEE divisions by (1 - alpha) (no FCHK / slow-path
                        // branches; operands are in [0.01, 1] so no special cases exist)
                        const float oma = 1.0f - al0;
                        const float inv = rcp_nr(oma);
                        const float t0 = Ta * inv;  // T / (1 - alpha), residual-corrected: the error must not
                        Ta = fma_(fma_(-t0, oma, Ta), inv, t0);  // accumulate along the list
                        wa = al0 * Ta;
                        const float cd = fma_(c.z, da2, fma_(c.y, da1, c.x * da0));
                        // the colour behind this entry enters only through its dot product with dL/dpixel
                        cga = G0 * fma_(bgTa, inv, (cd - bha) * Ta);
                        bha = fma_(al0, cd, oma * bha);
                    }
                    if (hit_b) {
                        const float oma = 1.0f - al1;
                        const float inv = rcp_nr(oma);
                        const float t0 = Tb * inv;
                        Tb = fma_(fma_(-t0, oma, Tb), inv, t0);
                        wb = al1 * Tb;
                        const float cd = fma_(c.z, db2, fma_(c.y, db1, c.x * db0));
                        cgb = G1 * fma_(bgTb, inv, (cd - bhb) * Tb);
                        bhb = fma_(al1, cd, oma * bhb);
                    }
                    // moments of cg about the Gaussian's centre (the flush below turns the tile's sums into
                    // dL/dmean2D and dL/dconic) and the colour gradient: packed over the two pixels, then added
                    const f32x2 cg2 = pk2(cga, cgb), w2 = pk2(wa, wb);
                    const f32x2 cgx = mul2(cg2, dx2), cgy = mul2(cg2, dy2);
                    const f32x2 q[9] = {cgx, cgy, mul2(cgx, dx2), mul2(cgx, dy2), mul2(cgy, dy2), cg2,
                                        mul2(w2, dp0), mul2(w2, dp1), mul2(w2, dp2)};
                    float v[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        float lo, hi;
                        upk2(q[i], lo, hi);
                        v[i] = lo + hi;
                    }
                    tr_reduce<9, 16>(v, lane);
                    if (owner) reds_add(a_acc + 48u * j, v[0]);
                }
                __syncwarp();  // the buffer is rewritten by the next group
            }
        }
        __syncthreads();
        // flush the round's tile-level sums: three 16-byte vector reductions per touched Gaussian
#pragma unroll
        for (int h = 0; h < ROUND / BLK; ++h) {
            const int st = h * BLK + threadIdx.x, idx = r * ROUND + st;
            if (idx < count) {
                float4 g0 = s_acc[3 * st], g1 = s_acc[3 * st + 1];
                const float4 g2 = s_acc[3 * st + 2];
                const bool nz = g0.x != 0.0f || g0.y != 0.0f || g0.z != 0.0f || g0.w != 0.0f || g1.x != 0.0f ||
                                g1.y != 0.0f || g1.z != 0.0f || g1.w != 0.0f || g2.x != 0.0f;
                if (nz) {
                    // moments -> gradients: dG/ddelta = -G (A dx + B dy, B dx + C dy), dG/dconic = -0.5 G (dx^2, dx dy, dy^2)
                    const float4 ra = s_ra[st], rb = s_rb[st];
                    const float opx = -rb.y * ddelx_dx, opy = -rb.y * ddely_dy, oph = -0.5f * rb.y;
                    const float mx = g0.x, my = g0.y;
                    g0.x = opx * fma_(ra.w, my, ra.z * mx);
                    g0.y = opy * fma_(rb.x, my, ra.w * mx);
                    g0.z *= oph; g0.w *= oph; g1.x *= oph;
                    float *dst = grad_blend + 12 * (size_t)s_id[st];
                    red_global_v4(dst, g0);
                    red_global_v4(dst + 4, g1);
                    red_global_v4(dst + 8, g2);
                    s_acc[3 * st] = make_float4(0, 0, 0, 0);
                    s_acc[3 * st + 1] = make_float4(0, 0, 0, 0);
                    s_acc[3 * st + 2] = make_float4(0, 0, 0, 0);
                }
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
