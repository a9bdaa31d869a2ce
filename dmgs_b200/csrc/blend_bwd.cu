// blend_bwd.cu -- back-to-front adjoint of the per-tile alpha blend; see blend.cu for
// the tile / warp-square layout, the autonomous-warp walk and the culling scheme, which are shared.
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

constexpr uint32_t ACC_ROW = 48u;                       // 9 sums (+3 pad) per list entry of the current group
constexpr uint32_t ACC_WARP_BYTES = 32u * ACC_ROW;      // 1.5 KB per warp

__global__ void __launch_bounds__(BLK, BWD_MIN_BLOCKS)
blend_bwd_kernel(const __grid_constant__ BlendArgs a, const uint32_t *__restrict__ tile_order, const uint2 *__restrict__ ranges,
                 const uint32_t *__restrict__ gidx, const float4 *__restrict__ rec, const float4 *__restrict__ rgb4,
                 const float *__restrict__ final_T, const uint32_t *__restrict__ n_contrib,
                 const float *__restrict__ dL_dpix, float *__restrict__ grad_blend, uint32_t *__restrict__ counter)
{
    __shared__ __align__(16) unsigned char s_cw[(BLK / 32) * CW_WARP_BYTES];    // per-warp compacted survivors
    __shared__ __align__(16) unsigned char s_acc[(BLK / 32) * ACC_WARP_BYTES];  // per-warp sums of the current group

    const int lane = threadIdx.x & 31;
    const float ddelx_dx = 0.5f * (float)a.W, ddely_dy = 0.5f * (float)a.H;
    const int slot = tr_slot9(lane);
    const bool owner = slot >= 0 && !(lane & 1);
    const uint32_t a_cw = smem_addr(s_cw) + (uint32_t)(threadIdx.x >> 5) * CW_WARP_BYTES;
    const uint32_t a_acc = smem_addr(s_acc) + (uint32_t)(threadIdx.x >> 5) * ACC_WARP_BYTES;
    const uint32_t a_own = a_acc + 4u * (uint32_t)(slot < 0 ? 0 : slot);
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (bool first = true;; first = false) {  // persistent warp: one 8x8 square per iteration (blend_common.cuh)
    int tile, px0, py0;
    if (!next_square(a, tile_order, counter, first, lane, tile, px0, py0)) break;
    const int px = px0 + (lane & 7), pya = py0 + (lane >> 3), pyb = pya + 4;
    const bool in_a = px < a.W && pya < a.H, in_b = px < a.W && pyb < a.H;
    const size_t pix_a = (size_t)pya * a.W + px, pix_b = (size_t)pyb * a.W + px, HW = (size_t)a.H * a.W;
    const int last_a = in_a ? (int)n_contrib[pix_a] : 0, last_b = in_b ? (int)n_contrib[pix_b] : 0;
    // only the first max(n_contrib) entries of the list matter to this square
    const int count = __reduce_max_sync(0xffffffffu, max(last_a, last_b));
    if (count == 0) continue;

    const float pxf = (float)px;
    const f32x2 npx = pk2(-pxf, -pxf), npy = pk2(-(float)pya, -(float)pyb);
    const float rx0 = (float)px0, rx1 = (float)(px0 + 7), ry0 = (float)py0, ry1 = (float)(py0 + 7);
    const uint32_t *__restrict__ list = gidx + ranges[tile].x;

    // per pixel: T (transmittance in front of the entry being visited), `behind` = (colour blended behind that
    // entry) . dL/dpixel, normalised by the transmittance behind it, and the background's share bgT
    float Ta = in_a ? final_T[pix_a] : 0.0f, Tb = in_b ? final_T[pix_b] : 0.0f;
    float da0 = 0, da1 = 0, da2 = 0, db0 = 0, db1 = 0, db2 = 0;
    if (in_a) { da0 = dL_dpix[pix_a]; da1 = dL_dpix[HW + pix_a]; da2 = dL_dpix[2 * HW + pix_a]; }
    if (in_b) { db0 = dL_dpix[pix_b]; db1 = dL_dpix[HW + pix_b]; db2 = dL_dpix[2 * HW + pix_b]; }
    const float bgTa = -Ta * dot3(a.bg[0], da0, a.bg[1], da1, a.bg[2], da2);
    const float bgTb = -Tb * dot3(a.bg[0], db0, a.bg[1], db1, a.bg[2], db2);
    float bha = 0.0f, bhb = 0.0f;
    const f32x2 dp0 = pk2(da0, db0), dp1 = pk2(da1, db1), dp2 = pk2(da2, db2);

    // back to front, group by group; software pipeline as in the forward (records one group ahead, indices two)
    int base = ((count - 1) / 32) * 32;
    float4 ra = make_float4(0, 0, 0, 0), rb = ra, col = ra;
    uint32_t id_cur = 0, id_next = 0;
    if (base + lane < count) {
        id_cur = list[base + lane];
        ra = ldg128(rec + 2 * (size_t)id_cur);
        rb = ldg128(rec + 2 * (size_t)id_cur + 1);
        col = ldg128(rgb4 + id_cur);
    }
    if (base >= 32) id_next = list[base - 32 + lane];

    for (; base >= 0; base -= 32) {
        const bool keep = base + lane < count && !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, rx0, rx1, ry0, ry1);
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        const int n = __popc(m);
        const int my_slot = __popc(m & lt_mask);
        const uint32_t id_keep = id_cur;
        if (keep) {  // compact the survivors in list order; their rows of sums start at zero
            cw_store(a_cw, my_slot, ra, rb, col, (uint32_t)(base + lane));
            sts128(a_acc + ACC_ROW * lane, 0.0f, 0.0f, 0.0f, 0.0f);
            sts128(a_acc + ACC_ROW * lane + 16u, 0.0f, 0.0f, 0.0f, 0.0f);
            sts128(a_acc + ACC_ROW * lane + 32u, 0.0f, 0.0f, 0.0f, 0.0f);
        }
        id_cur = id_next;
        if (base >= 32) {
            ra = ldg128(rec + 2 * (size_t)id_cur);
            rb = ldg128(rec + 2 * (size_t)id_cur + 1);
            col = ldg128(rgb4 + id_cur);
        }
        if (base >= 64) id_next = list[base - 64 + lane];
        if (!n) continue;
        __syncwarp();
        // back to front over the survivors: the alpha arithmetic (the forward's, so the contributor set is
        // identical) runs packed for the lane's two pixels, the recurrences one by one
        DMGS_UNROLL(BWD_UNROLL)
        for (int t = n - 1; t >= 0; --t) {
            const uint32_t cw = a_cw + CW_REC * (uint32_t)t;
            f32x2 power2, alpha2, dx2, dy2, G2;
            alpha_two(cw, npx, npy, power2, alpha2, dx2, dy2, G2);
            float pw0, pw1, al0, al1, G0, G1;
            upk2(power2, pw0, pw1);
            upk2(alpha2, al0, al1);
            upk2(G2, G0, G1);
            const float4 c = lds128(cw + 48u);  // r, g, b, list position
            const int j = (int)__float_as_uint(c.w);
            const bool hit_a = j < last_a && pw0 <= 0.0f && al0 >= 1.0f / 255.0f;
            const bool hit_b = j < last_b && pw1 <= 0.0f && al1 >= 1.0f / 255.0f;
            if (!__any_sync(0xffffffffu, hit_a || hit_b)) continue;
            // per-pixel work stops at cg = G * dL/dalpha and w = alpha * T; pixels that do not contribute keep
            // both at zero, so the products below need no other masking
            float cga = 0.0f, wa = 0.0f, cgb = 0.0f, wb = 0.0f;
            if (hit_a) {
                // one refined reciprocal replaces the IEEE divisions by (1 - alpha) (no FCHK / slow-path
                // branches; operands are in [0.01, 1] so no special cases exist)
                const float oma = 1.0f - al0;
                const float inv = rcp_nr(oma);
                const float t0 = Ta * inv;  // T / (1 - alpha), residual-corrected: the error must not
                Ta = fma_(fma_(-t0, oma, Ta), inv, t0);  // accumulate along the list
                wa = al0 * Ta;
                // the colour behind this entry enters only through its dot product with dL/dpixel
                const float cd = fma_(c.z, da2, fma_(c.y, da1, c.x * da0));
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
            // moments of cg about the Gaussian's centre (the flush below turns the sums into dL/dmean2D and
            // dL/dconic) and the colour gradient: packed over the two pixels, then added
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
            // every list entry is visited once per square: a plain store into the entry's row, no atomics
            if (owner) sts32(a_own + ACC_ROW * (uint32_t)(j & 31), v[0]);
        }
        __syncwarp();
        // flush, lane <-> entry again: three 16-byte vector reductions per surviving entry of this square
        if (keep) {
            float4 g0 = lds128(a_acc + ACC_ROW * lane), g1 = lds128(a_acc + ACC_ROW * lane + 16u);
            const float4 g2 = lds128(a_acc + ACC_ROW * lane + 32u);
            const bool nz = g0.x != 0.0f || g0.y != 0.0f || g0.z != 0.0f || g0.w != 0.0f || g1.x != 0.0f ||
                            g1.y != 0.0f || g1.z != 0.0f || g1.w != 0.0f || g2.x != 0.0f;
            if (nz) {
                // moments -> gradients: dG/ddelta = -G (A dx + B dy, B dx + C dy), dG/dconic = -0.5 G (dx^2, dx dy, dy^2)
                const uint32_t w = a_cw + CW_REC * (uint32_t)my_slot;
                const float4 ab = lds128(w + 16u), co = lds128(w + 32u);  // A, A, -B, -B | C, C, opacity, opacity
                const float A = ab.x, B = -ab.z, Cc = co.x, op = co.z;
                const float opx = -op * ddelx_dx, opy = -op * ddely_dy, oph = -0.5f * op;
                const float mx = g0.x, my = g0.y;
                g0.x = opx * fma_(B, my, A * mx);
                g0.y = opy * fma_(Cc, my, B * mx);
                g0.z *= oph; g0.w *= oph; g1.x *= oph;
                float *dst = grad_blend + 12 * (size_t)id_keep;
                red_global_v4(dst, g0);
                red_global_v4(dst + 4, g1);
                red_global_v4(dst + 8, g2);
            }
        }
        __syncwarp();  // the buffers are rewritten by the next group
    }
    }  // next square
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
    // the square counter sits behind the P x 12 sums of the scratch buffer and is zeroed with them (api.cu)
    uint32_t *counter = reinterpret_cast<uint32_t *>(grad_blend + 12 * (size_t)prm->P);
    blend_bwd_kernel<<<blend_grid(a, 1), BLK, 0, s>>>(a, BL.has_order ? at<uint32_t>(binning, BL.tile_order) : nullptr, at<uint2>(binning, BL.ranges), at<uint32_t>(binning, BL.gidx),
                                                      at<float4>(geom, GL.rec), at<float4>(geom, GL.rgb),
                                                      at<float>(image, IL.final_T), at<uint32_t>(image, IL.n_contrib),
                                                      dL_dpix, grad_blend, counter);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
