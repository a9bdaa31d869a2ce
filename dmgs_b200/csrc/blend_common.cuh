// blend_common.cuh -- pieces shared by the forward (blend.cu) and backward (blend_bwd.cu) blend kernels.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

// A 16x16 tile is four 8x8 squares, one per warp, two pixels per lane.  The warps never synchronise with each
// other, so the CTA size only decides how soon a finished warp gives its slot back (and how the four squares of a
// tile share L1): squares are numbered 4 * tile + quadrant and dealt to the warps of a 1-D grid.
#ifndef BLEND_BLK
#define BLEND_BLK 128
#endif
constexpr int BLK = BLEND_BLK;
#ifndef FWD_MIN_BLOCKS
#define FWD_MIN_BLOCKS 8
#endif
#ifndef BWD_MIN_BLOCKS
#define BWD_MIN_BLOCKS 8
#endif

// unroll factor of the survivor loops (experiments: more ILP per warp at a lower residency)
#ifndef FWD_UNROLL
#define FWD_UNROLL 1
#endif
#ifndef BWD_UNROLL
#define BWD_UNROLL 1
#endif
#define DMGS_PRAGMA_(x) _Pragma(#x)
#define DMGS_UNROLL(n) DMGS_PRAGMA_(unroll n)

struct BlendArgs {
    int W, H, gx, gy;
    float bg[3];
};

// Shared-memory reads in the inner loops go through explicit 32-bit shared addresses: with C++
// indexing nvcc rebuilds the cluster-window base (S2UR SR_CgaCtaId + ULEA) inside the hot loop,
// ~12 instructions and a scoreboard stall per entry (profiles/r1a).
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));  // volatile: the value is kept in a register, never rematerialised
    return r;
}
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void reds_add(uint32_t a, float v)
{
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void red_global_v4(float *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 1/x for x in [0.01, 1]: MUFU.RCP refined by one Newton step (error < 1 ulp, no slow path)
__device__ __forceinline__ float rcp_nr(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fma_(r, fma_(-x, r, 1.0f), r);
}

// ---- packed FP32 (sm_100: FFMA2 / FMUL2 / FADD2, two IEEE round-to-nearest results per instruction).  The blend
// kernels are bound by instruction ISSUE, not by the FP32 pipes (ncu: issue 70-80 %, fma pipe 40-48 %), so evaluating
// a lane's two pixels per instruction cuts the time.  Element-wise the results are those of the scalar operations,
// which keeps every forward decision bit-identical.  NOTE: never write mul2 followed by add2 where the contract wants two
// roundings -- ptxas 12.9 contracts that pair into FFMA2 even with -fmad=false; use fma2 (and say so in the contract).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 lds64(uint32_t a)
{
    f32x2 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32u(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Per-warp compacted survivor list of one 32-entry group.  A lane owns TWO pixels (same column, rows v and v + 4 of
// the warp's 8x8 square) and evaluates both with packed FP32, so every geometric field is stored twice: one
// broadcast 16-byte shared load yields two ready-made register pairs.  64 bytes per survivor:
//   {x, x, y, y} {A, A, -B, -B} {C, C, opacity, opacity} {r, g, b, position in the tile's list (u32 bits)}
constexpr uint32_t CW_REC = 64u;
constexpr uint32_t CW_WARP_BYTES = 32u * CW_REC;  // 2 KB per warp

__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void lds128x2(uint32_t a, f32x2 &lo, f32x2 &hi)
{
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(a));
}
// survivor `slot` of the warp's list <- list entry number `pos` (ra = x, y, A, B; rb = C, opacity, cut, -; colour)
__device__ __forceinline__ void cw_store(uint32_t a_cw, int slot, const float4 &ra, const float4 &rb, const float4 &col,
                                         uint32_t pos)
{
    const uint32_t w = a_cw + CW_REC * (uint32_t)slot;
    sts128(w, ra.x, ra.x, ra.y, ra.y);
    sts128(w + 16u, ra.z, ra.z, -ra.w, -ra.w);
    sts128(w + 32u, rb.x, rb.x, rb.y, rb.y);
    sts128(w + 48u, col.x, col.y, col.z, __uint_as_float(pos));
}
// read-only 16-byte gather (records and colours are written by preprocess, a different kernel)
__device__ __forceinline__ float4 ldg128(const float4 *p) { return __ldg(p); }

// alpha of ONE list entry for the lane's TWO pixels, in the arithmetic contract, element-wise:
//   dx = x - px, dy = y - py;  q = fma(dx, A dx, (C dy) dy);  power = fma(q, -0.5, (-B dx) dy)
//   G = dmgs_exp(power) (for power <= 0 only max(power, -80) of its clamp can act);  alpha = min(0.99, opacity G)
// cw: shared address of the survivor record; npx, npy: (-px, -px), (-py_a, -py_b).  alpha is meaningless where
// power > 0.
__device__ __forceinline__ void alpha_two(uint32_t cw, f32x2 npx, f32x2 npy, f32x2 &power, f32x2 &alpha, f32x2 &dxo, f32x2 &dyo,
                                          f32x2 &Go)
{
    f32x2 X, Y, A, NB, Cc, OP;
    lds128x2(cw, X, Y);
    lds128x2(cw + 16u, A, NB);
    lds128x2(cw + 32u, Cc, OP);
    const f32x2 dx = add2(X, npx), dy = add2(Y, npy);
    const f32x2 t1 = mul2(A, dx);
    const f32x2 t2 = mul2(mul2(Cc, dy), dy);
    const f32x2 t3 = mul2(mul2(NB, dx), dy);
    const f32x2 q = fma2(dx, t1, t2);
    power = fma2(q, pk2(-0.5f, -0.5f), t3);
    float p0, p1;
    upk2(power, p0, p1);
    const f32x2 xc = pk2(fmaxf(p0, -80.0f), fmaxf(p1, -80.0f));
    const f32x2 r = fma2(xc, pk2(1.44269502162933349609375f, 1.44269502162933349609375f), pk2(12582912.0f, 12582912.0f));
    const f32x2 jf = add2(r, pk2(-12582912.0f, -12582912.0f));
    f32x2 f = fma2(jf, pk2(-0.693145751953125f, -0.693145751953125f), xc);
    f = fma2(jf, pk2(-1.42860682030941723212e-6f, -1.42860682030941723212e-6f), f);
    f32x2 p = fma2(pk2(0x1.6d8360p-10f, 0x1.6d8360p-10f), f, pk2(0x1.127dd8p-7f, 0x1.127dd8p-7f));
    p = fma2(p, f, pk2(0x1.55549ep-5f, 0x1.55549ep-5f));
    p = fma2(p, f, pk2(0x1.5553e8p-3f, 0x1.5553e8p-3f));
    p = fma2(p, f, pk2(0.5f, 0.5f));
    p = fma2(p, f, pk2(1.0f, 1.0f));
    p = fma2(p, f, pk2(1.0f, 1.0f));
    float r0, r1, e0, e1;
    upk2(r, r0, r1);
    upk2(p, e0, e1);
    // exponent insert: bits(p) + (j << 23), j = bits(r) - 0x4B400000 (whose low 23 bits are zero: the shift drops it)
    e0 = __int_as_float(__float_as_int(e0) + (__float_as_int(r0) << 23));
    e1 = __int_as_float(__float_as_int(e1) + (__float_as_int(r1) << 23));
    Go = pk2(e0, e1);
    const f32x2 a = mul2(OP, Go);
    float a0, a1;
    upk2(a, a0, a1);
    alpha = pk2(fminf(0.99f, a0), fminf(0.99f, a1));
    dxo = dx;
    dyo = dy;
}

// exact minimum over the pixel rectangle [x0,x1]x[y0,y1] of q(d) = A dx^2 + 2 B dx dy + C dy^2,
// d = g - p, for a positive-definite conic; returns true when the entry can be skipped for every
// pixel of the rectangle: 0.5*q_min > cut (+ rounding margin).  cut = log(255*opacity) + 1e-3
// (preprocess), +inf for conics that are not positive definite, negative for opacity <= 1/255.
__device__ __forceinline__ bool cull_rect(float gx, float gy, float A, float B, float C, float cut, float x0,
                                          float x1, float y0, float y1)
{
    const float cx = fminf(fmaxf(gx, x0), x1), cy = fminf(fmaxf(gy, y0), y1);
    const float dx = gx - cx, dy = gy - cy;  // offset to the closest point of the rectangle
    float qmin = 0.0f, mag = 0.0f;
    if (dx != 0.0f || dy != 0.0f) {
        qmin = 3.0e38f;
        if (dx != 0.0f) {  // vertical edge x = cx: minimise over y in [y0, y1]
            const float ys = gy + __fdividef(B * dx, C);  // stationary point of q along the edge (an approximate
            // quotient moves the evaluation point by < 1e-3 px: q changes by < C * 1e-6, inside the margin below)
            const float dyc = gy - fminf(fmaxf(ys, y0), y1);
            const float t0 = A * dx * dx, t1 = 2.0f * B * dx * dyc, t2 = C * dyc * dyc;
            qmin = t0 + t1 + t2;
            mag = t0 + fabsf(t1) + t2;
        }
        if (dy != 0.0f) {  // horizontal edge y = cy
            const float xs = gx + __fdividef(B * dy, A);
            const float dxc = gx - fminf(fmaxf(xs, x0), x1);
            const float t0 = A * dxc * dxc, t1 = 2.0f * B * dxc * dy, t2 = C * dy * dy;
            const float q2 = t0 + t1 + t2;
            if (q2 < qmin) { qmin = q2; mag = t0 + fabsf(t1) + t2; }
        }
    }
    // skip only when certainly below the 1/255 threshold everywhere (NaNs compare false -> keep)
    return 0.5f * qmin > cut + 1.0e-5f * mag + 1.0e-3f;
}

// square sq = 4 * tile + q owns the 8x8 pixels (q & 1, q >> 1) of its tile; lane -> column lane & 7, rows lane >> 3
// and + 4.  tile_order (optional): tiles by descending list length -- the squares of the longest lists are handed
// out first.
//
// PERSISTENT WARPS: the grid is at most `blend_residency` CTAs per SM; a warp's first square is its global warp
// number, every further one comes from a device counter (zero at launch).  Two reasons: (1) dynamic hand-out in
// longest-first order keeps the tail short (with one square per warp the resident-warp average was 78 % of the
// theoretical 32 per SM); (2) a residency below 8 CTAs/SM leaves registers and warp slots of every SM to OTHER
// streams' kernels -- with 8 the blend kernels own all 64 K registers for their whole run and the latency-bound
// stages of other views (depth sort, tile placement) cannot start before a blend kernel drains (DESIGN.md
// section 6).  Returns false when the squares are used up.
__device__ __forceinline__ bool next_square(const BlendArgs &a, const uint32_t *__restrict__ tile_order,
                                            uint32_t *__restrict__ counter, bool first, int lane, int &tile, int &px0,
                                            int &py0)
{
    uint32_t sq;
    if (first) {
        sq = blockIdx.x * (BLK / 32) + (threadIdx.x >> 5);
    } else {
        sq = 0;
        if (lane == 0) sq = gridDim.x * (BLK / 32) + atomicAdd(counter, 1u);
        sq = __shfl_sync(0xffffffffu, sq, 0);
    }
    if (sq >= 4u * (uint32_t)(a.gx * a.gy)) return false;
    tile = (int)(sq >> 2);
    if (tile_order) tile = (int)tile_order[tile];
    const int ty = tile / a.gx, tx = tile - ty * a.gx, q = (int)(sq & 3u);
    px0 = tx * DMGS_TILE + (q & 1) * 8;
    py0 = ty * DMGS_TILE + (q >> 1) * 8;
    return true;
}
// CTAs per SM of the blend kernels' persistent grids (1..8); see dmgs_set_blend_residency
extern int g_blend_residency[2];
__host__ inline int blend_grid(const BlendArgs &a, int which)
{
    const int all = (4 * a.gx * a.gy + BLK / 32 - 1) / (BLK / 32);
    const int cap = num_sms() * g_blend_residency[which];
    return all < cap ? all : cap;
}

}  // namespace dmgs
