// blend_common.cuh -- pieces shared by the forward (blend.cu) and backward (blend_bwd.cu) blend kernels.
#pragma once
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int BLK = 256;
#ifndef BWD_MIN_BLOCKS
#define BWD_MIN_BLOCKS 5
#endif

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

__device__ __forceinline__ void warp_rect(int &px0, int &py0)
{
    const int w = threadIdx.x >> 5;
    px0 = blockIdx.x * DMGS_TILE + (w & 1) * 8;
    py0 = blockIdx.y * DMGS_TILE + (w >> 1) * 4;
}

}  // namespace dmgs
