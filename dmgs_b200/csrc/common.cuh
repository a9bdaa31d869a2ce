// common.cuh -- shared device helpers, state-buffer layouts and error plumbing.
//
// Arithmetic contract: the translation units are compiled with
// -fmad=false, so `a*b+c` is never contracted; every fused multiply-add is an explicit
// __fmaf_rn().  Division and sqrt are the IEEE-rounded defaults (no --use_fast_math), and
// exp() is dmgs_exp() below, a fixed sequence of IEEE operations.  The CPU oracle
// (oracle/splat_oracle.c) is written independently to the same contract; that is what
// makes keys, tile ranges, radii and contributor counts comparable bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dmgs_raster.h"

#define DMGS_TILE 16
#define DMGS_NEAR 0.2f
#define DMGS_DEFAULT_SMS 148    /* B200; used only when no device can be queried (host-only size calls) */
#define PLACE_MAX_TILES 16384   /* direct tile placement (place.cu) up to this many tiles, radix partition above */

namespace dmgs {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
int check_stage(const dmgs_params *prm, cudaStream_t s, const char *stage);
void count_launches(int n);  // kernels launched by this library (bench.py reports it)
// Per-DEVICE launch state (a process may drive several GPUs): multiprocessor count of the current device
// (DMGS_DEFAULT_SMS when none can be queried) and a once-per-device latch for the cudaFuncSetAttribute opt-ins.
int num_sms();
enum { ONCE_PREPROCESS_FWD = 0, ONCE_PREPROCESS_BWD, ONCE_SH_EXPAND, ONCE_PLACE, ONCE_BIND_FUSED, ONCE_TEXTURE_BWD, ONCE_SLOTS };
bool once_per_device(int slot);

#define DMGS_CUDA(call)                                                              \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            dmgs::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                            __FILE__, __LINE__);                                     \
            return (int)e__;                                                         \
        }                                                                            \
    } while (0)

// ----------------------------------------------------------------------------- layouts
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct GeomLayout {
    size_t depths, rec, rgb, clamped, cov3D, tiles, rect, order, offsets;  // inspection arrays
    size_t keys_a, keys_b, vals_b, hist, scan_tmp, stat, os_ghist, os_status, total;
    int sort_blocks, os_tiles;
};
struct BinLayout {
    size_t tiles, gidx, ranges;  // inspection arrays (final sorted list)
    size_t tiles_b, gidx_b, hist, scan_tmp;  // radix tile partition (sort.cu)
    size_t srec, table, gsum, tile_start;    // direct tile placement (place.cu)
    size_t tile_order;                       // tiles by descending list length (written with the ranges by place.cu)
    size_t total;
    int sort_blocks, has_order;
};
// direct tile placement plan (place.cu): nseg depth-ordered segments of seg Gaussians, one per warp
struct PlacePlan {
    int ok, wpb, seg, nseg, groups, rows_per_group;
    size_t smem;
};
PlacePlan place_plan(int32_t P, int T);      // the LAYOUT plan (200 KB of counters per SM): buffer sizes, path decision
PlacePlan place_plan_run(int32_t P, int T);  // the plan the kernels run with (dmgs_set_place_smem_kb): a prefix of its rows
extern int g_place_smem_kb;
struct ImgLayout {
    size_t final_T, n_contrib, counter, total;  // counter: next square of the persistent forward blend
};

constexpr int SORT_MIN_ITEMS_PER_BLOCK = 1024;  // 256 threads x 4 rounds (8 rounds above RADIX_SMALL_N items)
constexpr int64_t RADIX_SMALL_N = 2 * 1024 * 1024;
constexpr int SORT_MAX_BINS = 256;
constexpr int ONESWEEP_ITEMS = 2048;  // keys per tile of the depth sort (256 threads x 8)

static inline size_t scan_tmp_bytes(size_t n) { return align_up((n / 2048 + 2) * sizeof(uint32_t) * 2); }

static inline GeomLayout geom_layout(int32_t P)
{
    GeomLayout L;
    size_t n = (size_t)(P > 0 ? P : 1), o = 0;
    L.depths = o;  o += align_up(n * 4);
    L.rec = o;     o += align_up(n * 32);
    L.rgb = o;     o += align_up(n * 16);
    L.clamped = o; o += align_up(n);
    L.cov3D = o;   o += align_up(n * 24);
    L.tiles = o;   o += align_up(n * 4);
    L.rect = o;    o += align_up(n * 8);
    L.order = o;   o += align_up(n * 4);
    L.offsets = o; o += align_up(n * 4);
    L.keys_a = o;  o += align_up(n * 4);
    L.keys_b = o;  o += align_up(n * 4);
    L.vals_b = o;  o += align_up(n * 4);
    L.sort_blocks = (int)((n + SORT_MIN_ITEMS_PER_BLOCK - 1) / SORT_MIN_ITEMS_PER_BLOCK);
    size_t hist_n = (size_t)(L.sort_blocks + 1) * SORT_MAX_BINS + 1;  // table + bin totals
    L.hist = o;    o += align_up(hist_n * 4);
    size_t big = hist_n > n ? hist_n : n;
    L.scan_tmp = o; o += scan_tmp_bytes(big);
    // {max of ~key, max of key over the live depth keys, -, -, tile counters of the four onesweep passes}, then the
    // four global digit histograms: ONE memset clears both before preprocess
    L.stat = o;    o += 256;
    L.os_ghist = o; o += 4 * SORT_MAX_BINS * 4;
    L.os_tiles = (int)((n + ONESWEEP_ITEMS - 1) / ONESWEEP_ITEMS);
    L.os_status = o; o += align_up((size_t)4 * L.os_tiles * SORT_MAX_BINS * 4);  // decoupled look-back words
    L.total = o;
    return L;
}

static inline BinLayout bin_layout(int32_t P, int64_t R, int32_t W, int32_t H)
{
    BinLayout L;
    memset(&L, 0, sizeof(L));
    size_t n = (size_t)(R > 0 ? R : 1), o = 0;
    size_t T = (size_t)((W + DMGS_TILE - 1) / DMGS_TILE) * ((H + DMGS_TILE - 1) / DMGS_TILE);
    L.tiles = o;   o += align_up(n * 4);
    L.gidx = o;    o += align_up(n * 4);
    L.ranges = o;  o += align_up(T * 8);
    const PlacePlan pl = place_plan(P, (int)T);
    if (pl.ok) {
        L.srec = o;       o += align_up((size_t)(P > 0 ? P : 1) * 16);
        L.table = o;      o += align_up((size_t)pl.nseg * T * 4);
        L.gsum = o;       o += align_up((size_t)((pl.nseg + 1) / 2 + 1 > pl.groups ? (pl.nseg + 1) / 2 + 1 : pl.groups) * T * 4);
        L.tile_start = o; o += align_up(T * 4);
        L.tile_order = o; o += align_up(T * 4);
        L.has_order = (R > 0 && P > 0) ? 1 : 0;  // the cases in which the placement kernels run (api.cu)
    } else {
        L.tiles_b = o; o += align_up(n * 4);
        L.gidx_b = o;  o += align_up(n * 4);
        L.sort_blocks = (int)((n + SORT_MIN_ITEMS_PER_BLOCK - 1) / SORT_MIN_ITEMS_PER_BLOCK);
        size_t hist_n = (size_t)(L.sort_blocks + 1) * SORT_MAX_BINS + 1;  // table + bin totals
        L.hist = o;    o += align_up(hist_n * 4);
        L.scan_tmp = o; o += scan_tmp_bytes(hist_n);
    }
    L.total = o;
    return L;
}

static inline ImgLayout img_layout(int32_t W, int32_t H)
{
    ImgLayout L;
    size_t n = (size_t)W * H, o = 0;
    L.final_T = o;   o += align_up(n * 4);
    L.n_contrib = o; o += align_up(n * 4);
    L.counter = o;   o += align_up(4);
    L.total = o;
    return L;
}

template <typename T> static inline T *at(void *base, size_t off) { return reinterpret_cast<T *>((char *)base + off); }
template <typename T> static inline const T *at(const void *base, size_t off)
{
    return reinterpret_cast<const T *>((const char *)base + off);
}

// ----------------------------------------------------------------------------- device math
#ifdef __CUDACC__
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
// a0*b0 + a1*b1 + a2*b2 in the rounding order nvcc's default contraction gives that expression (the upstream
// rasteriser's GLM products and transformPoint helpers): the SECOND product is rounded on its own, the first and
// the third are fused -- fma(a2, b2, fma(a0, b0, a1*b1)) (oracle/upstream_arith.cu, scripts/arith_divergence.py)
__device__ __forceinline__ float dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    return fma_(a2, b2, fma_(a0, b0, a1 * b1));
}
// ndc2Pix as upstream writes it: ((v + 1.0) * S - 1.0) * 0.5 with DOUBLE literals, rounded to float once
__device__ __forceinline__ float ndc2pix(float v, int S)
{
    return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5);
}
// row r of a column-major 4x4 applied to (x, y, z, 1)
__device__ __forceinline__ float affine3(const float *m, int r, float x, float y, float z)
{
    return dot3(m[r], x, m[4 + r], y, m[8 + r], z) + m[12 + r];
}

// exp(x): clamp to [-80, 80], Cody-Waite reduction by ln2, degree-6 Horner, exponent insert.
__device__ __forceinline__ float dmgs_exp(float x)
{
    // explicit _rn operations: never contracted, whatever -fmad says for the translation unit
    x = fminf(fmaxf(x, -80.0f), 80.0f);
    // round(x * log2 e) by the magic-number trick, the product and the add FUSED: the packed-FP32 blend kernels
    // evaluate the same sequence with fma.rn.f32x2 (a separate mul.rn.f32x2 + add.rn.f32x2 pair is contracted by
    // ptxas 12.9 even under -fmad=false, so the contract makes the fusion explicit)
    const float r = fma_(x, 1.44269502162933349609375f, 12582912.0f);
    const float jf = __fadd_rn(r, -12582912.0f);
    const int j = __float_as_int(r) - 0x4B400000;
    float f = fma_(jf, -0.693145751953125f, x);
    f = fma_(jf, -1.42860682030941723212e-6f, f);
    float p = 0x1.6d8360p-10f;
    p = fma_(p, f, 0x1.127dd8p-7f);
    p = fma_(p, f, 0x1.55549ep-5f);
    p = fma_(p, f, 0x1.5553e8p-3f);
    p = fma_(p, f, 0.5f);
    p = fma_(p, f, 1.0f);
    p = fma_(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (j << 23));
}

// 16-byte vector reduction into global memory (REDG.E.ADD.F32x4): fire-and-forget, resolved in L2.  The accumulate
// modes of the per-Gaussian backward use reductions (not load-add-store) so that the views of a step running on
// different CUDA streams can add into ONE gradient buffer.
__device__ __forceinline__ void red_add_v4(float *p, float4 v)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ int clampi_f(float v, int hi) { return (int)fminf(fmaxf(v, 0.0f), (float)hi); }

// per-launch constants shared by the per-Gaussian kernels
struct DevParams {
    int P, sh_degree, M, W, H, sh_layout, sh_act;
    int gx, gy;
    float tanfovx, tanfovy, fx, fy, limx, limy, mod;
    float bg[3];
    float V[16];
    float PV[16];
    float cam[3];
};
#endif

}  // namespace dmgs
