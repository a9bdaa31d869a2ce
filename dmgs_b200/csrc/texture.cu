// texture.cu -- the stage-2 colour field: multiresolution hash-grid encoding + bias-free 3-layer MLP
// (/root/reference/geo/texture.py:47-111 `MLPTexture3D.sample_noact`, consumer
// scene/gaussian_geo_model_mlp_flex.py:313: features = sample_noact(gs_xyz).view(N, 3, 16)).
//
// The encoder of the reference is tinycudann's "HashGrid" (16 levels x 2 features, 2^19 entries per level, base
// resolution 16, finest 4096; fp16 parameters and output); its arithmetic lives in that third-party package, which
// is absent from the reference checkout and from this image: the level table, the dense / hashed index and the
// trilinear interpolation follow its published definition (oracle/texture_oracle.py restates them; parity
// UNPINNED).  The MLP (Linear(32,32) ReLU Linear(32,32) ReLU Linear(32,C), no bias, fp32) is the reference's own
// torch module.
//
// Forward: one thread per point -- 16 levels x 8 corner gathers of one half2 (the 24 MB table is L2 resident),
// features rounded to fp16 as the reference's encoder returns them, then the three layers on registers with the
// weights broadcast from shared memory.  Backward: a block of 128 points stages activations and their adjoints in
// shared memory ([feature][point], conflict free): phase A (thread <-> point) recomputes the activations and runs
// the adjoint chain, phase B (thread <-> weight entries) contracts the block's 128 points into the three weight
// gradients (3584 entries, one reduction each per block), phase C (thread <-> point) scatters the encoding
// gradient into the table with 8-byte vector reductions and forms dL/dxyz from the trilinear weights' derivative.
// Nothing here is tensor-core work at fp32 parity (TF32 would break the 1e-5 output bar), and the flops are small:
// 3584 FMA per point and direction.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int TX_LEVELS = 16, TX_FEAT = 2, TX_IN = TX_LEVELS * TX_FEAT, TX_HID = 32, TX_MAX_OUT = 64;
constexpr int TX_THREADS = 128;

struct TexLevels {
    float scale[TX_LEVELS];
    uint32_t res[TX_LEVELS], size[TX_LEVELS], offset[TX_LEVELS];  // entries (pairs of features), not floats
};
struct TexArgs {
    int64_t N;
    int C;
    float lo[3], hi[3];
    TexLevels lv;
};

// tinycudann's level table: scale_l = exp2(l * log2(per_level_scale)) * base - 1 (fp32), resolution = ceil(scale) + 1,
// entries = min(resolution^3 rounded up to 8, 2^log2_hashmap_size)
static TexLevels make_levels(uint32_t *total_entries)
{
    TexLevels lv;
    const float pls = (float)exp(log(4096.0 / 16.0) / (TX_LEVELS - 1));  // geo/texture.py:54-55 (float64 -> json float)
    const float log2_pls = log2f(pls);
    uint32_t off = 0;
    for (int l = 0; l < TX_LEVELS; ++l) {
        const float s = exp2f((float)l * log2_pls) * 16.0f - 1.0f;
        const uint32_t r = (uint32_t)ceilf(s) + 1u;
        uint64_t n = (uint64_t)r * r * r;
        n = (n + 7) / 8 * 8;
        if (n > (1u << 19)) n = 1u << 19;
        lv.scale[l] = s; lv.res[l] = r; lv.size[l] = (uint32_t)n; lv.offset[l] = off;
        off += (uint32_t)n;
    }
    *total_entries = off;
    return lv;
}

int64_t texture_grid_params()
{
    uint32_t e;
    make_levels(&e);
    return (int64_t)e * TX_FEAT;
}

__device__ __forceinline__ uint32_t tex_index(uint32_t x, uint32_t y, uint32_t z, uint32_t res, uint32_t size)
{
    // dense while the strides fit the level's table, the coherent prime hash otherwise
    uint32_t stride = 1, idx = 0;
    idx += x * stride; stride *= res;
    if (stride <= size) { idx += y * stride; stride *= res; }
    if (stride <= size) { idx += z * stride; stride *= res; }
    if (size < stride) idx = x ^ (y * 2654435761u) ^ (z * 805459861u);
    return idx % size;
}

__device__ __forceinline__ void tex_coords(const TexArgs &a, const float *__restrict__ xyz, int64_t n, float t[3], bool inside[3])
{
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float v = (xyz[3 * n + d] - a.lo[d]) / (a.hi[d] - a.lo[d]);
        inside[d] = v >= 0.0f && v <= 1.0f;  // torch.clamp passes the gradient inside [min, max]
        t[d] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

// features of one level for one point (fp32 accumulation of the eight corners, in corner order 0..7)
__device__ __forceinline__ float2 tex_level(const TexArgs &a, int l, const float t[3], const __half2 *__restrict__ grid)
{
    const float s = a.lv.scale[l];
    float w[3];
    uint32_t g[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float pos = fma_(s, t[d], 0.5f);
        const float fl = floorf(pos);
        w[d] = pos - fl;
        g[d] = (uint32_t)fl;
    }
    const uint32_t res = a.lv.res[l], size = a.lv.size[l];
    const __half2 *lev = grid + a.lv.offset[l];
    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float wx = (c & 1) ? w[0] : 1.0f - w[0], wy = (c & 2) ? w[1] : 1.0f - w[1], wz = (c & 4) ? w[2] : 1.0f - w[2];
        const float wgt = wx * wy * wz;
        const float2 v = __half22float2(lev[tex_index(g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + ((c >> 2) & 1), res, size)]);
        acc.x = fma_(wgt, v.x, acc.x);
        acc.y = fma_(wgt, v.y, acc.y);
    }
    return acc;
}

__global__ void __launch_bounds__(256) tex_cast_kernel(int64_t n2, const float2 *__restrict__ src, __half2 *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) dst[i] = __float22half2_rn(src[i]);
}

// y[j] = sum_i W[j][i] x[i] for j < OUT; W row-major [OUT][32] in shared memory (broadcast 16-byte loads)
template <int OUT, bool RELU>
__device__ __forceinline__ void tex_layer(const float *__restrict__ sW, const float (&x)[32], float *y)
{
#pragma unroll
    for (int j = 0; j < OUT; ++j) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(sW + j * 32 + i);
            acc = fma_(wv.x, x[i], acc); acc = fma_(wv.y, x[i + 1], acc);
            acc = fma_(wv.z, x[i + 2], acc); acc = fma_(wv.w, x[i + 3], acc);
        }
        y[j] = RELU ? fmaxf(acc, 0.0f) : acc;
    }
}

__global__ void __launch_bounds__(TX_THREADS)
texture_fwd_kernel(const __grid_constant__ TexArgs a, const float *__restrict__ xyz, const __half2 *__restrict__ grid,
                   const float *__restrict__ W0, const float *__restrict__ W1, const float *__restrict__ W2,
                   float *__restrict__ out, __half *__restrict__ enc_out)
{
    __shared__ __align__(16) float sW0[TX_HID * TX_IN], sW1[TX_HID * TX_HID], sW2[TX_MAX_OUT * TX_HID];
    for (int i = threadIdx.x; i < TX_HID * TX_IN; i += TX_THREADS) { sW0[i] = W0[i]; sW1[i] = W1[i]; }
    for (int i = threadIdx.x; i < a.C * TX_HID; i += TX_THREADS) sW2[i] = W2[i];
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * TX_THREADS + threadIdx.x;
    if (n >= a.N) return;
    float t[3];
    bool inside[3];
    tex_coords(a, xyz, n, t, inside);
    float x[TX_IN];
#pragma unroll
    for (int l = 0; l < TX_LEVELS; ++l) {
        const float2 f = tex_level(a, l, t, grid);
        const __half2 h = __float22half2_rn(f);  // the encoder returns fp16
        if (enc_out) reinterpret_cast<__half2 *>(enc_out + n * TX_IN)[l] = h;
        const float2 r = __half22float2(h);
        x[2 * l] = r.x; x[2 * l + 1] = r.y;
    }
    float h1[TX_HID], h2[TX_HID];
    tex_layer<TX_HID, true>(sW0, x, h1);
    tex_layer<TX_HID, true>(sW1, h1, h2);
    float *o = out + n * a.C;
    for (int c0 = 0; c0 < a.C; c0 += 4) {  // C is a multiple of 4 (checked by the launcher)
        float y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(sW2 + (c0 + k) * 32 + i);
                acc = fma_(wv.x, h2[i], acc); acc = fma_(wv.y, h2[i + 1], acc);
                acc = fma_(wv.z, h2[i + 2], acc); acc = fma_(wv.w, h2[i + 3], acc);
            }
            y[k] = acc;
        }
        *reinterpret_cast<float4 *>(o + c0) = make_float4(y[0], y[1], y[2], y[3]);
    }
}

// ------------------------------------------------------------------------------ backward
// shared-memory panels, [feature][point] with the point index fastest
struct TexBwdSmem {
    float x[TX_IN][TX_THREADS], h1[TX_HID][TX_THREADS], h2[TX_HID][TX_THREADS];
    float d1[TX_HID][TX_THREADS], d2[TX_HID][TX_THREADS], dout[TX_MAX_OUT][TX_THREADS];
};

__device__ __forceinline__ void red_add_v2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

__global__ void __launch_bounds__(TX_THREADS)
texture_bwd_kernel(const __grid_constant__ TexArgs a, const float *__restrict__ xyz, const __half2 *__restrict__ grid,
                   const __half *__restrict__ enc, const float *__restrict__ W0, const float *__restrict__ W1,
                   const float *__restrict__ W2, const float *__restrict__ dL_dout, float grid_grad_scale,
                   float *__restrict__ d_grid, float *__restrict__ dW0, float *__restrict__ dW1, float *__restrict__ dW2,
                   float *__restrict__ d_xyz)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TexBwdSmem &S = *reinterpret_cast<TexBwdSmem *>(smem_raw);
    float *sW0 = reinterpret_cast<float *>(smem_raw + sizeof(TexBwdSmem));
    float *sW1 = sW0 + TX_HID * TX_IN, *sW2 = sW1 + TX_HID * TX_HID;
    for (int i = threadIdx.x; i < TX_HID * TX_IN; i += TX_THREADS) { sW0[i] = W0[i]; sW1[i] = W1[i]; }
    for (int i = threadIdx.x; i < a.C * TX_HID; i += TX_THREADS) sW2[i] = W2[i];
    __syncthreads();
    const int p = threadIdx.x;
    const int64_t n = (int64_t)blockIdx.x * TX_THREADS + p;
    const bool live = n < a.N;
    float dx[TX_IN];
    // ---- phase A: thread <-> point.  Recompute the activations, run the adjoint chain, park everything the
    // weight gradients need in the panels (dead points contribute zeros).
    {
        float x[TX_IN], h1[TX_HID], h2[TX_HID];
#pragma unroll
        for (int l = 0; l < TX_LEVELS; ++l) {
            const float2 r = live ? __half22float2(reinterpret_cast<const __half2 *>(enc + n * TX_IN)[l]) : make_float2(0.0f, 0.0f);
            x[2 * l] = r.x; x[2 * l + 1] = r.y;
        }
        tex_layer<TX_HID, true>(sW0, x, h1);
        tex_layer<TX_HID, true>(sW1, h1, h2);
#pragma unroll
        for (int i = 0; i < TX_IN; ++i) { S.x[i][p] = x[i]; S.h1[i][p] = h1[i]; S.h2[i][p] = h2[i]; }
        // dh2 = W2^T dout, masked
        float d2[TX_HID];
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) d2[j] = 0.0f;
        for (int c = 0; c < a.C; ++c) {
            const float g = live ? dL_dout[n * a.C + c] : 0.0f;
            S.dout[c][p] = g;
#pragma unroll
            for (int j = 0; j < TX_HID; j += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(sW2 + c * 32 + j);
                d2[j] = fma_(wv.x, g, d2[j]); d2[j + 1] = fma_(wv.y, g, d2[j + 1]);
                d2[j + 2] = fma_(wv.z, g, d2[j + 2]); d2[j + 3] = fma_(wv.w, g, d2[j + 3]);
            }
        }
        float d1[TX_HID];
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) { d2[j] = h2[j] > 0.0f ? d2[j] : 0.0f; S.d2[j][p] = d2[j]; d1[j] = 0.0f; }
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) {  // dh1 = W1^T dh2
#pragma unroll
            for (int i = 0; i < TX_HID; i += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(sW1 + j * 32 + i);
                d1[i] = fma_(wv.x, d2[j], d1[i]); d1[i + 1] = fma_(wv.y, d2[j], d1[i + 1]);
                d1[i + 2] = fma_(wv.z, d2[j], d1[i + 2]); d1[i + 3] = fma_(wv.w, d2[j], d1[i + 3]);
            }
        }
#pragma unroll
        for (int i = 0; i < TX_IN; ++i) { d1[i] = h1[i] > 0.0f ? d1[i] : 0.0f; S.d1[i][p] = d1[i]; dx[i] = 0.0f; }
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) {  // dx = W0^T dh1
#pragma unroll
            for (int i = 0; i < TX_IN; i += 4) {
                const float4 wv = *reinterpret_cast<const float4 *>(sW0 + j * 32 + i);
                dx[i] = fma_(wv.x, d1[j], dx[i]); dx[i + 1] = fma_(wv.y, d1[j], dx[i + 1]);
                dx[i + 2] = fma_(wv.z, d1[j], dx[i + 2]); dx[i + 3] = fma_(wv.w, d1[j], dx[i + 3]);
            }
        }
    }
    __syncthreads();
    // ---- phase B: thread <-> weight entries.  entry e of [dW2 | dW1 | dW0] = <row of adjoints, row of activations>
    // over the block's 128 points; consecutive threads take consecutive entries of one output row, so the adjoint
    // row is a broadcast and the activation rows are distinct (conflict free)
    {
        const int n2 = a.C * TX_HID, total = n2 + 2 * TX_HID * TX_HID;
        for (int e = threadIdx.x; e < total; e += TX_THREADS) {
            const float *ra, *rb;
            float *dst;
            if (e < n2) { ra = S.dout[e >> 5]; rb = S.h2[e & 31]; dst = dW2 + e; }
            else if (e < n2 + TX_HID * TX_HID) { const int q = e - n2; ra = S.d2[q >> 5]; rb = S.h1[q & 31]; dst = dW1 + q; }
            else { const int q = e - n2 - TX_HID * TX_HID; ra = S.d1[q >> 5]; rb = S.x[q & 31]; dst = dW0 + q; }
            float acc = 0.0f;
#pragma unroll 8
            for (int k = 0; k < TX_THREADS; k += 4) {
                const float4 av = *reinterpret_cast<const float4 *>(ra + k), bv = *reinterpret_cast<const float4 *>(rb + k);
                acc = fma_(av.x, bv.x, acc); acc = fma_(av.y, bv.y, acc); acc = fma_(av.z, bv.z, acc); acc = fma_(av.w, bv.w, acc);
            }
            if (acc != 0.0f) atomicAdd(dst, acc);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TX_IN; ++i) S.x[i][p] = dx[i];  // own column only: read back below with a runtime level index
    if (!live) return;
    // ---- phase C: thread <-> point.  Scatter into the table (scaled as the reference's backward hooks scale the
    // encoder-parameter gradients) and dL/dxyz through the trilinear weights.
    float t[3];
    bool inside[3];
    tex_coords(a, xyz, n, t, inside);
    float dt[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll 1
    for (int l = 0; l < TX_LEVELS; ++l) {
        const float s = a.lv.scale[l];
        float w[3];
        uint32_t g[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float pos = fma_(s, t[d], 0.5f);
            const float fl = floorf(pos);
            w[d] = pos - fl;
            g[d] = (uint32_t)fl;
        }
        const uint32_t res = a.lv.res[l], size = a.lv.size[l], off = a.lv.offset[l];
        const float gx = S.x[2 * l][p], gy = S.x[2 * l + 1][p];
        float dw[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float wx = (c & 1) ? w[0] : 1.0f - w[0], wy = (c & 2) ? w[1] : 1.0f - w[1], wz = (c & 4) ? w[2] : 1.0f - w[2];
            const uint32_t idx = off + tex_index(g[0] + (c & 1), g[1] + ((c >> 1) & 1), g[2] + ((c >> 2) & 1), res, size);
            const float wgt = wx * wy * wz * grid_grad_scale;
            if (d_grid && (gx != 0.0f || gy != 0.0f)) red_add_v2(d_grid + 2 * (size_t)idx, wgt * gx, wgt * gy);
            if (d_xyz) {
                const float2 v = __half22float2(grid[idx]);
                const float gv = fma_(v.x, gx, v.y * gy);
                dw[0] += ((c & 1) ? 1.0f : -1.0f) * wy * wz * gv;
                dw[1] += ((c & 2) ? 1.0f : -1.0f) * wx * wz * gv;
                dw[2] += ((c & 4) ? 1.0f : -1.0f) * wx * wy * gv;
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) dt[d] = fma_(dw[d], s, dt[d]);
    }
    if (d_xyz) {
#pragma unroll
        for (int d = 0; d < 3; ++d) d_xyz[3 * n + d] = inside[d] ? dt[d] / (a.hi[d] - a.lo[d]) : 0.0f;
    }
}

static int fill_args(TexArgs &a, int64_t N, int C, const float *aabb6)
{
    if (N < 0 || C < 4 || C > TX_MAX_OUT || (C & 3)) { set_error("texture: channels must be a multiple of 4 in [4, %d] (got %d)", TX_MAX_OUT, C); return -14; }
    uint32_t e;
    a.N = N; a.C = C;
    for (int d = 0; d < 3; ++d) { a.lo[d] = aabb6[d]; a.hi[d] = aabb6[3 + d]; }
    a.lv = make_levels(&e);
    return 0;
}

int launch_texture_cast(int64_t n_params, const float *params, void *params_half, cudaStream_t s)
{
    if (n_params <= 0) return 0;
    const int64_t n2 = n_params / 2;
    tex_cast_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(n2, reinterpret_cast<const float2 *>(params),
                                                                 reinterpret_cast<__half2 *>(params_half));
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

int launch_texture_fwd(int64_t N, int C, const float *aabb6_host, const float *xyz, const void *grid_half, const float *W0,
                       const float *W1, const float *W2, float *out, void *enc_out, cudaStream_t s)
{
    TexArgs a;
    int rc = fill_args(a, N, C, aabb6_host);
    if (rc) return rc;
    if (N == 0) return 0;
    texture_fwd_kernel<<<(unsigned)((N + TX_THREADS - 1) / TX_THREADS), TX_THREADS, 0, s>>>(
        a, xyz, reinterpret_cast<const __half2 *>(grid_half), W0, W1, W2, out, reinterpret_cast<__half *>(enc_out));
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

int launch_texture_bwd(int64_t N, int C, const float *aabb6_host, const float *xyz, const void *grid_half, const void *enc,
                       const float *W0, const float *W1, const float *W2, const float *dL_dout, float grid_grad_scale,
                       float *d_grid, float *dW0, float *dW1, float *dW2, float *d_xyz, cudaStream_t s)
{
    TexArgs a;
    int rc = fill_args(a, N, C, aabb6_host);
    if (rc) return rc;
    if (N == 0) return 0;
    const size_t smem = sizeof(TexBwdSmem) + (size_t)(2 * TX_HID * TX_IN + TX_MAX_OUT * TX_HID) * sizeof(float);
    if (once_per_device(ONCE_TEXTURE_BWD))
        DMGS_CUDA(cudaFuncSetAttribute(texture_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    texture_bwd_kernel<<<(unsigned)((N + TX_THREADS - 1) / TX_THREADS), TX_THREADS, smem, s>>>(
        a, xyz, reinterpret_cast<const __half2 *>(grid_half), reinterpret_cast<const __half *>(enc), W0, W1, W2, dL_dout,
        grid_grad_scale, d_grid, dW0, dW1, dW2, d_xyz);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
