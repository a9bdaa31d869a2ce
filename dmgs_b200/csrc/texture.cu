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
// weights broadcast from shared memory.  Backward: two kernels, described at their definitions below (the MLP's
// adjoint chain + weight gradients over tiles of 128 points; the table scatter with sixteen lanes per point).
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

// (Weights in constant memory were measured and rejected: the compiler turns every weight into an LDC, and the kernels
// ran 1.4x (forward) to 3x (backward) slower than with broadcast shared-memory loads, profiles/r2_s3_texture.md.)
// two table entries (x, x+1 corners of one (y, z) pair).  Neighbours in x are neighbours in the table when the level is
// dense and for every even x when it is hashed (the x term of the hash is x itself): an aligned pair is ONE 8-byte load.
__device__ __forceinline__ void tex_pair(const __half2 *__restrict__ lev, uint32_t i0, uint32_t i1, float2 &v0, float2 &v1)
{
    if ((i0 ^ i1) == 1u) {
        const uint2 raw = *reinterpret_cast<const uint2 *>(lev + (i0 & ~1u));
        const uint32_t r0 = (i0 & 1u) ? raw.y : raw.x, r1 = (i0 & 1u) ? raw.x : raw.y;
        v0 = __half22float2(*reinterpret_cast<const __half2 *>(&r0));
        v1 = __half22float2(*reinterpret_cast<const __half2 *>(&r1));
    } else {
        v0 = __half22float2(lev[i0]);
        v1 = __half22float2(lev[i1]);
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
    for (int yz = 0; yz < 4; ++yz) {
        const float wy = (yz & 1) ? w[1] : 1.0f - w[1], wz = (yz & 2) ? w[2] : 1.0f - w[2];
        float2 v0, v1;
        tex_pair(lev, tex_index(g[0], g[1] + (yz & 1), g[2] + (yz >> 1), res, size),
                 tex_index(g[0] + 1, g[1] + (yz & 1), g[2] + (yz >> 1), res, size), v0, v1);
        const float wgt0 = (1.0f - w[0]) * wy * wz, wgt1 = w[0] * wy * wz;
        acc.x = fma_(wgt0, v0.x, acc.x);
        acc.y = fma_(wgt0, v0.y, acc.y);
        acc.x = fma_(wgt1, v1.x, acc.x);
        acc.y = fma_(wgt1, v1.y, acc.y);
    }
    return acc;
}

__global__ void __launch_bounds__(256) tex_cast_kernel(int64_t n2, const float2 *__restrict__ src, __half2 *__restrict__ dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) dst[i] = __float22half2_rn(src[i]);
}

// y[j] = sum_i W[j][i] x[i] for j < 32; W row-major [32][32] in shared memory (broadcast 16-byte loads)
template <bool RELU>
__device__ __forceinline__ void tex_layer(const float *__restrict__ sW, const float (&x)[32], float (&y)[32])
{
#pragma unroll
    for (int j = 0; j < 32; ++j) {
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

__global__ void __launch_bounds__(TX_THREADS, 4)
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
    tex_layer<true>(sW0, x, h1);
    tex_layer<true>(sW1, h1, h2);
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
// Two kernels.  (A) texture_mlp_bwd_kernel: the MLP's adjoint chain and the three weight gradients.  A CTA walks tiles
// of 128 points (grid-stride, so the weight-gradient accumulators stay in registers across tiles and are flushed once
// per CTA).  Per tile, thread <-> point recomputes the activations and runs the adjoint chain layer by layer; for
// every layer the adjoints go to a shared panel [row][point] and the activations to a panel
// [point][feature], and the four warps contract the tile's 128 points into dW: a lane owns a 2 x 4 block of the
// layer's 32 x 32 entries (rows 8 warp + 2 (lane / 8), columns 4 (lane % 8)), so four points cost six 16-byte shared
// loads for 32 FMAs -- the contraction is bound by the shared-memory load path (one cycle per operand register and
// warp), and the block shape is what minimises operands per FMA at 8 accumulators per thread.  Row strides of 132 / 36
// floats keep every access conflict free.  The panels are reused by the three layers (52 KB: three CTAs per SM).
// dL/d(encoding) leaves as fp32 [N][32].
// (B) texture_grid_bwd_kernel: lane <-> (point, level), sixteen lanes per point: vector reductions into the gradient
// table (x-neighbours that share an aligned 16-byte slot go as one red.v4), the gathers of the position gradient, a
// 16-lane butterfly for dL/dxyz.
constexpr int TXB_PTS = 128, TXB_ASTR = 132, TXB_BSTR = 36;
struct TexBwdSmem {
    float a[TX_MAX_OUT][TXB_ASTR];  // adjoint panel [row][point]
    float b[TXB_PTS][TXB_BSTR];     // activation panel [point][feature]
    float W0[TX_HID * TX_IN], W1[TX_HID * TX_HID], W2[TX_MAX_OUT * TX_HID];
};

__device__ __forceinline__ void red_add_v2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// acc[r][c] += sum over the tile's points of a[row0 + r][p] * b[p][col0 + c]
__device__ __forceinline__ void tex_contract(const TexBwdSmem &S, int row0, int col0, float (&acc)[2][4])
{
#pragma unroll 2
    for (int k = 0; k < TXB_PTS; k += 4) {
        const float4 a0 = *reinterpret_cast<const float4 *>(&S.a[row0][k]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&S.a[row0 + 1][k]);
        const float av[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 bv = *reinterpret_cast<const float4 *>(&S.b[k + q][col0]);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                acc[r][0] = fma_(av[r][q], bv.x, acc[r][0]); acc[r][1] = fma_(av[r][q], bv.y, acc[r][1]);
                acc[r][2] = fma_(av[r][q], bv.z, acc[r][2]); acc[r][3] = fma_(av[r][q], bv.w, acc[r][3]);
            }
        }
    }
}

// d[i] = sum_j W[j][i] * v[j]  (transposed product; W row-major [32][32] in shared memory)
__device__ __forceinline__ void tex_layer_t(const float *__restrict__ sW, const float (&v)[32], float (&d)[32])
{
#pragma unroll
    for (int i = 0; i < 32; ++i) d[i] = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(sW + j * 32 + i);
            d[i] = fma_(wv.x, v[j], d[i]); d[i + 1] = fma_(wv.y, v[j], d[i + 1]);
            d[i + 2] = fma_(wv.z, v[j], d[i + 2]); d[i + 3] = fma_(wv.w, v[j], d[i + 3]);
        }
    }
}

__device__ __forceinline__ void tex_store_b(TexBwdSmem &S, int p, const float (&v)[32])
{
#pragma unroll
    for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(&S.b[p][i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

__global__ void __launch_bounds__(TXB_PTS, 3)
texture_mlp_bwd_kernel(int64_t N, int C, const __half *__restrict__ enc, const float *__restrict__ W0,
                       const float *__restrict__ W1, const float *__restrict__ W2, const float *__restrict__ dL_dout,
                       float *__restrict__ d_enc, float *__restrict__ dW0, float *__restrict__ dW1, float *__restrict__ dW2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TexBwdSmem &S = *reinterpret_cast<TexBwdSmem *>(smem_raw);
    for (int i = threadIdx.x; i < TX_HID * TX_IN; i += TXB_PTS) { S.W0[i] = W0[i]; S.W1[i] = W1[i]; }
    for (int i = threadIdx.x; i < C * TX_HID; i += TXB_PTS) S.W2[i] = W2[i];
    const int p = threadIdx.x, lane = p & 31, warp = p >> 5;
    const int row0 = 8 * warp + 2 * (lane >> 3), col0 = 4 * (lane & 7);
    // this lane's 2 x 4 blocks: dW2 rows row0 (+32 in the second pass, channels 32..63), dW1, dW0
    float acc2[2][2][4], acc1[2][4], acc0[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { acc2[0][r][c] = 0.0f; acc2[1][r][c] = 0.0f; acc1[r][c] = 0.0f; acc0[r][c] = 0.0f; }
    // rows of the adjoint panel beyond C stay zero (the second dW2 pass reads whole 8-row groups)
    for (int i = p; i < TX_MAX_OUT * TXB_ASTR; i += TXB_PTS) (&S.a[0][0])[i] = 0.0f;
    const int64_t tiles = (N + TXB_PTS - 1) / TXB_PTS;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t n = tile * TXB_PTS + p;
        const bool live = n < N;
        float x[TX_IN], h1[TX_HID];
#pragma unroll
        for (int l = 0; l < TX_LEVELS; l += 4) {  // 64 bytes of fp16 features: four 16-byte loads
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (live) raw = *reinterpret_cast<const uint4 *>(enc + n * TX_IN + 2 * l);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 r = __half22float2(*reinterpret_cast<const __half2 *>(&w[q]));
                x[2 * (l + q)] = r.x; x[2 * (l + q) + 1] = r.y;
            }
        }
        tex_layer<true>(S.W0, x, h1);
        float d2[TX_HID];
        {
            float h2[TX_HID];
            tex_layer<true>(S.W1, h1, h2);
            // ---- layer 2: dout -> adjoint panel, h2 -> activation panel, dh2 = relu'(h2) W2^T dout
            tex_store_b(S, p, h2);
#pragma unroll
            for (int j = 0; j < TX_HID; ++j) d2[j] = 0.0f;
            for (int c0 = 0; c0 < C; c0 += 4) {
                float4 g4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (live) g4 = *reinterpret_cast<const float4 *>(dL_dout + n * C + c0);
                const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    S.a[c0 + q][p] = g[q];
#pragma unroll
                    for (int j = 0; j < TX_HID; j += 4) {
                        const float4 wv = *reinterpret_cast<const float4 *>(S.W2 + (c0 + q) * 32 + j);
                        d2[j] = fma_(wv.x, g[q], d2[j]); d2[j + 1] = fma_(wv.y, g[q], d2[j + 1]);
                        d2[j + 2] = fma_(wv.z, g[q], d2[j + 2]); d2[j + 3] = fma_(wv.w, g[q], d2[j + 3]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < TX_HID; ++j) d2[j] = h2[j] > 0.0f ? d2[j] : 0.0f;
        }
        __syncthreads();
        tex_contract(S, row0, col0, acc2[0]);
        if (C > 32 && row0 + 32 < ((C + 7) & ~7)) tex_contract(S, row0 + 32, col0, acc2[1]);
        __syncthreads();
        // ---- layer 1: dh2 -> adjoint panel, h1 -> activation panel, dh1 = relu'(h1) W1^T dh2
        float d1[TX_HID];
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) S.a[j][p] = d2[j];
        tex_store_b(S, p, h1);
        tex_layer_t(S.W1, d2, d1);
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) d1[j] = h1[j] > 0.0f ? d1[j] : 0.0f;
        __syncthreads();
        tex_contract(S, row0, col0, acc1);
        __syncthreads();
        // ---- layer 0: dh1 -> adjoint panel, x -> activation panel, dx = W0^T dh1
#pragma unroll
        for (int j = 0; j < TX_HID; ++j) S.a[j][p] = d1[j];
        tex_store_b(S, p, x);
        float dx[TX_IN];
        tex_layer_t(S.W0, d1, dx);
        if (live) {
#pragma unroll
            for (int i = 0; i < TX_IN; i += 4)
                *reinterpret_cast<float4 *>(d_enc + n * TX_IN + i) = make_float4(dx[i], dx[i + 1], dx[i + 2], dx[i + 3]);
        }
        __syncthreads();
        tex_contract(S, row0, col0, acc0);
        __syncthreads();
    }
    // one reduction per weight entry and CTA
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (row0 + r < C && acc2[0][r][c] != 0.0f) atomicAdd(dW2 + (row0 + r) * TX_HID + col0 + c, acc2[0][r][c]);
            if (row0 + 32 + r < C && acc2[1][r][c] != 0.0f) atomicAdd(dW2 + (row0 + 32 + r) * TX_HID + col0 + c, acc2[1][r][c]);
            if (acc1[r][c] != 0.0f) atomicAdd(dW1 + (row0 + r) * TX_HID + col0 + c, acc1[r][c]);
            if (acc0[r][c] != 0.0f) atomicAdd(dW0 + (row0 + r) * TX_IN + col0 + c, acc0[r][c]);
        }
    }
}

__global__ void __launch_bounds__(256)
texture_grid_bwd_kernel(const __grid_constant__ TexArgs a, const float *__restrict__ xyz, const __half2 *__restrict__ grid,
                        const float *__restrict__ d_enc, float grid_grad_scale, float *__restrict__ d_grid,
                        float *__restrict__ d_xyz)
{
    const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int l = (int)(t & 15);
    const int64_t n = t >> 4;
    const bool live = n < a.N;
    float dt[3] = {0.0f, 0.0f, 0.0f};
    bool inside[3] = {false, false, false};
    if (live) {
        float tc[3];
        tex_coords(a, xyz, n, tc, inside);
        const float2 gxy = *reinterpret_cast<const float2 *>(d_enc + n * TX_IN + 2 * l);
        const float gx = gxy.x, gy = gxy.y;
        if (gx != 0.0f || gy != 0.0f) {
            const float s = a.lv.scale[l];
            float w[3];
            uint32_t g[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float pos = fma_(s, tc[d], 0.5f);
                const float fl = floorf(pos);
                w[d] = pos - fl;
                g[d] = (uint32_t)fl;
            }
            const uint32_t res = a.lv.res[l], size = a.lv.size[l];
            const __half2 *lev = grid + a.lv.offset[l];
            float *dlev = d_grid ? d_grid + 2 * (size_t)a.lv.offset[l] : nullptr;
            float dw[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int yz = 0; yz < 4; ++yz) {
                const float wy = (yz & 1) ? w[1] : 1.0f - w[1], wz = (yz & 2) ? w[2] : 1.0f - w[2];
                const uint32_t i0 = tex_index(g[0], g[1] + (yz & 1), g[2] + (yz >> 1), res, size);
                const uint32_t i1 = tex_index(g[0] + 1, g[1] + (yz & 1), g[2] + (yz >> 1), res, size);
                if (dlev) {
                    const float wg0 = (1.0f - w[0]) * wy * wz * grid_grad_scale, wg1 = w[0] * wy * wz * grid_grad_scale;
                    if ((i0 ^ i1) == 1u) {  // the two entries share an aligned 16-byte slot
                        if (i0 & 1u) red_add_v4(dlev + 2 * (size_t)i1, wg1 * gx, wg1 * gy, wg0 * gx, wg0 * gy);
                        else red_add_v4(dlev + 2 * (size_t)i0, wg0 * gx, wg0 * gy, wg1 * gx, wg1 * gy);
                    } else {
                        red_add_v2(dlev + 2 * (size_t)i0, wg0 * gx, wg0 * gy);
                        red_add_v2(dlev + 2 * (size_t)i1, wg1 * gx, wg1 * gy);
                    }
                }
                if (d_xyz) {
                    float2 v0, v1;
                    tex_pair(lev, i0, i1, v0, v1);
                    const float gv0 = fma_(v0.x, gx, v0.y * gy), gv1 = fma_(v1.x, gx, v1.y * gy);
                    const float wx0 = 1.0f - w[0], wx1 = w[0];
                    dw[0] += wy * wz * (gv1 - gv0);
                    dw[1] += ((yz & 1) ? 1.0f : -1.0f) * wz * fma_(wx0, gv0, wx1 * gv1);
                    dw[2] += ((yz & 2) ? 1.0f : -1.0f) * wy * fma_(wx0, gv0, wx1 * gv1);
                }
            }
#pragma unroll
            for (int d = 0; d < 3; ++d) dt[d] = dw[d] * s;
        }
    }
    if (d_xyz) {  // the sixteen levels of a point sit in sixteen consecutive lanes
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int d = 0; d < 3; ++d) dt[d] += __shfl_xor_sync(0xffffffffu, dt[d], o);
        }
        if (live && l == 0) {
#pragma unroll
            for (int d = 0; d < 3; ++d) d_xyz[3 * n + d] = inside[d] ? dt[d] / (a.hi[d] - a.lo[d]) : 0.0f;
        }
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

size_t texture_bwd_scratch_bytes(int64_t N) { return (size_t)(N > 0 ? N : 0) * TX_IN * sizeof(float); }

int launch_texture_bwd(int64_t N, int C, const float *aabb6_host, const float *xyz, const void *grid_half, const void *enc,
                       const float *W0, const float *W1, const float *W2, const float *dL_dout, float grid_grad_scale,
                       float *d_grid, float *dW0, float *dW1, float *dW2, float *d_xyz, void *scratch, cudaStream_t s)
{
    TexArgs a;
    int rc = fill_args(a, N, C, aabb6_host);
    if (rc) return rc;
    if (N == 0) return 0;
    float *d_enc = reinterpret_cast<float *>(scratch);
    const size_t smem = sizeof(TexBwdSmem);
    if (once_per_device(ONCE_TEXTURE_BWD))
        DMGS_CUDA(cudaFuncSetAttribute(texture_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (N + TXB_PTS - 1) / TXB_PTS;
    const int64_t resident = 3 * (int64_t)num_sms();
    const unsigned blocks = (unsigned)(tiles < resident ? tiles : resident);
    const __half *e = reinterpret_cast<const __half *>(enc);
    texture_mlp_bwd_kernel<<<blocks, TXB_PTS, smem, s>>>(N, C, e, W0, W1, W2, dL_dout, d_enc, dW0, dW1, dW2);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    if (d_grid || d_xyz) {
        texture_grid_bwd_kernel<<<(unsigned)((N * TX_LEVELS + 255) / 256), 256, 0, s>>>(
            a, xyz, reinterpret_cast<const __half2 *>(grid_half), d_enc, grid_grad_scale, d_grid, d_xyz);
        DMGS_CUDA(cudaGetLastError());
        count_launches(1);
    }
    return 0;
}

}  // namespace dmgs
