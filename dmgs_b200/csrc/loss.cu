// loss.cu -- fused L1 + SSIM image loss, forward and backward.
//
// Replaces utils/loss_utils.py:17-18 (l1_loss) and :35-63 (ssim/_ssim: five grouped 11x11 Gaussian
// convolutions + ~15 element-wise passes over [3,H,W], and as many again in autograd's backward) as
// used by every trainer right after the render (train_geo_stage2.py:115-116).
//
//   forward  (l1_ssim_fwd_kernel): one CTA per 32x32 tile of one image plane.  The 42x42 patches of
//            x (render) and y (ground truth) are staged once in shared memory (zero padding outside
//            the image, like conv2d(padding=5)); a horizontal then a vertical 11-tap pass give the five
//            windowed moments E[x], E[y], E[xx], E[yy], E[xy]; the SSIM map value and its three partial
//            derivatives are formed in registers.  The CTA writes its |x-y| and SSIM partial sums to
//            its own slot (no atomics: the final sum is a fixed-order reduction, bit-reproducible) and
//            the three derivative fields Da, Db, Dc that the backward convolves.
//   backward (l1_ssim_bwd_kernel): dSSIM/dx = G*Da + 2x (G*Db) + y (G*Dc) (G symmetric), same staging;
//            adds the L1 term sign(x-y) and scales by the upstream gradient read from device memory.
//
// HBM traffic per plane pixel: forward 8 B read + 12 B written, backward 20 B read + 4 B written:
// 44 B against >= 40 full-image passes (> 320 B) of the op-by-op PyTorch formulation.
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int LT = 32;             // output tile edge
constexpr int LR = 5;              // window radius (11 taps)
constexpr int LP = LT + 2 * LR;    // staged patch edge (42)
constexpr int LPS = 48;            // row stride of the staged patches: 16-byte aligned rows, 16 floats from column 28 fit
constexpr int LTHREADS = 256;

struct Window11 {
    float g[11];
};

// Stages NF patches (42 rows x 48 columns, zero outside the image / beyond column 41) at once: every
// thread first issues all its global loads (8 per patch), then stores them -- the loads of a CTA are in
// flight together instead of one dependent load -> store round trip per element (first profile: 85 % of
// the samples sat on the store waiting for its load).
constexpr int LSTAGE = (LP * LPS + LTHREADS - 1) / LTHREADS;  // 8 elements per thread and patch
template <int NF>
__device__ __forceinline__ void stage_patches(float *const (&dst)[NF], const float *const (&src)[NF], int H, int W, int y0,
                                               int x0)
{
    float v[NF][LSTAGE];
#pragma unroll
    for (int k = 0; k < LSTAGE; ++k) {
        const int i = threadIdx.x + k * LTHREADS;
        const int r = i / LPS, c = i - r * LPS;
        const int y = y0 + r - LR, x = x0 + c - LR;
        const bool in = i < LP * LPS && c < LP && y >= 0 && y < H && x >= 0 && x < W;
        const size_t o = in ? (size_t)y * W + x : 0;
#pragma unroll
        for (int f = 0; f < NF; ++f) v[f][k] = in ? src[f][o] : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < LSTAGE; ++k) {
        const int i = threadIdx.x + k * LTHREADS;
        if (i < LP * LPS) {
#pragma unroll
            for (int f = 0; f < NF; ++f) dst[f][i] = v[f][k];
        }
    }
}

// 16 consecutive floats of a staged row starting at a multiple of 4: four 16-byte shared loads
__device__ __forceinline__ void load16(float *v, const float *row)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = *reinterpret_cast<const float4 *>(row + 4 * q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// Register blocking (both passes): a thread produces FOUR adjacent outputs from one sliding window of
// 14 inputs, so shared-memory traffic per output drops ~2.6x against one output per thread and the
// kernels sit at the FP32-issue / shared-memory balance point instead of being LDS bound.

// sums[(plane * tiles + tile) * 2 + {0,1}] = sum |x-y|, sum ssim_map over the tile
__global__ void __launch_bounds__(LTHREADS, 4)
l1_ssim_fwd_kernel(int H, int W, int tiles_x, const __grid_constant__ Window11 win, const float *__restrict__ img,
                   const float *__restrict__ gt, float *__restrict__ deriv, float *__restrict__ sums)
{
    __shared__ __align__(16) float sx[LP * LPS], sy[LP * LPS];
    __shared__ __align__(16) float hm[5][LP * LT];
    __shared__ float red[2][LTHREADS / 32];
    const int plane = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * LT, x0 = tx * LT;
    const size_t HW = (size_t)H * W;
    const float *xp = img + plane * HW, *yp = gt + plane * HW;
    {
        float *const dst[2] = {sx, sy};
        const float *const src[2] = {xp, yp};
        stage_patches<2>(dst, src, H, W, y0, x0);
    }
    __syncthreads();
    // horizontal pass: item = (patch row, group of 4 output columns)
    for (int i = threadIdx.x; i < LP * (LT / 4); i += LTHREADS) {
        const int r = i >> 3, cg = i & 7;
        float a[16], b[16];
        load16(a, sx + r * LPS + 4 * cg);
        load16(b, sy + r * LPS + 4 * cg);
        float o[5][4];
#pragma unroll
        for (int f = 0; f < 5; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[f][j] = 0.0f;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            const float aa = a[k] * a[k], bb = b[k] * b[k], ab = a[k] * b[k];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = k - j;  // tap index of input k for output j
                if (t >= 0 && t < 11) {
                    const float g = win.g[t];
                    o[0][j] = fma_(g, a[k], o[0][j]);
                    o[1][j] = fma_(g, b[k], o[1][j]);
                    o[2][j] = fma_(g, aa, o[2][j]);
                    o[3][j] = fma_(g, bb, o[3][j]);
                    o[4][j] = fma_(g, ab, o[4][j]);
                }
            }
        }
#pragma unroll
        for (int f = 0; f < 5; ++f)
            *reinterpret_cast<float4 *>(&hm[f][r * LT + 4 * cg]) = make_float4(o[f][0], o[f][1], o[f][2], o[f][3]);
    }
    __syncthreads();
    // vertical pass: thread = (column, group of 4 output rows)
    float l1 = 0.0f, ss = 0.0f;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    {
        const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
        float o[5][4];
#pragma unroll
        for (int f = 0; f < 5; ++f)
#pragma unroll
            for (int j = 0; j < 4; ++j) o[f][j] = 0.0f;
#pragma unroll
        for (int k = 0; k < 14; ++k) {
            float v[5];
#pragma unroll
            for (int f = 0; f < 5; ++f) v[f] = hm[f][(r0 + k) * LT + c];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = k - j;
                if (t >= 0 && t < 11) {
                    const float g = win.g[t];
#pragma unroll
                    for (int f = 0; f < 5; ++f) o[f][j] = fma_(g, v[f], o[f][j]);
                }
            }
        }
        const int x = x0 + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + j, y = y0 + r;
            if (y < H && x < W) {
                const float m1 = o[0][j], m2 = o[1][j];
                const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
                const float s11 = o[2][j] - m11, s22 = o[3][j] - m22, s12 = o[4][j] - m12;
                const float num1 = 2.0f * m12 + C1, num2 = 2.0f * s12 + C2;
                const float den1 = m11 + m22 + C1, den2 = s11 + s22 + C2;
                const float inv = 1.0f / (den1 * den2);
                const float map = (num1 * num2) * inv;
                // partial derivatives of the map w.r.t. (mu1, sigma1^2, sigma12) at this pixel
                const float d_mu1 = 2.0f * m2 * num2 * inv - map * (2.0f * m1) / den1;
                const float d_s11 = -map / den2;
                const float d_s12 = 2.0f * num1 * inv;
                const size_t p = (size_t)y * W + x;
                float *d = deriv + (size_t)plane * 3 * HW;
                d[p] = d_mu1 - 2.0f * m1 * d_s11 - m2 * d_s12;  // field convolved as is
                d[HW + p] = d_s11;                               // field whose convolution is multiplied by 2x
                d[2 * HW + p] = d_s12;                           // field whose convolution is multiplied by y
                ss += map;
                l1 += fabsf(sx[(r + LR) * LPS + c + LR] - sy[(r + LR) * LPS + c + LR]);
            }
        }
    }
    // fixed-order block reduction -> the CTA's slot
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0, b = 0;
        for (int k = 0; k < LTHREADS / 32; ++k) { a += red[0][k]; b += red[1][k]; }
        float *o = sums + ((size_t)plane * gridDim.x + tile) * 2;
        o[0] = a; o[1] = b;
    }
}

// out[plane*2 + {0,1}] = mean |x-y|, mean ssim over the plane (double accumulation, fixed order)
__global__ void __launch_bounds__(256)
l1_ssim_finish_kernel(int tiles, double inv_n, const float *__restrict__ sums, float *__restrict__ out)
{
    __shared__ double red[2][8];
    const int plane = blockIdx.x;
    double a = 0, b = 0;
    for (int t = threadIdx.x; t < tiles; t += 256) {
        a += sums[((size_t)plane * tiles + t) * 2];
        b += sums[((size_t)plane * tiles + t) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0; b = 0;
        for (int k = 0; k < 8; ++k) { a += red[0][k]; b += red[1][k]; }
        out[plane * 2] = (float)(a * inv_n);
        out[plane * 2 + 1] = (float)(b * inv_n);
    }
}

// grad[p] = up_l1[plane] * sign(x - y) + up_ssim[plane] * (G*Da + 2 x G*Db + y G*Dc); up_* are device
// arrays (the upstream gradient times the loss weights, already divided by the pixel count)
__global__ void __launch_bounds__(LTHREADS, 4)
l1_ssim_bwd_kernel(int H, int W, int tiles_x, const __grid_constant__ Window11 win, const float *__restrict__ img,
                   const float *__restrict__ gt, const float *__restrict__ deriv, const float *__restrict__ up,
                   float *__restrict__ grad)
{
    __shared__ __align__(16) float sd[3][LP * LPS];
    __shared__ __align__(16) float hm[3][LP * LT];
    const int plane = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * LT, x0 = tx * LT;
    const size_t HW = (size_t)H * W;
    const float *d = deriv + (size_t)plane * 3 * HW;
    {
        float *const dst[3] = {sd[0], sd[1], sd[2]};
        const float *const src[3] = {d, d + HW, d + 2 * HW};
        stage_patches<3>(dst, src, H, W, y0, x0);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LP * (LT / 4); i += LTHREADS) {
        const int r = i >> 3, cg = i & 7;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            float a[16], o[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            load16(a, sd[f] + r * LPS + 4 * cg);
#pragma unroll
            for (int k = 0; k < 14; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = k - j;
                    if (t >= 0 && t < 11) o[j] = fma_(win.g[t], a[k], o[j]);
                }
            *reinterpret_cast<float4 *>(&hm[f][r * LT + 4 * cg]) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();
    const float up_l1 = up[plane * 2], up_ss = up[plane * 2 + 1];
    const int c = threadIdx.x & 31, r0 = (threadIdx.x >> 5) * 4;
    float o[3][4];
#pragma unroll
    for (int f = 0; f < 3; ++f)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[f][j] = 0.0f;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
        float v[3];
#pragma unroll
        for (int f = 0; f < 3; ++f) v[f] = hm[f][(r0 + k) * LT + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int t = k - j;
            if (t >= 0 && t < 11) {
                const float g = win.g[t];
#pragma unroll
                for (int f = 0; f < 3; ++f) o[f][j] = fma_(g, v[f], o[f][j]);
            }
        }
    }
    const int x = x0 + c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + r0 + j;
        if (y >= H || x >= W) continue;
        const size_t p = (size_t)plane * HW + (size_t)y * W + x;
        const float xv = img[p], yv = gt[p];
        const float df = xv - yv;
        const float sgn = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
        grad[p] = fma_(up_ss, o[0][j] + 2.0f * xv * o[1][j] + yv * o[2][j], up_l1 * sgn);
    }
}

static size_t loss_tiles(int H, int W) { return (size_t)((H + LT - 1) / LT) * ((W + LT - 1) / LT); }

size_t loss_scratch_bytes(int planes, int H, int W)
{
    return align_up((size_t)planes * 3 * H * W * sizeof(float)) + align_up((size_t)planes * loss_tiles(H, W) * 2 * sizeof(float));
}

int launch_l1_ssim_fwd(int planes, int H, int W, const float *window11, const float *img, const float *gt,
                       void *scratch, float *out_means, cudaStream_t s)
{
    Window11 win;
    for (int i = 0; i < 11; ++i) win.g[i] = window11[i];
    const int tiles = (int)loss_tiles(H, W), tiles_x = (W + LT - 1) / LT;
    float *deriv = reinterpret_cast<float *>(scratch);
    float *sums = reinterpret_cast<float *>(reinterpret_cast<char *>(scratch) + align_up((size_t)planes * 3 * H * W * sizeof(float)));
    l1_ssim_fwd_kernel<<<dim3(tiles, planes), LTHREADS, 0, s>>>(H, W, tiles_x, win, img, gt, deriv, sums);
    l1_ssim_finish_kernel<<<planes, 256, 0, s>>>(tiles, 1.0 / ((double)H * W), sums, out_means);
    DMGS_CUDA(cudaGetLastError());
    count_launches(2);
    return 0;
}

int launch_l1_ssim_bwd(int planes, int H, int W, const float *window11, const float *img, const float *gt,
                       const void *scratch, const float *upstream, float *grad, cudaStream_t s)
{
    Window11 win;
    for (int i = 0; i < 11; ++i) win.g[i] = window11[i];
    const int tiles = (int)loss_tiles(H, W), tiles_x = (W + LT - 1) / LT;
    l1_ssim_bwd_kernel<<<dim3(tiles, planes), LTHREADS, 0, s>>>(H, W, tiles_x, win, img, gt,
                                                                 reinterpret_cast<const float *>(scratch), upstream, grad);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
