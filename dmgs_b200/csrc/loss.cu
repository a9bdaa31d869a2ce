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
constexpr int LPS = LP + 1;        // padded row stride of the staged patches
constexpr int LTHREADS = 256;

struct Window11 {
    float g[11];
};

__device__ __forceinline__ void stage_patch(float *dst, const float *__restrict__ src, int H, int W, int y0, int x0)
{
    for (int i = threadIdx.x; i < LP * LP; i += LTHREADS) {
        const int r = i / LP, c = i - r * LP;
        const int y = y0 + r - LR, x = x0 + c - LR;
        dst[r * LPS + c] = (y >= 0 && y < H && x >= 0 && x < W) ? src[(size_t)y * W + x] : 0.0f;
    }
}

// sums[(plane * tiles + tile) * 2 + {0,1}] = sum |x-y|, sum ssim_map over the tile
__global__ void __launch_bounds__(LTHREADS)
l1_ssim_fwd_kernel(int H, int W, int tiles_x, const __grid_constant__ Window11 win, const float *__restrict__ img,
                   const float *__restrict__ gt, float *__restrict__ deriv, float *__restrict__ sums)
{
    __shared__ float sx[LP * LPS], sy[LP * LPS];
    __shared__ float hm[5][LP * LT];
    __shared__ float red[2][LTHREADS / 32];
    const int plane = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * LT, x0 = tx * LT;
    const size_t HW = (size_t)H * W;
    const float *xp = img + plane * HW, *yp = gt + plane * HW;
    stage_patch(sx, xp, H, W, y0, x0);
    stage_patch(sy, yp, H, W, y0, x0);
    __syncthreads();
    // horizontal pass: rows of the patch x output columns
    for (int i = threadIdx.x; i < LP * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        const float *rx = sx + r * LPS + c, *ry = sy + r * LPS + c;
        float m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float a = rx[k], b = ry[k], g = win.g[k];
            m1 = fma_(g, a, m1);
            m2 = fma_(g, b, m2);
            e11 = fma_(g, a * a, e11);
            e22 = fma_(g, b * b, e22);
            e12 = fma_(g, a * b, e12);
        }
        hm[0][i] = m1; hm[1][i] = m2; hm[2][i] = e11; hm[3][i] = e22; hm[4][i] = e12;
    }
    __syncthreads();
    float l1 = 0.0f, ss = 0.0f;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    for (int i = threadIdx.x; i < LT * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        const int y = y0 + r, x = x0 + c;
        float m1 = 0, m2 = 0, e11 = 0, e22 = 0, e12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const int j = (r + k) * LT + c;
            const float g = win.g[k];
            m1 = fma_(g, hm[0][j], m1);
            m2 = fma_(g, hm[1][j], m2);
            e11 = fma_(g, hm[2][j], e11);
            e22 = fma_(g, hm[3][j], e22);
            e12 = fma_(g, hm[4][j], e12);
        }
        if (y < H && x < W) {
            const float m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            const float s11 = e11 - m11, s22 = e22 - m22, s12 = e12 - m12;
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
__global__ void __launch_bounds__(LTHREADS)
l1_ssim_bwd_kernel(int H, int W, int tiles_x, const __grid_constant__ Window11 win, const float *__restrict__ img,
                   const float *__restrict__ gt, const float *__restrict__ deriv, const float *__restrict__ up,
                   float *__restrict__ grad)
{
    __shared__ float sd[3][LP * LPS];
    __shared__ float hm[3][LP * LT];
    const int plane = blockIdx.y, tile = blockIdx.x;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const int y0 = ty * LT, x0 = tx * LT;
    const size_t HW = (size_t)H * W;
    const float *d = deriv + (size_t)plane * 3 * HW;
    for (int f = 0; f < 3; ++f) stage_patch(sd[f], d + f * HW, H, W, y0, x0);
    __syncthreads();
    for (int i = threadIdx.x; i < LP * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        float a = 0, b = 0, cc = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float g = win.g[k];
            a = fma_(g, sd[0][r * LPS + c + k], a);
            b = fma_(g, sd[1][r * LPS + c + k], b);
            cc = fma_(g, sd[2][r * LPS + c + k], cc);
        }
        hm[0][i] = a; hm[1][i] = b; hm[2][i] = cc;
    }
    __syncthreads();
    const float up_l1 = up[plane * 2], up_ss = up[plane * 2 + 1];
    for (int i = threadIdx.x; i < LT * LT; i += LTHREADS) {
        const int r = i / LT, c = i - r * LT;
        const int y = y0 + r, x = x0 + c;
        if (y >= H || x >= W) continue;
        float a = 0, b = 0, cc = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const int j = (r + k) * LT + c;
            const float g = win.g[k];
            a = fma_(g, hm[0][j], a);
            b = fma_(g, hm[1][j], b);
            cc = fma_(g, hm[2][j], cc);
        }
        const size_t p = (size_t)plane * HW + (size_t)y * W + x;
        const float xv = img[p], yv = gt[p];
        const float df = xv - yv;
        const float sgn = df > 0.0f ? 1.0f : (df < 0.0f ? -1.0f : 0.0f);
        grad[p] = fma_(up_ss, a + 2.0f * xv * b + yv * cc, up_l1 * sgn);
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
