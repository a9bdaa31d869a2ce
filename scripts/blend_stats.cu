// blend_stats.cu -- DIAGNOSTIC (not part of libdmgs_raster.so): counts what the blend kernels' warp-rectangle
// culling does on a rendered frame, for candidate warp rectangles 8x4 (what the kernels use), 8x8 and 16x16.
// Reads the state buffers of a finished forward (ranges, sorted indices, 32-byte records, n_contrib).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -shared -Xcompiler -fPIC \
//        -o scripts/_blend_stats.so scripts/blend_stats.cu        (scripts/blend_stats.py builds and drives it)
// Counters per rectangle shape s (stride 8):
//   [0] (rect, entry) pairs in range for the backward (entry index < max n_contrib of the rectangle)
//   [1] ... that survive cull_rect            [2] ... that survive and have >= 1 contributing pixel
//   [3] contributing (pixel, entry) pairs     [4] pairs with >= 1 contributing pixel (cull ignored; == [2] if the
//   cull is conservative)                     [5] rectangles with any work   [6] 32-entry groups visited
#include "../dmgs_b200/csrc/blend_common.cuh"

using namespace dmgs;

__global__ void __launch_bounds__(256) blend_stats_kernel(BlendArgs a, const uint2 *ranges, const uint32_t *gidx,
                                                          const float4 *rec, const uint32_t *n_contrib,
                                                          unsigned long long *out)
{
    __shared__ uint32_t s_hit[32][8];   // ballot of contributing pixels per staged entry and 8x4 rectangle
    __shared__ float4 s_ra[32], s_rb[32];
    __shared__ int s_wmax[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int px0 = blockIdx.x * DMGS_TILE + (w & 1) * 8, py0 = blockIdx.y * DMGS_TILE + (w >> 1) * 4;
    const int px = px0 + (lane & 7), py = py0 + (lane >> 3);
    const bool inside = px < a.W && py < a.H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 rng = ranges[blockIdx.y * a.gx + blockIdx.x];
    const int last = inside ? (int)n_contrib[(size_t)py * a.W + px] : 0;
    const int wmax = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) s_wmax[w] = wmax;
    __syncthreads();
    int count = 0;
    for (int i = 0; i < 8; ++i) count = max(count, s_wmax[i]);
    unsigned long long c[3][7];
    for (int s = 0; s < 3; ++s)
        for (int k = 0; k < 7; ++k) c[s][k] = 0;
    for (int base = 0; base < count; base += 32) {
        if (threadIdx.x < 32 && base + threadIdx.x < count) {
            const uint32_t g = gidx[rng.x + base + threadIdx.x];
            s_ra[threadIdx.x] = rec[2 * (size_t)g];
            s_rb[threadIdx.x] = rec[2 * (size_t)g + 1];
        }
        __syncthreads();
        const int nb = min(32, count - base);
        for (int e = 0; e < nb; ++e) {
            const float4 ra = s_ra[e], rb = s_rb[e];
            const float dx = ra.x - pxf, dy = ra.y - pyf;
            const float t1 = ra.z * dx, t2 = (rb.x * dy) * dy, t3 = ((-ra.w) * dx) * dy;
            const float power = fma_(fma_(dx, t1, t2), -0.5f, t3);
            const float alpha = fminf(0.99f, rb.y * dmgs_exp(power));
            const bool hit = base + e < last && power <= 0.0f && alpha >= 1.0f / 255.0f;
            const uint32_t m = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_hit[e][w] = m;
        }
        __syncthreads();
        if (threadIdx.x < nb) {
            const int e = threadIdx.x, idx = base + e;
            const float4 ra = s_ra[e], rb = s_rb[e];
            const float tx = (float)(blockIdx.x * DMGS_TILE), ty = (float)(blockIdx.y * DMGS_TILE);
            // shape 0: 8x4 (warp w: x block w&1, y block w>>1)
            for (int q = 0; q < 8; ++q) {
                if (idx >= s_wmax[q]) continue;
                const float x0 = tx + (q & 1) * 8, y0 = ty + (q >> 1) * 4;
                const bool keep = !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, x0, x0 + 7, y0, y0 + 3);
                const uint32_t m = s_hit[e][q];
                c[0][0]++; c[0][1] += keep; c[0][2] += keep && m; c[0][3] += __popc(m); c[0][4] += m != 0;
            }
            // shape 1: 8x8 (x block q&1, y block q>>1 = warps q&1 + 4*(q>>1) and +2)
            for (int q = 0; q < 4; ++q) {
                const int w0 = (q & 1) + 4 * (q >> 1), w1 = w0 + 2;
                if (idx >= max(s_wmax[w0], s_wmax[w1])) continue;
                const float x0 = tx + (q & 1) * 8, y0 = ty + (q >> 1) * 8;
                const bool keep = !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, x0, x0 + 7, y0, y0 + 7);
                const uint32_t m0 = s_hit[e][w0], m1 = s_hit[e][w1];
                c[1][0]++; c[1][1] += keep; c[1][2] += keep && (m0 | m1); c[1][3] += __popc(m0) + __popc(m1);
                c[1][4] += (m0 | m1) != 0;
            }
            // shape 2: the whole 16x16 tile
            {
                const bool keep = !cull_rect(ra.x, ra.y, ra.z, ra.w, rb.x, rb.z, tx, tx + 15, ty, ty + 15);
                uint32_t any = 0, n = 0;
                for (int q = 0; q < 8; ++q) { any |= s_hit[e][q]; n += __popc(s_hit[e][q]); }
                c[2][0]++; c[2][1] += keep; c[2][2] += keep && any; c[2][3] += n; c[2][4] += any != 0;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        for (int q = 0; q < 8; ++q) { c[0][5] += s_wmax[q] > 0; c[0][6] += (s_wmax[q] + 31) / 32; }
        for (int q = 0; q < 4; ++q) {
            const int w0 = (q & 1) + 4 * (q >> 1), m = max(s_wmax[w0], s_wmax[w0 + 2]);
            c[1][5] += m > 0; c[1][6] += (m + 31) / 32;
        }
        c[2][5] += count > 0; c[2][6] += (count + 31) / 32;
    }
    if (threadIdx.x < 32)
        for (int s = 0; s < 3; ++s)
            for (int k = 0; k < 7; ++k) {
                unsigned long long v = c[s][k];
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0 && v) atomicAdd(&out[s * 8 + k], v);
            }
}

extern "C" int blend_stats(int W, int H, const void *ranges, const void *gidx, const void *rec, const void *n_contrib,
                           void *out24, void *stream)
{
    BlendArgs a;
    a.W = W; a.H = H;
    a.gx = (W + DMGS_TILE - 1) / DMGS_TILE; a.gy = (H + DMGS_TILE - 1) / DMGS_TILE;
    a.bg[0] = a.bg[1] = a.bg[2] = 0.0f;
    blend_stats_kernel<<<dim3(a.gx, a.gy), 256, 0, (cudaStream_t)stream>>>(a, (const uint2 *)ranges, (const uint32_t *)gidx,
                                                                           (const float4 *)rec, (const uint32_t *)n_contrib,
                                                                           (unsigned long long *)out24);
    return (int)cudaGetLastError();
}
