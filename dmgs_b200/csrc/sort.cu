// sort.cu -- in-house stable radix machinery, prefix scan, instance emission, tile ranges.
//
// Replaces cub::DeviceScan + duplicateWithKeys + cub::DeviceRadixSort(64-bit keys) +
// identifyTileRanges with a different decomposition:
//   1. the P Gaussians are sorted by depth bits (4 x 8-bit stable LSD passes over P items);
//   2. tiles-touched is scanned in depth order and the (tile, gaussian) instances are emitted in
//      that order, so every tile's sub-sequence is already depth-sorted with ties in ascending
//      Gaussian index -- exactly the order a stable sort of (tile<<32 | depth) keys produces;
//   3. the R instances are stably partitioned by tile id: ceil(log2 T) bits in two LSD passes.
// R-sized traffic is 2 passes x 20 B instead of 6 passes x 24 B + histogram for a 64-bit sort.
// All passes are spin-free (histogram -> scan -> ranked scatter), so nothing can dead-lock.
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 2048

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a 256-thread block; returns block total in *total
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) {
        const uint32_t s = warp_sums[k];
        if (k < w) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const uint32_t *__restrict__ in, const uint32_t *__restrict__ gather, int64_t n,
                   uint32_t *__restrict__ block_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) s += gather ? in[gather[i]] : in[i];
    }
    uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_spine_kernel(uint32_t *__restrict__ block_sums, int nb, uint32_t *__restrict__ total)
{
    uint32_t carry = 0;
    for (int b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        const int i = b0 + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0;
        uint32_t tot;
        const uint32_t ex = block_excl_scan(v, &tot);
        if (i < nb) block_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t *in, const uint32_t *__restrict__ gather, uint32_t *out,
                  int64_t n, const uint32_t *__restrict__ block_sums)
{
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        v[k] = i < n ? (gather ? in[gather[i]] : in[i]) : 0;
        s += v[k];
    }
    uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n) out[i] = ex;
        ex += v[k];
    }
}

int exclusive_scan_u32(const uint32_t *in, const uint32_t *gather, uint32_t *out, int64_t n, uint32_t *total,
                       uint32_t *scan_tmp, cudaStream_t s)
{
    if (n <= 0) {
        if (total) DMGS_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), s));
        return 0;
    }
    const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, gather, n, scan_tmp);
    scan_spine_kernel<<<1, SCAN_THREADS, 0, s>>>(scan_tmp, nb, total);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, gather, out, n, scan_tmp);
    DMGS_CUDA(cudaGetLastError());
    count_launches(3);
    return 0;
}

// ------------------------------------------------------------------------------ radix pass
constexpr int RP_THREADS = 256;
constexpr int RP_WARPS = RP_THREADS / 32;
constexpr int RP_ROUNDS = SORT_ITEMS_PER_BLOCK / RP_THREADS;  // 32 rounds of 32 items per warp

__global__ void __launch_bounds__(RP_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift, int bins, int nblocks,
                  uint32_t *__restrict__ hist)
{
    __shared__ uint32_t h[SORT_MAX_BINS];
    for (int d = threadIdx.x; d < bins; d += RP_THREADS) h[d] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_ITEMS_PER_BLOCK;
    const uint32_t mask = (uint32_t)bins - 1;
#pragma unroll 4
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RP_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < bins; d += RP_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = h[d];
}

__global__ void __launch_bounds__(RP_THREADS)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n, int shift, int bins,
                     int nblocks, const uint32_t *__restrict__ hist_scanned)
{
    __shared__ uint32_t wh[RP_WARPS][SORT_MAX_BINS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RP_WARPS * SORT_MAX_BINS; d += RP_THREADS) (&wh[0][0])[d] = 0;
    __syncthreads();
    const uint32_t mask = (uint32_t)bins - 1;
    const uint32_t lt = (1u << lane) - 1u;
    // each warp owns a contiguous run of the block's chunk (stability: warp order, round order, lane order)
    const int64_t wbase = (int64_t)blockIdx.x * SORT_ITEMS_PER_BLOCK + (int64_t)w * (RP_ROUNDS * 32);
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t d = (keys_in[i] >> shift) & mask;
            const uint32_t peers = __match_any_sync(act, d);
            if ((peers & lt) == 0) wh[w][d] += __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    for (int d = threadIdx.x; d < bins; d += RP_THREADS) {
        uint32_t base = hist_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int k = 0; k < RP_WARPS; ++k) {
            const uint32_t t = wh[k][d];
            wh[k][d] = base;
            base += t;
        }
    }
    __syncthreads();
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        const uint32_t act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const uint32_t key = keys_in[i], val = vals_in[i];
            const uint32_t d = (key >> shift) & mask;
            const uint32_t peers = __match_any_sync(act, d);
            const uint32_t pos = wh[w][d] + __popc(peers & lt);
            __syncwarp(act);
            if ((peers & lt) == 0) wh[w][d] += __popc(peers);
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncwarp();
    }
}

int radix_pass(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out, uint32_t *vals_out, int64_t n,
               int shift, int bits, uint32_t *hist, uint32_t *scan_tmp, cudaStream_t s)
{
    if (n <= 0) return 0;
    const int bins = 1 << bits;
    const int nblocks = (int)((n + SORT_ITEMS_PER_BLOCK - 1) / SORT_ITEMS_PER_BLOCK);
    radix_hist_kernel<<<nblocks, RP_THREADS, 0, s>>>(keys_in, n, shift, bins, nblocks, hist);
    int rc = exclusive_scan_u32(hist, nullptr, hist, (int64_t)bins * nblocks, nullptr, scan_tmp, s);
    if (rc) return rc;
    radix_scatter_kernel<<<nblocks, RP_THREADS, 0, s>>>(keys_in, vals_in, keys_out, vals_out, n, shift, bins, nblocks, hist);
    DMGS_CUDA(cudaGetLastError());
    count_launches(2);
    return 0;
}

// ------------------------------------------------------------------------------ instance emission
__global__ void __launch_bounds__(256)
emit_instances_kernel(int P, const uint32_t *__restrict__ order, const uint32_t *__restrict__ offsets,
                      const uint32_t *__restrict__ tiles, const uint2 *__restrict__ rect, int gx,
                      uint32_t *__restrict__ inst_tile, uint32_t *__restrict__ inst_gidx)
{
    const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t g = 0, cnt = 0, off = 0;
    uint2 rc = make_uint2(0, 0);
    if (sidx < P) {
        g = order[sidx];
        cnt = tiles[g];
        if (cnt) {
            off = offsets[sidx];
            rc = rect[g];
        }
    }
    // small rectangles: one thread each; large ones: the whole warp cooperates
    const uint32_t big = __ballot_sync(0xffffffffu, cnt > 32);
    if (cnt && cnt <= 32) {
        const int x0 = rc.x & 0xffff, x1 = rc.x >> 16, y0 = rc.y & 0xffff, y1 = rc.y >> 16;
        uint32_t o = off;
        for (int ty = y0; ty < y1; ++ty)
            for (int tx = x0; tx < x1; ++tx) {
                inst_tile[o] = (uint32_t)(ty * gx + tx);
                inst_gidx[o] = g;
                ++o;
            }
    }
    uint32_t todo = big;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t g2 = __shfl_sync(0xffffffffu, g, src), c2 = __shfl_sync(0xffffffffu, cnt, src);
        const uint32_t o2 = __shfl_sync(0xffffffffu, off, src);
        const uint32_t rx = __shfl_sync(0xffffffffu, rc.x, src), ry = __shfl_sync(0xffffffffu, rc.y, src);
        const int x0 = rx & 0xffff, x1 = rx >> 16, y0 = ry & 0xffff;
        const int wdt = x1 - x0;
        for (uint32_t j = lane; j < c2; j += 32) {
            const int ty = y0 + (int)(j / wdt), tx = x0 + (int)(j % wdt);
            inst_tile[o2 + j] = (uint32_t)(ty * gx + tx);
            inst_gidx[o2 + j] = g2;
        }
    }
}

int launch_emit_instances(int P, const uint32_t *order, const uint32_t *offsets, const uint32_t *tiles,
                          const uint2 *rect, int gx, uint32_t *inst_tile, uint32_t *inst_gidx, cudaStream_t s)
{
    if (P <= 0) return 0;
    emit_instances_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, order, offsets, tiles, rect, gx, inst_tile, inst_gidx);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------ tile ranges
__global__ void __launch_bounds__(256)
tile_ranges_kernel(int64_t R, const uint32_t *__restrict__ sorted_tiles, uint2 *__restrict__ ranges)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= R) return;
    const uint32_t t = sorted_tiles[j];
    if (j == 0 || sorted_tiles[j - 1] != t) ranges[t].x = (uint32_t)j;
    if (j == R - 1 || sorted_tiles[j + 1] != t) ranges[t].y = (uint32_t)(j + 1);
}

int launch_tile_ranges(int64_t R, const uint32_t *sorted_tiles, uint2 *ranges, int T, cudaStream_t s)
{
    DMGS_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)T, s));
    if (R <= 0) return 0;
    tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, sorted_tiles, ranges);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

__global__ void sorted_keys_kernel(int64_t R, const uint32_t *__restrict__ sorted_tiles,
                                   const uint32_t *__restrict__ sorted_gidx, const float *__restrict__ depths,
                                   uint64_t *__restrict__ keys_out)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= R) return;
    keys_out[j] = ((uint64_t)sorted_tiles[j] << 32) | (uint64_t)__float_as_uint(depths[sorted_gidx[j]]);
}

int launch_sorted_keys(int64_t R, const uint32_t *sorted_tiles, const uint32_t *sorted_gidx, const float *depths,
                       uint64_t *keys_out, cudaStream_t s)
{
    if (R <= 0) return 0;
    sorted_keys_kernel<<<(unsigned)((R + 255) / 256), 256, 0, s>>>(R, sorted_tiles, sorted_gidx, depths, keys_out);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
