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
// One stable LSD pass = histogram kernel -> scan of the [bins x blocks] table -> ranked scatter.
// A block owns 2048 consecutive items; warp w owns the contiguous run [w*256, (w+1)*256) of it and
// walks it in 8 rounds of 32 (stability = warp order, round order, lane order).  Keys live in
// registers (all loads issued up front), equal-digit lanes are found with `bits` warp ballots
// (MATCH.ANY saturates the ADU pipe on sm_100: profiles/r1_v1), and per-warp digit counters in
// shared memory turn the block's scanned table column into final positions.
constexpr int RP_THREADS = 256;
constexpr int RP_WARPS = RP_THREADS / 32;
#ifndef SCATTER_MIN_BLOCKS
#define SCATTER_MIN_BLOCKS 4
#endif

// Adaptive passes: `stat` (optional) holds {max of ~key, max of key} over the live keys, i.e. their
// minimum and maximum.  The passes sort on (key - min), whose significant bits are those of
// (max - min): a pass over a digit above them has nothing to do, its kernels exit at once, and every
// kernel finds its input in buffer A or B from the number of passes that did run before it.
struct PassSel {
    bool active;
    int src;        // 0: (keys_a, vals_a) -> (keys_b, vals_b); 1: the other way round
    uint32_t kmin;  // subtracted from every key before the digit is taken
};
__device__ __forceinline__ PassSel pass_select(const uint32_t *__restrict__ stat, int pass, int shift)
{
    PassSel r;
    if (!stat) {
        r.active = true;
        r.src = pass & 1;
        r.kmin = 0;
        return r;
    }
    const uint32_t kmin = ~stat[0], kmax = stat[1];
    const uint32_t range = kmax >= kmin ? kmax - kmin : 0u;  // no live key: nothing to sort
    r.active = (range >> shift) != 0 || (pass == 0 && kmax >= kmin);
    r.src = pass & 1;  // 8-bit digits from bit 0: the active passes are a prefix 0..k-1 of the sequence
    r.kmin = kmin;
    return r;
}

template <int RP_ROUNDS>
__global__ void __launch_bounds__(RP_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys_a, const uint32_t *__restrict__ keys_b, int64_t n, int shift,
                  int bins, int nblocks, uint32_t *__restrict__ hist, const uint32_t *__restrict__ stat, int pass,
                  int bits)
{
    (void)bits;
    const PassSel ps = pass_select(stat, pass, shift);
    if (!ps.active) return;
    const uint32_t *__restrict__ keys = ps.src ? keys_b : keys_a;
    __shared__ uint32_t h[SORT_MAX_BINS];
    for (int d = threadIdx.x; d < bins; d += RP_THREADS) h[d] = 0;
    constexpr int ITEMS = RP_ROUNDS * RP_THREADS;
    const int64_t base = (int64_t)blockIdx.x * ITEMS;
    const uint32_t mask = (uint32_t)bins - 1;
    uint32_t k[RP_ROUNDS];
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RP_THREADS + threadIdx.x;
        k[r] = i < n ? keys[i] : 0xFFFFFFFFu;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RP_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[((k[r] - ps.kmin) >> shift) & mask], 1u);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < bins; d += RP_THREADS) hist[(size_t)d * nblocks + blockIdx.x] = h[d];
}

template <int BITS>
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, uint32_t act)
{
    uint32_t peers = act;
#pragma unroll
    for (int b = 0; b < BITS; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

template <int BITS, int RP_ROUNDS>
__global__ void __launch_bounds__(RP_THREADS, SCATTER_MIN_BLOCKS)
radix_scatter_kernel(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int shift,
                     int nblocks, const uint32_t *__restrict__ row_prefix, const uint32_t *__restrict__ bin_total,
                     const uint32_t *__restrict__ stat, int pass)
{
    const PassSel ps = pass_select(stat, pass, shift);
    if (!ps.active) return;
    const uint32_t *__restrict__ keys_in = ps.src ? keys_b : keys_a, *__restrict__ vals_in = ps.src ? vals_b : vals_a;
    uint32_t *__restrict__ keys_out = ps.src ? keys_a : keys_b, *__restrict__ vals_out = ps.src ? vals_a : vals_b;
    constexpr int BINS = 1 << BITS;
    constexpr int ITEMS = RP_ROUNDS * RP_THREADS;
    __shared__ uint32_t wh[RP_WARPS][BINS];
    __shared__ uint32_t bin_local[BINS];   // first slot of digit d inside the block's reordered chunk
    __shared__ uint32_t bin_global[BINS];  // global position of that slot
    __shared__ uint32_t skeys[ITEMS];
    __shared__ uint32_t svals[ITEMS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RP_WARPS * BINS; d += RP_THREADS) (&wh[0][0])[d] = 0;
    const uint32_t mask = (uint32_t)BINS - 1;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t bbase = (int64_t)blockIdx.x * ITEMS;
    const int64_t wbase = bbase + (int64_t)w * (RP_ROUNDS * 32) + lane;
    uint32_t key[RP_ROUNDS], val[RP_ROUNDS], rank[RP_ROUNDS];
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = wbase + r * 32;
        key[r] = i < n ? keys_in[i] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const int64_t i = wbase + r * 32;
        val[r] = i < n ? vals_in[i] : 0u;
    }
    __syncthreads();
    // phase 1: stable rank of every item among the warp's items with the same digit.  Lanes with
    // equal digits are found with BITS ballots; the lowest such lane bumps the warp's counter with
    // one shared-memory atomic (same-warp atomics to one address retire in program order, so the
    // rounds need no barrier between them and overlap freely) and broadcasts the old value.
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        const bool valid = wbase + r * 32 < n;
        const uint32_t act = __ballot_sync(0xffffffffu, valid);
        const uint32_t d = ((key[r] - ps.kmin) >> shift) & mask;
        const uint32_t peers = digit_peers<BITS>(d, act);
        uint32_t old = 0;
        if (valid && (peers & lt) == 0) old = atomicAdd(&wh[w][d], (uint32_t)__popc(peers));
        old = __shfl_sync(0xffffffffu, old, peers ? __ffs(peers) - 1 : 0);
        rank[r] = old + __popc(peers & lt);
    }
    __syncthreads();
    // phase 2: per digit, exclusive prefix over warps; block-level exclusive scan over digits;
    // global base = (scan of the bin totals) + (this block's prefix inside the bin's table row)
    {
        const int d = threadIdx.x;
        uint32_t tot = 0, gtot = 0;
        if (d < BINS) {
#pragma unroll
            for (int k = 0; k < RP_WARPS; ++k) {
                const uint32_t t = wh[k][d];
                wh[k][d] = tot;
                tot += t;
            }
            gtot = bin_total[d];
        }
        uint32_t dummy;
        const uint32_t lex = block_excl_scan(tot, &dummy);
        const uint32_t gex = block_excl_scan(gtot, &dummy);
        if (d < BINS) {
            bin_local[d] = lex;
            bin_global[d] = gex + row_prefix[(size_t)d * nblocks + blockIdx.x];
        }
    }
    __syncthreads();
    // phase 3: shared-memory reorder
#pragma unroll
    for (int r = 0; r < RP_ROUNDS; ++r) {
        if (wbase + r * 32 < n) {
            const uint32_t d = ((key[r] - ps.kmin) >> shift) & mask;
            const uint32_t pos = bin_local[d] + wh[w][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }
    __syncthreads();
    // phase 4: coalesced write-out (each digit's run is contiguous both in shared and in global memory)
    const int count = (int)min((int64_t)ITEMS, n - bbase);
#pragma unroll 4
    for (int j = threadIdx.x; j < count; j += RP_THREADS) {
        const uint32_t k = skeys[j];
        const uint32_t d = ((k - ps.kmin) >> shift) & mask;
        const uint32_t out = bin_global[d] + ((uint32_t)j - bin_local[d]);
        keys_out[out] = k;
        vals_out[out] = svals[j];
    }
}

// one block per digit: exclusive scan of the digit's table row (over blocks) + the row total
__global__ void __launch_bounds__(256)
radix_rowscan_kernel(uint32_t *__restrict__ hist, int nblocks, uint32_t *__restrict__ bin_total,
                     const uint32_t *__restrict__ stat, int pass, int shift, int bits)
{
    (void)bits;
    if (!pass_select(stat, pass, shift).active) return;
    uint32_t *row = hist + (size_t)blockIdx.x * nblocks;
    uint32_t carry = 0;
    for (int b0 = 0; b0 < nblocks; b0 += 256 * 4) {
        const int i0 = b0 + threadIdx.x * 4;
        uint32_t v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k] = (i0 + k < nblocks) ? row[i0 + k] : 0; s += v[k]; }
        uint32_t tot;
        uint32_t ex = carry + block_excl_scan(s, &tot);
#pragma unroll
        for (int k = 0; k < 4; ++k) { if (i0 + k < nblocks) row[i0 + k] = ex; ex += v[k]; }
        carry += tot;
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = carry;
}

template <int BITS>
static void launch_scatter(int rounds, uint32_t *ka, uint32_t *va, uint32_t *kb, uint32_t *vb, int64_t n, int shift,
                           int nblocks, const uint32_t *hist, const uint32_t *bin_total, const uint32_t *stat, int pass,
                           cudaStream_t s)
{
    if (rounds == 4)
        radix_scatter_kernel<BITS, 4><<<nblocks, RP_THREADS, 0, s>>>(ka, va, kb, vb, n, shift, nblocks, hist, bin_total, stat, pass);
    else
        radix_scatter_kernel<BITS, 8><<<nblocks, RP_THREADS, 0, s>>>(ka, va, kb, vb, n, shift, nblocks, hist, bin_total, stat, pass);
}

// One stable LSD pass on `bits` bits starting at `shift`.  Without `stat`, pass p reads (keys_a, vals_a)
// when p is even and (keys_b, vals_b) when odd and writes the other pair.  With `stat` (adaptive depth
// sort, 8-bit digits) passes over digits that do not vary are skipped on the device and the buffers
// alternate over the passes that did run.
int radix_pass(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int pass, int shift,
               int bits, uint32_t *hist, const uint32_t *stat, cudaStream_t s)
{
    if (n <= 0) return 0;
    if (bits < 1 || bits > 8) { set_error("radix_pass: bits=%d unsupported", bits); return -8; }
    const int bins = 1 << bits;
    // small inputs get 1024-item tiles: more blocks than resident slots (148 SMs x 3-4 CTAs)
    const int rounds = n <= RADIX_SMALL_N ? 4 : 8;
    const int items = rounds * RP_THREADS;
    const int nblocks = (int)((n + items - 1) / items);
    uint32_t *bin_total = hist + (size_t)SORT_MAX_BINS * nblocks;  // 256 spare entries behind the table
    if (rounds == 4) radix_hist_kernel<4><<<nblocks, RP_THREADS, 0, s>>>(keys_a, keys_b, n, shift, bins, nblocks, hist, stat, pass, bits);
    else radix_hist_kernel<8><<<nblocks, RP_THREADS, 0, s>>>(keys_a, keys_b, n, shift, bins, nblocks, hist, stat, pass, bits);
    radix_rowscan_kernel<<<bins, 256, 0, s>>>(hist, nblocks, bin_total, stat, pass, shift, bits);
    switch (bits) {
    case 1: launch_scatter<1>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 2: launch_scatter<2>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 3: launch_scatter<3>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 4: launch_scatter<4>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 5: launch_scatter<5>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 6: launch_scatter<6>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    case 7: launch_scatter<7>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    default: launch_scatter<8>(rounds, keys_a, vals_a, keys_b, vals_b, n, shift, nblocks, hist, bin_total, stat, pass, s); break;
    }
    DMGS_CUDA(cudaGetLastError());
    count_launches(3);
    return 0;
}

// ------------------------------------------------------------------------------ onesweep depth sort
// The depth sort of the P Gaussians is launch-latency bound (8 MB of keys per pass): three kernels per digit were
// 9-12 small launches per frame.  Onesweep: ONE kernel builds the global histograms of all four digits (and clears
// the look-back words), then ONE kernel per digit ranks a 2048-key tile exactly like radix_scatter_kernel and gets
// the number of equal-digit keys in earlier tiles by decoupled look-back: every tile publishes its per-digit count
// (AGG) as soon as it is known, then walks back over its predecessors' words until it meets an inclusive PREFIX,
// and publishes its own.  Tiles are numbered by an atomic counter in the order their blocks START, so every
// predecessor a block waits for is already running: the spin cannot dead-lock.
constexpr uint32_t OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_VALUE = (1u << 30) - 1u;

__global__ void __launch_bounds__(RP_THREADS)
depth_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, const uint32_t *__restrict__ stat,
                  uint32_t *__restrict__ ghist, uint32_t *__restrict__ status, int ntiles)
{
    __shared__ uint32_t h[4][SORT_MAX_BINS];
    for (int p = 0; p < 4; ++p) {
        h[p][threadIdx.x] = 0;
        status[((size_t)p * ntiles + blockIdx.x) * SORT_MAX_BINS + threadIdx.x] = 0;
    }
    const uint32_t kmin = stat ? ~stat[0] : 0u;
    const int64_t base = (int64_t)blockIdx.x * ONESWEEP_ITEMS;
    uint32_t k[ONESWEEP_ITEMS / RP_THREADS];
#pragma unroll
    for (int r = 0; r < ONESWEEP_ITEMS / RP_THREADS; ++r) {
        const int64_t i = base + (int64_t)r * RP_THREADS + threadIdx.x;
        k[r] = i < n ? keys[i] : 0xFFFFFFFFu;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ONESWEEP_ITEMS / RP_THREADS; ++r) {
        if (base + (int64_t)r * RP_THREADS + threadIdx.x < n) {
            const uint32_t v = k[r] - kmin;
#pragma unroll
            for (int p = 0; p < 4; ++p) atomicAdd(&h[p][(v >> (8 * p)) & 255u], 1u);
        }
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t c = h[p][threadIdx.x];
        if (c) atomicAdd(&ghist[p * SORT_MAX_BINS + threadIdx.x], c);
    }
}

__global__ void __launch_bounds__(RP_THREADS, SCATTER_MIN_BLOCKS)
onesweep_pass_kernel(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n, int pass,
                     int ntiles, const uint32_t *__restrict__ ghist, uint32_t *status, uint32_t *counters,
                     const uint32_t *__restrict__ stat)
{
    const int shift = 8 * pass;
    const PassSel ps = pass_select(stat, pass, shift);
    if (!ps.active) return;
    const uint32_t *__restrict__ keys_in = ps.src ? keys_b : keys_a, *__restrict__ vals_in = ps.src ? vals_b : vals_a;
    uint32_t *__restrict__ keys_out = ps.src ? keys_a : keys_b, *__restrict__ vals_out = ps.src ? vals_a : vals_b;
    constexpr int BINS = SORT_MAX_BINS, ROUNDS = ONESWEEP_ITEMS / RP_THREADS, ITEMS = ONESWEEP_ITEMS;
    __shared__ uint32_t wh[RP_WARPS][BINS];
    __shared__ uint32_t bin_local[BINS];   // first slot of digit d inside the tile's reordered chunk
    __shared__ uint32_t bin_global[BINS];  // global position of that slot
    __shared__ uint32_t skeys[ITEMS];
    __shared__ uint32_t svals[ITEMS];
    __shared__ uint32_t s_tile;
    if (threadIdx.x == 0) s_tile = atomicAdd(&counters[pass], 1u);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RP_WARPS * BINS; d += RP_THREADS) (&wh[0][0])[d] = 0;
    __syncthreads();
    const int tile = (int)s_tile;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t bbase = (int64_t)tile * ITEMS;
    const int64_t wbase = bbase + (int64_t)w * (ROUNDS * 32) + lane;
    uint32_t key[ROUNDS], val[ROUNDS], rank[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int64_t i = wbase + r * 32;
        key[r] = i < n ? keys_in[i] : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const int64_t i = wbase + r * 32;
        val[r] = i < n ? vals_in[i] : 0u;
    }
    // stable rank of every key among the warp's keys with the same digit (see radix_scatter_kernel)
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        const bool valid = wbase + r * 32 < n;
        const uint32_t act = __ballot_sync(0xffffffffu, valid);
        const uint32_t d = ((key[r] - ps.kmin) >> shift) & 255u;
        const uint32_t peers = digit_peers<8>(d, act);
        uint32_t old = 0;
        if (valid && (peers & lt) == 0) old = atomicAdd(&wh[w][d], (uint32_t)__popc(peers));
        old = __shfl_sync(0xffffffffu, old, peers ? __ffs(peers) - 1 : 0);
        rank[r] = old + __popc(peers & lt);
    }
    __syncthreads();
    {
        const int d = threadIdx.x;  // one digit per thread (RP_THREADS == BINS)
        uint32_t tot = 0;
#pragma unroll
        for (int k = 0; k < RP_WARPS; ++k) {
            const uint32_t t = wh[k][d];
            wh[k][d] = tot;
            tot += t;
        }
        // publish, then look back
        volatile uint32_t *st = status + (size_t)pass * ntiles * BINS;
        st[(size_t)tile * BINS + d] = (tile == 0 ? OS_PREFIX : OS_AGG) | tot;
        const uint32_t gtot = ghist[pass * BINS + d];
        uint32_t dummy;
        const uint32_t lex = block_excl_scan(tot, &dummy);
        const uint32_t gex = block_excl_scan(gtot, &dummy);
        uint32_t excl = 0;
        if (tile > 0) {
            // eight predecessors per round trip: all tiles of a small sort start together, so the walk back to the
            // first inclusive prefix is long and its latency, not its traffic, is what a pass costs
            bool found = false;
            for (int t = tile - 1; !found; t -= 8) {
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = t - u >= 0 ? (uint32_t)st[(size_t)(t - u) * BINS + d] : (uint32_t)(2u << 30);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (found) break;
                    uint32_t x = v[u];
                    while ((x & (OS_AGG | OS_PREFIX)) == 0) x = st[(size_t)(t - u) * BINS + d];
                    excl += x & OS_VALUE;
                    found = (x & OS_PREFIX) != 0;
                }
            }
            st[(size_t)tile * BINS + d] = OS_PREFIX | (excl + tot);
        }
        bin_local[d] = lex;
        bin_global[d] = gex + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
        if (wbase + r * 32 < n) {
            const uint32_t d = ((key[r] - ps.kmin) >> shift) & 255u;
            const uint32_t pos = bin_local[d] + wh[w][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = val[r];
        }
    }
    __syncthreads();
    const int count = (int)min((int64_t)ITEMS, n - bbase);
#pragma unroll 4
    for (int j = threadIdx.x; j < count; j += RP_THREADS) {
        const uint32_t k = skeys[j];
        const uint32_t d = ((k - ps.kmin) >> shift) & 255u;
        const uint32_t out = bin_global[d] + ((uint32_t)j - bin_local[d]);
        keys_out[out] = k;
        vals_out[out] = svals[j];
    }
}

int depth_sort_onesweep(uint32_t *keys_a, uint32_t *vals_a, uint32_t *keys_b, uint32_t *vals_b, int64_t n,
                        uint32_t *ghist, uint32_t *counters, uint32_t *status, const uint32_t *stat, cudaStream_t s)
{
    if (n <= 0) return 0;
    static_assert(RP_THREADS == SORT_MAX_BINS, "one digit per thread");
    const int ntiles = (int)((n + ONESWEEP_ITEMS - 1) / ONESWEEP_ITEMS);
    if (!stat) {
        DMGS_CUDA(cudaMemsetAsync(ghist, 0, 4 * SORT_MAX_BINS * sizeof(uint32_t), s));
        DMGS_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(uint32_t), s));
    }
    depth_hist_kernel<<<ntiles, RP_THREADS, 0, s>>>(keys_a, n, stat, ghist, status, ntiles);
    for (int pass = 0; pass < 4; ++pass)
        onesweep_pass_kernel<<<ntiles, RP_THREADS, 0, s>>>(keys_a, vals_a, keys_b, vals_b, n, pass, ntiles, ghist, status,
                                                           counters, stat);
    DMGS_CUDA(cudaGetLastError());
    count_launches(5);
    return 0;
}

// ------------------------------------------------------------------------------ instance emission
// Load-balanced emission: a warp takes 32 depth-ordered Gaussians whose instances occupy one
// contiguous output range; lanes walk that range with stride 32 (fully coalesced stores) and find
// the owning Gaussian of each slot by a 5-step search over the warp's offsets.
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
        off = offsets[sidx];
        if (cnt) rc = rect[g];
    }
    // offsets are an exclusive scan in this order: lane l's instances are [off_l, off_l + cnt_l)
    const uint32_t first = __shfl_sync(0xffffffffu, off, 0);
    const uint32_t last_off = __shfl_sync(0xffffffffu, off, 31), last_cnt = __shfl_sync(0xffffffffu, cnt, 31);
    uint32_t total = last_off + last_cnt - first;
    if (sidx - lane + 31 >= P) {  // partial warp at the tail: lanes beyond P hold zeros
        const uint32_t end = off + cnt;
        uint32_t m = sidx < P ? end : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
        total = m > first ? m - first : 0;
    }
    const uint32_t rel = off - first;  // start of this lane's run relative to the warp's range
    for (uint32_t jb = 0; jb < total; jb += 32) {  // warp-uniform trip count (shuffles inside)
        const uint32_t j = jb + lane;
        // largest lane l with rel_l <= j and cnt_l > 0 covering j: binary search on rel (non-decreasing)
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int cand = lo + step;
            const uint32_t r2 = __shfl_sync(0xffffffffu, rel, cand & 31);
            const bool okl = (sidx - lane + cand) < P;
            if (cand < 32 && okl && r2 <= j) lo = cand;
        }
        const uint32_t g2 = __shfl_sync(0xffffffffu, g, lo);
        const uint32_t rx = __shfl_sync(0xffffffffu, rc.x, lo), ry = __shfl_sync(0xffffffffu, rc.y, lo);
        const uint32_t r0 = __shfl_sync(0xffffffffu, rel, lo);
        const int x0 = rx & 0xffff, x1 = rx >> 16, y0 = ry & 0xffff;
        const uint32_t wdt = (uint32_t)(x1 - x0), k = j - r0;
        if (j < total) {
            const uint32_t ty = y0 + k / wdt, tx = x0 + k % wdt;
            inst_tile[first + j] = ty * (uint32_t)gx + tx;
            inst_gidx[first + j] = g2;
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
    // four consecutive entries per thread (one 16-byte load) + the entry before them
    const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (j0 >= R) return;
    uint32_t t[5];
    if (j0 + 3 < R) {
        const uint4 v = *reinterpret_cast<const uint4 *>(sorted_tiles + j0);
        t[1] = v.x; t[2] = v.y; t[3] = v.z; t[4] = v.w;
    } else {
        for (int k = 0; k < 4; ++k) t[1 + k] = j0 + k < R ? sorted_tiles[j0 + k] : 0xFFFFFFFFu;
    }
    t[0] = j0 > 0 ? sorted_tiles[j0 - 1] : 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t j = j0 + k;
        if (j >= R) break;
        if (t[k] != t[k + 1]) {
            ranges[t[k + 1]].x = (uint32_t)j;
            if (j > 0) ranges[t[k]].y = (uint32_t)j;
        }
        if (j == R - 1) ranges[t[k + 1]].y = (uint32_t)R;
    }
}

int launch_tile_ranges(int64_t R, const uint32_t *sorted_tiles, uint2 *ranges, int T, cudaStream_t s)
{
    DMGS_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint2) * (size_t)T, s));
    if (R <= 0) return 0;
    tile_ranges_kernel<<<(unsigned)((R + 1023) / 1024), 256, 0, s>>>(R, sorted_tiles, ranges);
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
