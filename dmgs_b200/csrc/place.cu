// place.cu -- direct tile placement: the per-tile, depth-ordered Gaussian lists without moving
// (tile, Gaussian) pairs through a radix sort.
//
// Replaces duplicateWithKeys + SortPairs(64-bit) + identifyTileRanges for images of
// up to PLACE_MAX_TILES tiles; larger images use the radix tile partition in sort.cu.  Input is the
// depth-sorted Gaussian order produced by dmgs_preprocess_forward (stable, ties in ascending index).
//
//   0  sorted_rect_kernel  packs every Gaussian's tile rectangle in depth order (16-byte records);
//   A  tile_count_kernel   the depth-ordered Gaussians are cut into `nseg` contiguous segments, one
//                          per warp; each warp counts the tiles its Gaussians touch in its own
//                          shared-memory counters and writes one row of table[nseg][T];
//   B  group_scan / tile_scan
//                          exclusive scan of the per-CTA sums down every tile column and of the tile
//                          totals: together with the rows of the table they give every segment its first
//                          slot in every tile's list, and ranges[t] falls out for free;
//   C  tile_place_kernel   each warp re-walks its segment IN ORDER, one Gaussian per step with lanes
//                          over the tiles of its rectangle (distinct tiles, so no conflicts), bumping
//                          its shared-memory cursors and writing the Gaussian index to its final slot.
//
// Because segments are contiguous in depth order and a warp walks its segment sequentially, every
// tile's list comes out in depth order with ties in ascending Gaussian index -- exactly the order a
// stable sort of (tile << 32 | depth bits) keys produces (the parity tests rebuild those keys and
// compare them bit for bit).  R-sized traffic is ONE 4-byte write per instance (plus the table),
// against 48 B/instance for emission + two radix passes and 152 B/instance for a 64-bit sort.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

// Shared memory per SM given to the per-warp tile counters.  200 KB (20 warps/SM at 2500 tiles) is the LAYOUT plan:
// it sizes the count table and decides whether an image takes the placement path at all.  The kernels may RUN with
// less (dmgs_set_place_smem_kb): they are latency bound (34 % issue utilisation), but at 200 KB no CTA of another
// stream fits beside them on the SM; at 128 KB six blend CTAs do, and the walk is only 4 % slower.  A smaller
// budget means fewer, longer segments, i.e. a prefix of the rows the layout reserved.
int g_place_smem_kb = [] {
    const char *v = getenv("DMGS_PLACE_SMEM_KB");
    const int kb = v ? atoi(v) : 0;
    return kb >= 64 && kb <= 200 ? kb : 200;
}();

static PlacePlan place_plan_budget(int32_t P, int T, size_t budget)
{
    PlacePlan p;
    memset(&p, 0, sizeof(p));
    const size_t per_warp = (size_t)T * 4;
    if (T < 1 || T > PLACE_MAX_TILES) return p;
    // test hook: DMGS_TILE_PARTITION=radix forces the radix tile partition (sort.cu) for any image size
    const char *force = getenv("DMGS_TILE_PARTITION");
    if (force && strcmp(force, "radix") == 0) return p;
    int wps = (int)(budget / per_warp);
    if (wps > 32) wps = 32;
    if (wps < 2) return p;
    p.wpb = wps >= 4 ? 4 : wps >= 2 ? 2 : 1;
    const int nseg_max = num_sms() * (wps / p.wpb) * p.wpb;
    const int64_t n = P > 0 ? P : 1;
    int seg = (int)((n + nseg_max - 1) / nseg_max);
    if (seg < 64) seg = 64;
    seg = (seg + 31) / 32 * 32;
    p.seg = seg;
    p.nseg = (int)((n + seg - 1) / seg);
    p.rows_per_group = p.wpb;  // one group of table rows per CTA of the count / place kernels
    p.groups = (p.nseg + p.wpb - 1) / p.wpb;
    p.smem = (size_t)p.wpb * per_warp;
    p.ok = 1;
    return p;
}
PlacePlan place_plan(int32_t P, int T) { return place_plan_budget(P, T, 200 * 1024); }
PlacePlan place_plan_run(int32_t P, int T)
{
    const PlacePlan full = place_plan(P, T);
    if (!full.ok || g_place_smem_kb >= 200) return full;
    const PlacePlan p = place_plan_budget(P, T, (size_t)g_place_smem_kb * 1024);
    // never more rows than the layout reserved (table: nseg rows; per-CTA sums: ceil(nseg / 2) rows)
    return p.ok && p.nseg <= full.nseg && p.groups <= (full.nseg + 1) / 2 + 1 ? p : full;
}

// ------------------------------------------------------------------------------ depth-ordered rectangles
// srec[s] = {x0 | x1 << 16, y0 | y1 << 16, ceil(2^32 / width), Gaussian index} of the s-th Gaussian in
// depth order: the walkers below read ONE coalesced 16-byte record per Gaussian instead of chasing
// order[s] -> rect[g] through two dependent gathers per 32-Gaussian batch.
__global__ void __launch_bounds__(256)
sorted_rect_kernel(int P, const uint32_t *__restrict__ order_a, const uint32_t *__restrict__ order_b,
                   const uint32_t *__restrict__ stat, const uint2 *__restrict__ rect, uint4 *__restrict__ srec)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P) return;
    // the adaptive depth sort leaves its result in A or B depending on how many passes ran
    const uint32_t kmin = ~stat[0], kmax = stat[1];
    const uint32_t range = kmax >= kmin ? kmax - kmin : 0u;
    int ran = kmax >= kmin ? 1 : 0;  // pass 0 always runs when there is a live key
#pragma unroll
    for (int q = 1; q < 4; ++q) ran += (range >> (8 * q)) != 0;
    const uint32_t g = (ran & 1) ? order_b[s] : order_a[s];
    const uint2 rc = rect[g];
    const uint32_t wd = (rc.x >> 16) - (rc.x & 0xffff);
    const uint32_t magic = wd > 1 ? (uint32_t)((0x100000000ull + wd - 1) / wd) : 0u;
    srec[s] = make_uint4(rc.x, rc.y, magic, g);
}

// ------------------------------------------------------------------------------ A: count
// Every warp counts the tiles its segment touches as a 2-D difference image: a rectangle [x0,x1) x [y0,y1)
// is FOUR shared atomics (+1, -1, -1, +1 at its corners on a (gy+1) x (gx+1) grid), whatever its size, instead
// of one per covered tile; two prefix passes (lanes over rows, then lanes over columns; row stride gx+1 is
// odd for the common image sizes, so both passes are bank-conflict free) turn the image into the per-tile counts.  Each warp writes its row
// of table[nseg][T]; the CTA also writes the sum of its rows to gsum[cta][T], the coarse level of the column scan.
__global__ void __launch_bounds__(128)
tile_count_kernel(int P, int T, int gx, int seg, int nseg, const uint4 *__restrict__ srec, uint32_t *__restrict__ table,
                  uint32_t *__restrict__ gsum)
{
    extern __shared__ uint32_t s_cnt[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int sg = blockIdx.x * wpb + w;
    const int gy = T / gx, sx = gx + 1, G = sx * (gy + 1);
    uint32_t *cnt = s_cnt + (size_t)w * G;
    for (int t = lane; t < G; t += 32) cnt[t] = 0;
    __syncwarp();
    if (sg < nseg) {
        const int s0 = sg * seg, s1 = min(P, s0 + seg);
        for (int sb = s0 + lane; sb < s1; sb += 32) {
            const uint4 rc = srec[sb];
            const int x0 = rc.x & 0xffff, x1 = rc.x >> 16, y0 = rc.y & 0xffff, y1 = rc.y >> 16;
            if (x1 > x0 && y1 > y0) {
                atomicAdd(cnt + y0 * sx + x0, 1u);
                atomicAdd(cnt + y0 * sx + x1, 0xffffffffu);
                atomicAdd(cnt + y1 * sx + x0, 0xffffffffu);
                atomicAdd(cnt + y1 * sx + x1, 1u);
            }
        }
        __syncwarp();
        for (int y = lane; y < gy; y += 32) {  // prefix along x
            uint32_t run = 0;
            for (int x = 0; x < gx; ++x) { run += cnt[y * sx + x]; cnt[y * sx + x] = run; }
        }
        __syncwarp();
        uint32_t *row = table + (size_t)sg * T;
        for (int x = lane; x < gx; x += 32) {  // prefix along y; the finished counts go straight to the table
            uint32_t run = 0;
            for (int y = 0; y < gy; ++y) { run += cnt[y * sx + x]; cnt[y * sx + x] = run; row[y * gx + x] = run; }
        }
    }
    __syncthreads();
    uint32_t *grow = gsum + (size_t)blockIdx.x * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        const int y = t / gx, x = t - y * gx;
        uint32_t sum = 0;
        for (int k = 0; k < wpb; ++k) sum += s_cnt[(size_t)k * G + y * sx + x];
        grow[t] = sum;
    }
}

// ------------------------------------------------------------------------------ B: column scan
// Exclusive scan of the per-CTA sums down every tile column (in place) and the tile totals.  A block owns
// 32 adjacent tiles; its 16 warps cut the column into 16 slabs of consecutive groups.  Lanes <-> tiles, so
// every load and store is a coalesced 128-byte row segment of gsum[group][tile] (the first version read
// down the columns: one sector per lane and load).  Pass 1 sums the slabs, a 16-entry scan across the
// warps gives every slab its base, pass 2 writes the running prefix.
constexpr int GS_WARPS = 16;
__global__ void __launch_bounds__(GS_WARPS * 32)
group_scan_kernel(int T, int groups, uint32_t *__restrict__ gsum, uint32_t *__restrict__ tile_total)
{
    __shared__ uint32_t slab[GS_WARPS][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const int per = (groups + GS_WARPS - 1) / GS_WARPS;
    const int g0 = w * per, g1 = min(groups, g0 + per);
    uint32_t sum = 0;
    if (t < T) {
#pragma unroll 8
        for (int g = g0; g < g1; ++g) sum += gsum[(size_t)g * T + t];
    }
    slab[w][lane] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
#pragma unroll
    for (int k = 0; k < GS_WARPS; ++k) {
        const uint32_t v = slab[k][lane];
        if (k < w) run += v;
        total += v;
    }
    if (t < T) {
#pragma unroll 8
        for (int g = g0; g < g1; ++g) {
            const uint32_t v = gsum[(size_t)g * T + t];
            gsum[(size_t)g * T + t] = run;
            run += v;
        }
        if (w == 0) tile_total[t] = total;
    }
}

// one block: exclusive scan of the tile totals (in place -> tile_start) and the tile ranges
// (empty tiles keep (0,0), as the reference's zero-initialised ranges do)
// capacity / overflow: the instance list holds `capacity` entries; if the total exceeds it nothing can be
// placed: every range stays (0,0), *overflow receives the total (0 otherwise) and the place kernel exits.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, uint32_t *__restrict__ tile_start, uint2 *__restrict__ ranges, uint32_t capacity,
                 uint32_t *__restrict__ overflow, uint32_t *__restrict__ tile_order)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t bucket[33];  // tiles per list-length class (class = bit length of the list length)
    const int per = (T + (int)blockDim.x - 1) / (int)blockDim.x;
    const int t0 = threadIdx.x * per, t1 = min(T, t0 + per);
    if (threadIdx.x < 33) bucket[threadIdx.x] = 0;
    uint32_t local = 0;
    for (int t = t0; t < t1; ++t) local += tile_start[t];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (threadIdx.x < 32) warp_sums[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 31) warp_sums[w] = incl;
    for (int t = t0; t < t1; ++t) atomicAdd(&bucket[32 - __clz(tile_start[t])], 1u);
    __syncthreads();
    uint32_t base = 0, total = 0;
    for (int k = 0; k < 32; ++k) {
        if (k < w) base += warp_sums[k];
        total += warp_sums[k];
    }
    // longest lists first: class c starts after all longer classes (the blend kernels hand the squares of the
    // tiles out in this order, so the long walks start first and the short ones fill the tail of the grid)
    uint32_t my_class_base = 0;
    if (threadIdx.x < 33)
        for (int c = 32; c > (int)threadIdx.x; --c) my_class_base += bucket[c];
    __syncthreads();
    if (threadIdx.x < 33) bucket[threadIdx.x] = my_class_base;
    __syncthreads();
    const bool fits = total <= capacity;
    if (threadIdx.x == 0 && overflow) *overflow = fits ? 0u : total;
    uint32_t run = base + incl - local;
    for (int t = t0; t < t1; ++t) {
        const uint32_t tot = tile_start[t];
        tile_start[t] = run;
        ranges[t] = (tot && fits) ? make_uint2(run, run + tot) : make_uint2(0u, 0u);
        run += tot;
        tile_order[atomicAdd(&bucket[32 - __clz(tot)], 1u)] = (uint32_t)t;
    }
}

// ------------------------------------------------------------------------------ C: place
// The CTA first turns its rows of counts into cursors (tile start + CTAs before + warps before), then
// every warp walks its segment in depth order: four Gaussians per step (eight lanes each over the tiles of a
// rectangle) when their rectangles are disjoint, else one Gaussian per step with all lanes over its tiles -- the tiles
// touched in one step are all distinct, so the read-modify-write of the cursors needs no atomics and keeps the order.
__global__ void __launch_bounds__(128)
tile_place_kernel(int P, int T, int gx, int seg, int nseg, const uint4 *__restrict__ srec,
                  const uint32_t *__restrict__ table, const uint32_t *__restrict__ gsum,
                  const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ overflow,
                  uint32_t *__restrict__ out_gidx)
{
    extern __shared__ uint32_t s_cur[];
    if (overflow && *overflow) return;  // the list does not fit its buffer: nothing is placed (uniform)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int sg = blockIdx.x * wpb + w;
    uint32_t *cur = s_cur + (size_t)w * T;
    if (sg < nseg) {
        // eight independent loads in flight per lane (the plain loop was one L2 round trip per 32 tiles)
        const uint32_t *row = table + (size_t)sg * T;
        int t = lane;
        for (; t + 7 * 32 < T; t += 8 * 32) {
            uint32_t v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = row[t + 32 * u];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[t + 32 * u] = v[u];
        }
        for (; t < T; t += 32) cur[t] = row[t];
    } else {
        for (int t = lane; t < T; t += 32) cur[t] = 0;
    }
    __syncthreads();
    const uint32_t *grow = gsum + (size_t)blockIdx.x * T;
    for (int t0 = threadIdx.x; t0 < T; t0 += 4 * blockDim.x) {
        uint32_t run[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * blockDim.x;
            run[u] = t < T ? tile_start[t] + grow[t] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * blockDim.x;
            if (t < T)
                for (int k = 0; k < wpb; ++k) {
                    const uint32_t c = s_cur[(size_t)k * T + t];
                    s_cur[(size_t)k * T + t] = run[u];
                    run[u] += c;
                }
        }
    }
    __syncthreads();
    if (sg >= nseg) return;
    const int s0 = sg * seg, s1 = min(P, s0 + seg);
    uint4 nxt = s0 + lane < s1 ? srec[s0 + lane] : make_uint4(0, 0, 0, 0);
    const int oct = lane >> 3, sub = lane & 7;
    for (int sb = s0; sb < s1; sb += 32) {
        // lane <-> Gaussian: 32 depth-ordered records at once; the next batch is already in flight
        const uint4 rc = nxt;
        const int sn = sb + 32 + lane;
        nxt = sn < s1 ? srec[sn] : make_uint4(0, 0, 0, 0);
        const uint32_t x0 = rc.x & 0xffff, y0 = rc.y & 0xffff, x1 = rc.x >> 16, y1 = rc.y >> 16, wd = x1 - x0;
        const uint32_t n = wd * (y1 - y0), tb = y0 * (uint32_t)gx + x0;
        const uint32_t live = __ballot_sync(0xffffffffu, n != 0);
        if (!live) continue;
        // FOUR Gaussians per step, eight lanes each, whenever the four rectangles of a quad are pairwise disjoint
        // (then no two of them bump the same cursor and the order inside every tile's list is untouched); a quad
        // with an overlap -- or with a rectangle of more than 64 tiles -- is walked one Gaussian at a time by the
        // whole warp, as before.  A lane compares its rectangle with the earlier ones of its quad.
        bool cf = n > 64u;
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            const uint32_t ox = __shfl_up_sync(0xffffffffu, rc.x, d), oy = __shfl_up_sync(0xffffffffu, rc.y, d);
            const uint32_t ox0 = ox & 0xffff, ox1 = ox >> 16, oy0 = oy & 0xffff, oy1 = oy >> 16;
            if ((lane & 3) >= d && n != 0 && ox1 > ox0 && oy1 > oy0 && x0 < ox1 && ox0 < x1 && y0 < oy1 && oy0 < y1) cf = true;
        }
        const uint32_t cmask = __ballot_sync(0xffffffffu, cf);
        for (int q = 0; q < 8; ++q) {
            const uint32_t lq = (live >> (4 * q)) & 0xfu;
            if (!lq) continue;
            if ((cmask >> (4 * q)) & 0xfu) {
                uint32_t m = lq << (4 * q);
                while (m) {
                    const int i = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t ni = __shfl_sync(0xffffffffu, n, i), wi = __shfl_sync(0xffffffffu, wd, i);
                    const uint32_t tbase = __shfl_sync(0xffffffffu, tb, i);
                    const uint32_t mg = __shfl_sync(0xffffffffu, rc.z, i), gi = __shfl_sync(0xffffffffu, rc.w, i);
                    for (uint32_t k = lane; k < ni; k += 32) {
                        const uint32_t qq = mg ? __umulhi(k, mg) : k;  // k / width (exact for k < 2^18)
                        const uint32_t t = tbase + qq * (uint32_t)gx + (k - qq * wi);
                        const uint32_t slot = cur[t];
                        cur[t] = slot + 1;
                        out_gidx[slot] = gi;
                    }
                    __syncwarp();
                }
            } else {
                const int i = 4 * q + oct;
                const uint32_t ni = __shfl_sync(0xffffffffu, n, i), wi = __shfl_sync(0xffffffffu, wd, i);
                const uint32_t tbase = __shfl_sync(0xffffffffu, tb, i);
                const uint32_t mg = __shfl_sync(0xffffffffu, rc.z, i), gi = __shfl_sync(0xffffffffu, rc.w, i);
                for (uint32_t k = sub; k < ni; k += 8) {
                    const uint32_t qq = mg ? __umulhi(k, mg) : k;
                    const uint32_t t = tbase + qq * (uint32_t)gx + (k - qq * wi);
                    const uint32_t slot = cur[t];
                    cur[t] = slot + 1;
                    out_gidx[slot] = gi;
                }
                __syncwarp();
            }
        }
    }
}

// ------------------------------------------------------------------------------ inspection helper
// sorted tile ids rebuilt from the ranges (the placement path never materialises them)
__global__ void fill_tiles_kernel(int T, const uint2 *__restrict__ ranges, uint32_t *__restrict__ sorted_tiles)
{
    const int t = blockIdx.x;
    const uint2 r = ranges[t];
    for (uint32_t j = r.x + threadIdx.x; j < r.y; j += blockDim.x) sorted_tiles[j] = (uint32_t)t;
}

int launch_fill_tiles(int T, const uint2 *ranges, uint32_t *sorted_tiles, cudaStream_t s)
{
    if (T <= 0) return 0;
    fill_tiles_kernel<<<T, 128, 0, s>>>(T, ranges, sorted_tiles);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

int launch_tile_placement(const PlacePlan &pl, int P, int T, int gx, const uint32_t *order_a, const uint32_t *order_b,
                          const uint32_t *stat, const uint2 *rect,
                          uint4 *srec, uint32_t *table, uint32_t *gsum, uint32_t *tile_start, uint2 *ranges,
                          uint32_t *out_gidx, int64_t capacity, uint32_t *overflow, uint32_t *tile_order, cudaStream_t s)
{
    if (once_per_device(ONCE_PLACE)) {
        DMGS_CUDA(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        DMGS_CUDA(cudaFuncSetAttribute(tile_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        // the plan counts on ~200 KB of shared memory per SM: ask for the largest carve-out
        DMGS_CUDA(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        DMGS_CUDA(cudaFuncSetAttribute(tile_place_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    const int threads = pl.wpb * 32;
    const int blocks = pl.groups;
    sorted_rect_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, order_a, order_b, stat, rect, srec);
    const size_t count_smem = (size_t)pl.wpb * (gx + 1) * (T / gx + 1) * sizeof(uint32_t);  // the difference images
    tile_count_kernel<<<blocks, threads, count_smem, s>>>(P, T, gx, pl.seg, pl.nseg, srec, table, gsum);
    group_scan_kernel<<<(T + 31) / 32, GS_WARPS * 32, 0, s>>>(T, pl.groups, gsum, tile_start);
    tile_scan_kernel<<<1, 1024, 0, s>>>(T, tile_start, ranges, (uint32_t)capacity, overflow, tile_order);
    tile_place_kernel<<<blocks, threads, pl.smem, s>>>(P, T, gx, pl.seg, pl.nseg, srec, table, gsum, tile_start, overflow, out_gidx);
    DMGS_CUDA(cudaGetLastError());
    count_launches(5);
    return 0;
}

}  // namespace dmgs
