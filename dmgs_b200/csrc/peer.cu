// peer.cu -- all-reduce (sum) of the flat per-Gaussian gradient buffer over NVLink peer memory, for the
// view-partitioned training step (the collective follows the last view's
// per-Gaussian backward on every rank).
//
// Every rank owns the slice [rank * n/N, (rank+1) * n/N) of the buffer.  One kernel per rank
//   * reads its slice from ALL ranks' buffers (its own from HBM, the others through NVLink peer
//     mappings, or -- when the buffers are bound to an NVSwitch multicast object -- with ONE
//     multimem.ld_reduce that performs the addition inside the switch),
//   * scales the sum (view averaging), and
//   * writes the result back into the same slice of every rank's buffer (peer stores, or one
//     multimem.st that the switch broadcasts).
// During the kernel only rank r touches slice r of anybody's buffer, so the exchange is in place and
// needs no staging copy; the caller brackets the launch with two device-side barriers over the
// symmetric-memory signal pads (all accumulators complete before, all slices delivered after).
// NVLink traffic per rank: (N-1)/N * n * 4 B in each direction with peer loads/stores, n/N * 4 B in
// each direction with multimem -- against 2 (N-1)/N * n * 4 B each way for a ring all-reduce.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

struct PeerPtrs {
    float *p[DMGS_MAX_PEERS];
};

// Peer buffers are written by kernels that finished before the barrier preceding this launch and are
// homed in their owner's L2: plain cache-global accesses (no L1 allocation) observe them; no
// system-scope ordering is needed inside the kernel.
__device__ __forceinline__ float4 ld_sys(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st_sys(float *p, float4 v) { __stcg(reinterpret_cast<float4 *>(p), v); }

template <int WORLD>
__global__ void __launch_bounds__(256)
allreduce_peer_kernel(const __grid_constant__ PeerPtrs peers, int rank, long long begin4, long long end4, float scale)
{
    // two 16-byte groups per thread and iteration, all 2 * WORLD loads issued before the first store:
    // NVLink round trips are ~2-3 us, so bytes in flight decide the bandwidth
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < end4; q += 2 * stride) {
        const long long q1 = q + stride;
        const bool two = q1 < end4;
        float4 v[WORLD], w[WORLD];
#pragma unroll
        for (int k = 0; k < WORLD; ++k) {
            const float *src = peers.p[(rank + k) % WORLD];
            v[k] = ld_sys(src + 4 * q);
            if (two) w[k] = ld_sys(src + 4 * q1);
        }
        float4 a = v[0], b = two ? w[0] : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 1; k < WORLD; ++k) {
            a.x += v[k].x; a.y += v[k].y; a.z += v[k].z; a.w += v[k].w;
            if (two) { b.x += w[k].x; b.y += w[k].y; b.z += w[k].z; b.w += w[k].w; }
        }
        a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
        b.x *= scale; b.y *= scale; b.z *= scale; b.w *= scale;
#pragma unroll
        for (int k = 0; k < WORLD; ++k) {
            float *dst = peers.p[(rank + k) % WORLD];
            st_sys(dst + 4 * q, a);
            if (two) st_sys(dst + 4 * q1, b);
        }
    }
}

__device__ __forceinline__ float4 mm_ld_reduce(float *p)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st(float *p, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// U 16-byte groups per thread and iteration: all U in-switch reductions are issued before the first multicast store
template <int U>
__global__ void __launch_bounds__(1024)
allreduce_multimem_kernel(float *mc, long long begin4, long long end4, float scale)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q < end4; q += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < end4) v[u] = mm_ld_reduce(mc + 4 * (q + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (q + u * stride < end4) {
                v[u].x *= scale; v[u].y *= scale; v[u].z *= scale; v[u].w *= scale;
                mm_st(mc + 4 * (q + u * stride), v[u]);
            }
    }
}

// launch shape of the multimem kernel: {CTAs per SM, threads per CTA, groups in flight per thread}.  Measured on one
// 8 x B200 box (profiles/r2_allreduce_sweep.jsonl): the exchange is bound by what the fabric sustains (0.60-0.66 ms
// for 248 MB whatever the shape; running peer loads/stores beside the switch reduction on a share of the buffer
// does not help either), a small grid with eight reductions in flight per thread is the best of them.
// DMGS_AR_SHAPE="ctas,threads,unroll" overrides (experiments)
static void ar_shape(int *ctas_per_sm, int *threads, int *unroll)
{
    int cc = 1, tt = 256, uu = 8;
    if (const char *e = getenv("DMGS_AR_SHAPE")) sscanf(e, "%d,%d,%d", &cc, &tt, &uu);
    if (cc < 1) cc = 1;
    if (tt < 32 || tt > 1024 || (tt & 31)) tt = 256;
    if (uu != 1 && uu != 2 && uu != 4 && uu != 8) uu = 4;
    *ctas_per_sm = cc; *threads = tt; *unroll = uu;
}

int launch_allreduce_peer(int64_t n, int world, int rank, const void *const *peer_ptrs_host, void *multicast_ptr,
                          float scale, cudaStream_t s)
{
    if (world < 1 || world > DMGS_MAX_PEERS || rank < 0 || rank >= world) { set_error("allreduce_peer: bad rank %d / world %d", rank, world); return -13; }
    if (n < 0 || (n & 3)) { set_error("allreduce_peer: element count must be a multiple of 4"); return -13; }
    const long long n4 = n / 4, per = (n4 + world - 1) / world;
    const long long begin4 = (long long)rank * per, end4 = begin4 + per < n4 ? begin4 + per : n4;
    if (end4 <= begin4) return 0;
    int cps, threads, unroll;
    ar_shape(&cps, &threads, &unroll);
    long long blocks = (end4 - begin4 + 255) / 256;
    long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (multicast_ptr) {
        float *mc = reinterpret_cast<float *>(multicast_ptr);
        blocks = (end4 - begin4 + threads - 1) / threads;
        cap = (long long)num_sms() * cps;
        if (blocks > cap) blocks = cap;
        const unsigned g = (unsigned)blocks;
        switch (unroll) {
        case 1: allreduce_multimem_kernel<1><<<g, threads, 0, s>>>(mc, begin4, end4, scale); break;
        case 2: allreduce_multimem_kernel<2><<<g, threads, 0, s>>>(mc, begin4, end4, scale); break;
        case 8: allreduce_multimem_kernel<8><<<g, threads, 0, s>>>(mc, begin4, end4, scale); break;
        default: allreduce_multimem_kernel<4><<<g, threads, 0, s>>>(mc, begin4, end4, scale); break;
        }
    } else {
        PeerPtrs pp;
        for (int k = 0; k < DMGS_MAX_PEERS; ++k) {
            pp.p[k] = k < world ? reinterpret_cast<float *>(const_cast<void *>(peer_ptrs_host[k])) : nullptr;
            if (k < world && (!pp.p[k] || ((uintptr_t)pp.p[k] & 15))) { set_error("allreduce_peer: peer pointer %d NULL or unaligned", k); return -13; }
        }
        const unsigned g = (unsigned)blocks;
        switch (world) {
        case 1: allreduce_peer_kernel<1><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 2: allreduce_peer_kernel<2><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 3: allreduce_peer_kernel<3><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 4: allreduce_peer_kernel<4><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 5: allreduce_peer_kernel<5><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 6: allreduce_peer_kernel<6><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        case 7: allreduce_peer_kernel<7><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        default: allreduce_peer_kernel<8><<<g, 256, 0, s>>>(pp, rank, begin4, end4, scale); break;
        }
    }
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
