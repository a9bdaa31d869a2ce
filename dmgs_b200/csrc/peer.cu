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

#include "adam_math.cuh"
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

// ------------------------------------------------------------------------------ exchange fused with the optimiser
// The training step's tail is "all-reduce the gradients, then every rank runs the SAME Adam step on its replica":
// N redundant optimiser passes over p, g, m, v.  Fused: rank r owns slice r of every parameter tensor.  One kernel
// per rank sums slice r of all ranks' gradients (in the switch: multimem.ld_reduce; or peer loads), applies Adam with
// ITS shard of the optimiser state (exp_avg / exp_avg_sq exist once per box, not once per rank) and broadcasts the
// new PARAMETERS to all replicas (multimem.st / peer stores) -- reduce-scatter + update + all-gather in one pass, the
// same NVLink volume as the all-reduce alone, and the separate Adam launch (0.34 ms for 59 M parameters) is gone.
// Parameters and gradients live in two flat symmetric buffers with the same field layout (16-byte aligned fields).
struct XSeg {
    long long goff4, poff4;    // first 16-byte group of the tensor in the flat gradient / parameter buffer
    long long begin4, end4;    // the groups of the tensor this rank owns
    float *m, *v;              // this rank's shard of the state, indexed from begin4
    float step_lo, step_hi;
    int period, split;
};
struct XArgs {
    XSeg seg[DMGS_ADAM_MAX_SEGMENTS];
    AdamConsts c;
    float *gmc, *pmc;          // multicast mappings (NULL: peer loads / stores)
    float *plocal;
    int rank;
};

template <int WORLD, bool MC>
__global__ void __launch_bounds__(256)
adam_exchange_kernel(const __grid_constant__ XArgs a, const __grid_constant__ PeerPtrs gpeers, const __grid_constant__ PeerPtrs ppeers)
{
    const XSeg &sg = a.seg[blockIdx.y];
    constexpr int U = 4;  // 16-byte groups in flight per thread: the switch round trip is microseconds
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q0 = sg.begin4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < sg.end4; q0 += U * stride) {
        float4 g[U], p[U], m[U], v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = q0 + u * stride;
            if (q < sg.end4) {
                if (MC) {
                    g[u] = mm_ld_reduce(a.gmc + 4 * (sg.goff4 + q));
                } else {
                    g[u] = ld_sys(gpeers.p[a.rank] + 4 * (sg.goff4 + q));
#pragma unroll
                    for (int k = 1; k < WORLD; ++k) {
                        const float4 t = ld_sys(gpeers.p[(a.rank + k) % WORLD] + 4 * (sg.goff4 + q));
                        g[u].x += t.x; g[u].y += t.y; g[u].z += t.z; g[u].w += t.w;
                    }
                }
                p[u] = *reinterpret_cast<const float4 *>(a.plocal + 4 * (sg.poff4 + q));
                m[u] = *reinterpret_cast<const float4 *>(sg.m + 4 * (q - sg.begin4));
                v[u] = *reinterpret_cast<const float4 *>(sg.v + 4 * (q - sg.begin4));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long q = q0 + u * stride;
            if (q < sg.end4) {
                const uint32_t ph = sg.period > 0 ? (uint32_t)((unsigned long long)(4 * q) % (unsigned)sg.period) : 0u;
                float *pp = &p[u].x, *mp = &m[u].x, *vp = &v[u].x;
                const float *gp = &g[u].x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float st = (sg.period > 0 && (ph + k) % (uint32_t)sg.period >= (uint32_t)sg.split) ? sg.step_hi : sg.step_lo;
                    adam_math(a.c, st, pp[k], gp[k], mp[k], vp[k]);
                }
                *reinterpret_cast<float4 *>(sg.m + 4 * (q - sg.begin4)) = m[u];
                *reinterpret_cast<float4 *>(sg.v + 4 * (q - sg.begin4)) = v[u];
                if (MC) {
                    mm_st(a.pmc + 4 * (sg.poff4 + q), p[u]);
                } else {
#pragma unroll
                    for (int k = 0; k < WORLD; ++k) st_sys(ppeers.p[(a.rank + k) % WORLD] + 4 * (sg.poff4 + q), p[u]);
                }
            }
        }
    }
}

void adam_exchange_shard(int64_t n, int world, int rank, int64_t *begin4, int64_t *end4)
{
    const int64_t n4 = (n + 3) / 4, per = (n4 + world - 1) / world;
    int64_t b = (int64_t)rank * per, e = b + per;
    if (b > n4) b = n4;
    if (e > n4) e = n4;
    *begin4 = b; *end4 = e;
}

int launch_adam_exchange(int world, int rank, int nseg, const dmgs_adam_xsegment *segs, const void *const *grad_peers_host,
                         void *grad_multicast, const void *const *param_peers_host, void *param_multicast, double beta1,
                         double beta2, double eps, int64_t step, float grad_scale, cudaStream_t s)
{
    if (world < 1 || world > DMGS_MAX_PEERS || rank < 0 || rank >= world) { set_error("adam_exchange: bad rank %d / world %d", rank, world); return -13; }
    if (nseg < 1 || nseg > DMGS_ADAM_MAX_SEGMENTS) { set_error("adam_exchange: 1..%d segments per call, got %d", DMGS_ADAM_MAX_SEGMENTS, nseg); return -12; }
    if (step < 1) { set_error("adam_exchange: step counts from 1"); return -12; }
    if ((grad_multicast != nullptr) != (param_multicast != nullptr)) { set_error("adam_exchange: both or neither buffer needs a multicast mapping"); return -13; }
    XArgs a;
    memset(&a, 0, sizeof(a));
    PeerPtrs gp, pp;
    for (int k = 0; k < DMGS_MAX_PEERS; ++k) {
        gp.p[k] = k < world ? reinterpret_cast<float *>(const_cast<void *>(grad_peers_host[k])) : nullptr;
        pp.p[k] = k < world ? reinterpret_cast<float *>(const_cast<void *>(param_peers_host[k])) : nullptr;
        if (k < world && (!gp.p[k] || !pp.p[k] || (((uintptr_t)gp.p[k] | (uintptr_t)pp.p[k]) & 15))) { set_error("adam_exchange: peer pointer %d NULL or unaligned", k); return -13; }
    }
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    long long most = 0;
    for (int k = 0; k < nseg; ++k) {
        const dmgs_adam_xsegment &u = segs[k];
        if (u.n < 0 || (u.grad_offset & 3) || (u.param_offset & 3) || u.grad_offset < 0 || u.param_offset < 0) { set_error("adam_exchange: segment %d: offsets must be non-negative multiples of 4 floats", k); return -12; }
        int64_t b4, e4;
        adam_exchange_shard(u.n, world, rank, &b4, &e4);
        if (e4 > b4 && (!u.exp_avg_shard || !u.exp_avg_sq_shard || (((uintptr_t)u.exp_avg_shard | (uintptr_t)u.exp_avg_sq_shard) & 15))) { set_error("adam_exchange: segment %d: state shard NULL or unaligned", k); return -12; }
        XSeg &d = a.seg[k];
        d.goff4 = u.grad_offset / 4; d.poff4 = u.param_offset / 4;
        d.begin4 = b4; d.end4 = e4;
        d.m = u.exp_avg_shard; d.v = u.exp_avg_sq_shard;
        d.step_lo = (float)(u.lr / bc1);
        d.step_hi = (float)((u.period > 0 ? u.lr_hi : u.lr) / bc1);
        d.period = u.period; d.split = u.split;
        if (e4 - b4 > most) most = e4 - b4;
    }
    a.c.b2 = (float)beta2;
    a.c.one_minus_b1 = (float)(1.0 - beta1);
    a.c.one_minus_b2 = (float)(1.0 - beta2);
    a.c.sqrt_bc2 = (float)sqrt(bc2);
    a.c.eps = (float)eps; a.c.grad_scale = grad_scale;
    a.gmc = reinterpret_cast<float *>(grad_multicast); a.pmc = reinterpret_cast<float *>(param_multicast);
    a.plocal = pp.p[rank]; a.rank = rank;
    if (most == 0) return 0;
    long long blocks = (most + 255) / 256;
    const long long cap = (long long)num_sms() * 2;
    if (blocks > cap) blocks = cap;
    const dim3 grid((unsigned)blocks, (unsigned)nseg);
    if (a.gmc) {
        adam_exchange_kernel<1, true><<<grid, 256, 0, s>>>(a, gp, pp);
    } else {
        switch (world) {
        case 1: adam_exchange_kernel<1, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 2: adam_exchange_kernel<2, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 3: adam_exchange_kernel<3, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 4: adam_exchange_kernel<4, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 5: adam_exchange_kernel<5, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 6: adam_exchange_kernel<6, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        case 7: adam_exchange_kernel<7, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        default: adam_exchange_kernel<8, false><<<grid, 256, 0, s>>>(a, gp, pp); break;
        }
    }
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
