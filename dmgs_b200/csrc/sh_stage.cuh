// sh_stage.cuh -- TMA (cp.async.bulk) staging of per-Gaussian SH rows through shared memory.
//
// A Gaussian's SH coefficients are 48 contiguous floats (192 B) in both layouts ([P,16,3] and
// [P,3,16]).  One thread per Gaussian reading its own row makes every load instruction touch 32
// different lines; instead each lane issues ONE bulk copy of its 192-byte row into a padded
// shared-memory row (stride 52 floats: 16-byte aligned and conflict-free for 128-bit reads), the
// warp waits on one mbarrier, and the row is read back with 12 LDS.128.  Gradient rows go the
// other way: registers -> shared row -> cp.async.bulk store, or cp.reduce.async.bulk (.add.f32)
// when gradients of several views are accumulated -- the read-modify-write then happens in L2.
// Rows of culled Gaussians are never fetched.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmgs {

constexpr int SH_ROW_FLOATS = 48;                   // 16 coefficients x 3 channels
constexpr int SH_ROW_BYTES = SH_ROW_FLOATS * 4;     // 192
constexpr int SH_ROW_STRIDE = 52;                   // floats
#ifndef PRE_BLK
#define PRE_BLK 128   /* threads per CTA of the per-Gaussian kernels: 128 (8 K registers, 26 KB of staged rows) fits beside the blend and placement CTAs of other views where 256 did not (H0: 1741 -> 1766 frames/s, alone 3 % faster) */
#endif
constexpr int SH_STAGE_THREADS = PRE_BLK;
constexpr int SH_STAGE_SMEM = SH_STAGE_THREADS * SH_ROW_STRIDE * 4 + (SH_STAGE_THREADS / 32) * 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();  // a lost transaction must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_reduce_add_f32(void *dst, uint32_t src_smem, uint32_t bytes)
{
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                 ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct ShStage {
    float *row;       // this lane's padded shared-memory row
    uint32_t bar;     // this warp's mbarrier (shared address)
    uint32_t issued;  // lanes of the warp with a copy in flight

    __device__ __forceinline__ void init(unsigned char *dsm)
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        row = reinterpret_cast<float *>(dsm) + (size_t)threadIdx.x * SH_ROW_STRIDE;
        bar = smem_u32(dsm + SH_STAGE_THREADS * SH_ROW_STRIDE * 4 + w * 8);
        if (lane == 0) mbar_init(bar, 1);
        __syncwarp();
        issued = 0;
    }
    // warp-collective: lanes with need==true fetch their 192-byte row
    __device__ __forceinline__ void load(bool need, const float *src)
    {
        issued = __ballot_sync(0xffffffffu, need);
        if (issued) {
            if ((threadIdx.x & 31) == 0) mbar_expect_tx(bar, (uint32_t)__popc(issued) * SH_ROW_BYTES);
            __syncwarp();
            if (need) bulk_load(smem_u32(row), src, SH_ROW_BYTES, bar);
        }
    }
    __device__ __forceinline__ void wait()
    {
        if (issued) mbar_wait(bar, 0);
    }
    __device__ __forceinline__ void read(float *r) const
    {
#pragma unroll
        for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) {
            const float4 v = reinterpret_cast<const float4 *>(row)[j];
            r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
    }
    // writes this lane's row to global memory (store or L2 reduce-add); call flush() before exit
    __device__ __forceinline__ void write_out(const float *r, float *dst, bool accumulate)
    {
#pragma unroll
        for (int j = 0; j < SH_ROW_FLOATS / 4; ++j)
            reinterpret_cast<float4 *>(row)[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        fence_proxy_async();
        if (accumulate) bulk_reduce_add_f32(dst, smem_u32(row), SH_ROW_BYTES);
        else bulk_store(dst, smem_u32(row), SH_ROW_BYTES);
        bulk_commit();
    }
    // same, for a row that was already built in place in shared memory
    __device__ __forceinline__ void flush_row(float *dst, bool accumulate)
    {
        fence_proxy_async();
        if (accumulate) bulk_reduce_add_f32(dst, smem_u32(row), SH_ROW_BYTES);
        else bulk_store(dst, smem_u32(row), SH_ROW_BYTES);
        bulk_commit();
    }
    __device__ __forceinline__ void flush() { bulk_wait_read_all(); }
};

}  // namespace dmgs
