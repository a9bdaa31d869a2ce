// cub_stage.cu -- LIBRARY STAGE COMPARATOR, not the reference and not part of the product.
// Times exactly the CUB calls upstream's rasteriser makes for binning: an InclusiveSum over the
// P tiles-touched counts and a DeviceRadixSort::SortPairs of R (64-bit key = tile << 32 | depth bits, 32-bit value)
// on bits [0, 32 + ceil(log2 T)), on synthetic keys of the bench's shape, so that the placement path's stage times
// (bench.py "preprocess_sort_scan" minus preprocess, and "binning") have a library number next to them.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/_cub_stage scripts/cub_stage.cu
//   scripts/_cub_stage <P> <R> <tiles> [iters]      -> one JSON line
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(err__)); return 1; } } while (0)

__global__ void make_keys(uint64_t *keys, uint32_t *vals, int64_t R, uint32_t T, uint32_t P)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    const uint32_t tile = (uint32_t)(h % T);
    const float depth = 2.0f + 4.0f * (float)((h >> 40) & 0xFFFFF) / 1048576.0f;  // camera-space depths in [2, 6)
    keys[i] = ((uint64_t)tile << 32) | (uint64_t)__float_as_uint(depth);
    vals[i] = (uint32_t)((h >> 13) % P);
}
__global__ void make_counts(uint32_t *c, int P)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) c[i] = (uint32_t)((i * 2654435761u) >> 29);  // 0..7 tiles per Gaussian
}
__global__ void ranges_kernel(int64_t R, const uint64_t *keys, uint2 *ranges)
{  // upstream's identifyTileRanges: one thread per sorted instance
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[t].x = 0;
    else {
        const uint32_t p = (uint32_t)(keys[i - 1] >> 32);
        if (p != t) { ranges[p].y = (uint32_t)i; ranges[t].x = (uint32_t)i; }
    }
    if (i == R - 1) ranges[t].y = (uint32_t)R;
}

int main(int argc, char **argv)
{
    const int P = argc > 1 ? atoi(argv[1]) : 1000000;
    const int64_t R = argc > 2 ? atoll(argv[2]) : 7376946;
    const uint32_t T = argc > 3 ? (uint32_t)atoi(argv[3]) : 2500;
    const int iters = argc > 4 ? atoi(argv[4]) : 20;
    int tbits = 0;
    while ((1u << tbits) < T) ++tbits;
    uint64_t *k0, *k1; uint32_t *v0, *v1, *cnt, *off; uint2 *ranges;
    CK(cudaMalloc(&k0, R * 8)); CK(cudaMalloc(&k1, R * 8)); CK(cudaMalloc(&v0, R * 4)); CK(cudaMalloc(&v1, R * 4));
    CK(cudaMalloc(&cnt, (size_t)P * 4)); CK(cudaMalloc(&off, (size_t)P * 4)); CK(cudaMalloc(&ranges, (size_t)T * 8));
    make_counts<<<(P + 255) / 256, 256>>>(cnt, P);
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, k0, k1, v0, v1, R, 0, 32 + tbits);
    cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, cnt, off, P);
    void *tmp; CK(cudaMalloc(&tmp, std::max(tmp_sort, tmp_scan)));
    cudaEvent_t e[4];
    for (auto &x : e) CK(cudaEventCreate(&x));
    std::vector<float> ts, tc, tr;
    for (int it = 0; it < iters + 3; ++it) {
        make_keys<<<(unsigned)((R + 255) / 256), 256>>>(k0, v0, R, T, (uint32_t)P);  // fresh unsorted input (also evicts L2)
        CK(cudaMemset(ranges, 0, (size_t)T * 8));
        CK(cudaEventRecord(e[0]));
        cub::DeviceScan::InclusiveSum(tmp, tmp_scan, cnt, off, P);
        CK(cudaEventRecord(e[1]));
        cub::DeviceRadixSort::SortPairs(tmp, tmp_sort, k0, k1, v0, v1, R, 0, 32 + tbits);
        CK(cudaEventRecord(e[2]));
        ranges_kernel<<<(unsigned)((R + 255) / 256), 256>>>(R, k1, ranges);
        CK(cudaEventRecord(e[3]));
        CK(cudaDeviceSynchronize());
        if (it < 3) continue;
        float a, b, c;
        cudaEventElapsedTime(&a, e[0], e[1]); cudaEventElapsedTime(&b, e[1], e[2]); cudaEventElapsedTime(&c, e[2], e[3]);
        tc.push_back(a); ts.push_back(b); tr.push_back(c);
    }
    auto med = [](std::vector<float> v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    printf("{\"what\": \"CUB library stage comparator (not the reference): DeviceScan::InclusiveSum over P counts, "
           "DeviceRadixSort::SortPairs of R 64-bit keys / 32-bit values on %d bits, tile-range pass\", \"P\": %d, \"R\": %lld, "
           "\"tiles\": %u, \"iters\": %d, \"scan_ms\": %.4f, \"sort_ms\": %.4f, \"ranges_ms\": %.4f, \"total_ms\": %.4f, "
           "\"sort_GBps_24B_per_pass_model\": %.1f}\n",
           32 + tbits, P, (long long)R, T, iters, med(tc), med(ts), med(tr), med(tc) + med(ts) + med(tr),
           (double)R * 24.0 * ((32 + tbits + 7) / 8) / (med(ts) * 1e-3) / 1e9);
    return 0;
}
