// frustum.cu -- view-frustum test of query points / face centroids and ordered compaction of the
// visible faces: the step right before the binding in stage 3 and in
// the COLMAP stage 2.
//
// Replaces `in_frustum` (scene/gaussian_geo_model_finetune.py:33-48; the COLMAP variant with the
// cube offset and the screen pieces, scene/gaussian_geo_model_mlp_flex_colmap.py:32-76) and its use
// at finetune.py:405-409: `query_p = verts[faces].mean(dim=1)`, `mask = in_frustum(P, query_p)`,
// `faces = faces[mask]` -- a gather, a mean, a matmul, five element-wise kernels, a boolean-index
// compaction (nonzero + gather) in PyTorch; here one mask kernel + a block scan + one scatter.
//
// Arithmetic (fp32, matched by oracle/splat_oracle.c:orc_in_frustum bit for bit):
//   centroid = ((v0 + v1) + v2) / 3;  p_j = fma(z, M[2][j], fma(y, M[1][j], x * M[0][j])) + row_j,
//   row = M[3] (+ sum_i M[i] * cube_len/2 for the COLMAP variant); w = p_3 + 1e-6; q = p_xyz / w (IEEE);
//   visible = lo_j < q_j < hi_j for j = 0..2 and w > 0, with (lo, hi) = (-1.05, 1.05) or the piece's box.
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

struct FrustumArgs {
    float M[12];   // rows 0..2 of the [4,4] right-multiplied projection matrix
    float row[4];  // translation row (with the cube offset folded in)
    float lo[3], hi[3];
};

__device__ __forceinline__ bool frustum_test(const FrustumArgs &a, float x, float y, float z)
{
    float p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = fma_(z, a.M[8 + j], fma_(y, a.M[4 + j], x * a.M[j])) + a.row[j];
    const float w = p[3] + 1e-6f;
    bool ok = w > 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float q = p[j] / w;
        ok = ok && (q > a.lo[j]) && (q < a.hi[j]);
    }
    return ok;
}

// mode 0: points [N,3]; mode 1: centroids of faces [N,3] (int64) over verts
template <int FACES>
__global__ void __launch_bounds__(256)
frustum_mask_kernel(int64_t N, const __grid_constant__ FrustumArgs a, const float *__restrict__ pts,
                    const int64_t *__restrict__ faces, uint8_t *__restrict__ mask, uint32_t *__restrict__ block_count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool vis = false;
    if (i < N) {
        float x, y, z;
        if (FACES) {
            const int64_t f0 = faces[3 * i], f1 = faces[3 * i + 1], f2 = faces[3 * i + 2];
            x = ((pts[3 * f0] + pts[3 * f1]) + pts[3 * f2]) / 3.0f;
            y = ((pts[3 * f0 + 1] + pts[3 * f1 + 1]) + pts[3 * f2 + 1]) / 3.0f;
            z = ((pts[3 * f0 + 2] + pts[3 * f1 + 2]) + pts[3 * f2 + 2]) / 3.0f;
        } else {
            x = pts[3 * i]; y = pts[3 * i + 1]; z = pts[3 * i + 2];
        }
        vis = frustum_test(a, x, y, z);
        mask[i] = vis ? 1 : 0;
    }
    if (block_count) {
        const int c = __syncthreads_count(vis);
        if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
    }
}

// one block: exclusive scan of the per-block counts (in place) and the total
__global__ void __launch_bounds__(1024)
frustum_scan_kernel(int nb, uint32_t *__restrict__ block_count, int32_t *__restrict__ total)
{
    __shared__ uint32_t warp_sums[32];
    const int per = (nb + 1023) / 1024;
    const int b0 = threadIdx.x * per, b1 = min(nb, b0 + per);
    uint32_t local = 0;
    for (int b = b0; b < b1; ++b) local += block_count[b];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
    for (int k = 0; k < 32; ++k) {
        if (k < w) base += warp_sums[k];
        tot += warp_sums[k];
    }
    uint32_t run = base + incl - local;
    for (int b = b0; b < b1; ++b) {
        const uint32_t c = block_count[b];
        block_count[b] = run;
        run += c;
    }
    if (threadIdx.x == 0) *total = (int32_t)tot;
}

// faces_out[rank of i among visible faces] = faces[i], order preserved (what faces[mask] returns)
__global__ void __launch_bounds__(256)
frustum_compact_kernel(int64_t N, const uint8_t *__restrict__ mask, const int64_t *__restrict__ faces,
                       const uint32_t *__restrict__ block_offset, int64_t *__restrict__ faces_out,
                       int32_t *__restrict__ index_out)
{
    __shared__ uint32_t warp_base[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vis = i < N && mask[i];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t bal = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) warp_base[w] = __popc(bal);
    __syncthreads();
    uint32_t base = block_offset[blockIdx.x];
    for (int k = 0; k < w; ++k) base += warp_base[k];
    if (vis) {
        const uint32_t r = base + __popc(bal & ((1u << lane) - 1u));
        if (faces_out) {
            faces_out[3 * (size_t)r] = faces[3 * i];
            faces_out[3 * (size_t)r + 1] = faces[3 * i + 1];
            faces_out[3 * (size_t)r + 2] = faces[3 * i + 2];
        }
        if (index_out) index_out[r] = (int32_t)i;
    }
}

size_t frustum_scratch_bytes(int64_t N) { return align_up((size_t)((N + 255) / 256 + 1) * sizeof(uint32_t)); }

int launch_frustum(int64_t N, const float *proj16_host, float cube_len, int has_cube, int piece_id, int n_piece,
                   const float *pts, const int64_t *faces, uint8_t *mask, int64_t *faces_out, int32_t *index_out,
                   int32_t *count_out, void *scratch, cudaStream_t s)
{
    FrustumArgs a;
    for (int i = 0; i < 12; ++i) a.M[i] = proj16_host[i];
    for (int j = 0; j < 4; ++j) {
        float r = proj16_host[12 + j];
        if (has_cube) {
            // (proj_matrix_3x3 * (cube_len / 2.)).sum(dim=0) added to the translation row (colmap.py:41)
            const float h = cube_len / 2.0f;
            const float sum = (proj16_host[j] * h + proj16_host[4 + j] * h) + proj16_host[8 + j] * h;
            r = r + sum;
        }
        a.row[j] = r;
    }
    // screen pieces of the COLMAP variant (colmap.py:46-72); piece_id < 0 = the whole screen
    float lo[3] = {-1.0f, -1.0f, -1.0f}, hi[3] = {1.0f, 1.0f, 1.0f};
    if (piece_id >= 0) {
        if (n_piece == 2 && piece_id < 2) {
            if (piece_id == 0) hi[0] = 0.0f; else lo[0] = 0.0f;
        } else if (n_piece == 4 && piece_id < 4) {
            if (piece_id & 1) lo[0] = 0.0f; else hi[0] = 0.0f;
            if (piece_id & 2) lo[1] = 0.0f; else hi[1] = 0.0f;
        } else {
            set_error("in_frustum: unsupported piece_id %d / n_piece %d", piece_id, n_piece);
            return -11;
        }
    }
    for (int j = 0; j < 3; ++j) { a.lo[j] = lo[j] - 0.05f; a.hi[j] = hi[j] + 0.05f; }
    if (N <= 0) {
        if (count_out) DMGS_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), s));
        return 0;
    }
    const int nb = (int)((N + 255) / 256);
    const bool compact = count_out != nullptr;
    uint32_t *bc = compact ? reinterpret_cast<uint32_t *>(scratch) : nullptr;
    if (faces)
        frustum_mask_kernel<1><<<nb, 256, 0, s>>>(N, a, pts, faces, mask, bc);
    else
        frustum_mask_kernel<0><<<nb, 256, 0, s>>>(N, a, pts, nullptr, mask, bc);
    int n = 1;
    if (compact) {
        frustum_scan_kernel<<<1, 1024, 0, s>>>(nb, bc, count_out);
        frustum_compact_kernel<<<nb, 256, 0, s>>>(N, mask, faces, bc, faces ? faces_out : nullptr, index_out);
        n = 3;
    }
    DMGS_CUDA(cudaGetLastError());
    count_launches(n);
    return 0;
}

}  // namespace dmgs
