// binding.cu -- fused mesh-face -> Gaussian binding, forward and backward.
// Replaces the ~25 eager PyTorch kernels of scene/gaussian_geo_model_mlp_flex.py:267-311
// (face frame, barycentric means, affine cov3D_L) and :370-385 (get_covariance_dyn: R L (R L)^T,
// strip_symmetric, broadcast to the k Gaussians of the face) with one pass per direction.
// One thread per face; HBM-bound: reads 24 B of indices + 36 B of vertices, writes k*(12+24) B.
#include "common.cuh"
#include "kernels.cuh"
#include "bind_math.cuh"

namespace dmgs {

__global__ void __launch_bounds__(256)
bind_fwd_kernel(int64_t F, int k, const float *__restrict__ verts, const int64_t *__restrict__ faces,
                const float *__restrict__ bc, float rad_base, float thin_z, const float *__restrict__ g_ptr, int adaptive,
                float *__restrict__ xyz, float *__restrict__ cov6, float *__restrict__ rot_t2w)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float g = g_ptr ? __ldg(g_ptr) : 1.0f;
    const int64_t i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    float v0[3], v1[3], v2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { v0[c] = verts[3 * i0 + c]; v1[c] = verts[3 * i1 + c]; v2[c] = verts[3 * i2 + c]; }
    if (xyz) {
        for (int j = 0; j < k; ++j) {
            const float b0 = bc[3 * j], b1 = bc[3 * j + 1], b2 = bc[3 * j + 2];
#pragma unroll
            for (int c = 0; c < 3; ++c) xyz[(f * k + j) * 3 + c] = dot3(b0, v0[c], b1, v1[c], b2, v2[c]);
        }
    }
    if (!cov6 && !rot_t2w) return;
    Frame fr;
    face_frame(v0, v1, v2, fr);
    const float R[3][3] = {{fr.xh[0], fr.yh[0], fr.nh[0]}, {fr.xh[1], fr.yh[1], fr.nh[1]}, {fr.xh[2], fr.yh[2], fr.nh[2]}};
    if (rot_t2w) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) rot_t2w[9 * f + 3 * a + b] = R[a][b];
    }
    if (!cov6) return;
    float L00, L01, L11;
    tri_factor(fr, rad_base, adaptive, L00, L01, L11);
    float c6[6];
    bind_cov6(fr, L00, L01, L11, thin_z, g, c6);
    for (int j = 0; j < k; ++j) {
        float2 *dst = reinterpret_cast<float2 *>(cov6 + (f * k + j) * 6);
        dst[0] = make_float2(c6[0], c6[1]);
        dst[1] = make_float2(c6[2], c6[3]);
        dst[2] = make_float2(c6[4], c6[5]);
    }
}

int launch_bind_fwd(int64_t F, int k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                    float thin_z, const float *g, int adaptive, float *xyz, float *cov6, float *rot_t2w, cudaStream_t s)
{
    if (F <= 0) return 0;
    bind_fwd_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(F, k, verts, faces, bc, rad_base, thin_z, g, adaptive,
                                                                 xyz, cov6, rot_t2w);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

__global__ void __launch_bounds__(256)
bind_bwd_kernel(int64_t F, int k, const float *__restrict__ verts, const int64_t *__restrict__ faces,
                const float *__restrict__ bc, float rad_base, float thin_z, const float *__restrict__ g_ptr, int adaptive,
                const float *__restrict__ dL_dxyz, const float *__restrict__ dL_dcov6,
                const float *__restrict__ dL_drot, float *__restrict__ dverts, float *__restrict__ dg)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float dg_local = 0.0f;
    const float g = g_ptr ? __ldg(g_ptr) : 1.0f;
    if (f < F) {
        const int64_t i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
        float v0[3], v1[3], v2[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { v0[c] = verts[3 * i0 + c]; v1[c] = verts[3 * i1 + c]; v2[c] = verts[3 * i2 + c]; }
        float dv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        if (dL_dxyz) {
            for (int j = 0; j < k; ++j) {
                const float b0 = bc[3 * j], b1 = bc[3 * j + 1], b2 = bc[3 * j + 2];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float d = dL_dxyz[(f * k + j) * 3 + c];
                    dv[0][c] = fma_(b0, d, dv[0][c]);
                    dv[1][c] = fma_(b1, d, dv[1][c]);
                    dv[2][c] = fma_(b2, d, dv[2][c]);
                }
            }
        }
        if (dL_dcov6 || dL_drot) {
            Frame fr;
            face_frame(v0, v1, v2, fr);
            float dR[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            if (dL_drot) {
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) dR[a][b] = dL_drot[9 * f + 3 * a + b];
            }
            if (dL_dcov6) {
                float L00, L01, L11;
                tri_factor(fr, rad_base, adaptive, L00, L01, L11);
                float G6[6] = {0, 0, 0, 0, 0, 0};
                for (int j = 0; j < k; ++j)
#pragma unroll
                    for (int c = 0; c < 6; ++c) G6[c] += dL_dcov6[(f * k + j) * 6 + c];
                bind_cov_adjoint(fr, L00, L01, L11, thin_z, g, G6, dR, dg_local);
            }
            frame_adjoint(fr, dR, dv);
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicAdd(dverts + 3 * i0 + a, dv[0][a]);
            atomicAdd(dverts + 3 * i1 + a, dv[1][a]);
            atomicAdd(dverts + 3 * i2 + a, dv[2][a]);
        }
    }
    if (dg) {
        // block reduction of the scale-factor gradient, one atomic per block
        __shared__ float red[8];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) dg_local += __shfl_xor_sync(0xffffffffu, dg_local, d);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dg_local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w];
            atomicAdd(dg, t);
        }
    }
}

int launch_bind_bwd(int64_t F, int k, const float *verts, const int64_t *faces, const float *bc, float rad_base,
                    float thin_z, const float *g, int adaptive, const float *dL_dxyz, const float *dL_dcov6,
                    const float *dL_drot, float *dverts, float *dg, cudaStream_t s)
{
    if (F <= 0) return 0;
    bind_bwd_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(F, k, verts, faces, bc, rad_base, thin_z, g, adaptive,
                                                                 dL_dxyz, dL_dcov6, dL_drot, dverts, dg);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
