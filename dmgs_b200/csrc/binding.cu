// binding.cu -- fused mesh-face -> Gaussian binding, forward and backward.
// Replaces the ~25 eager PyTorch kernels of scene/gaussian_geo_model_mlp_flex.py:267-311
// (face frame, barycentric means, affine cov3D_L) and :370-385 (get_covariance_dyn: R L (R L)^T,
// strip_symmetric, broadcast to the k Gaussians of the face) with one pass per direction.
// One thread per face; HBM-bound: reads 24 B of indices + 36 B of vertices, writes k*(12+24) B.
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

struct Frame {
    float xh[3], yh[3], nh[3], e1[3], e2[3], l;
};

__device__ __forceinline__ void cross3(const float *a, const float *b, float *c)
{
    c[0] = fma_(-a[2], b[1], a[1] * b[2]);
    c[1] = fma_(-a[0], b[2], a[2] * b[0]);
    c[2] = fma_(-a[1], b[0], a[0] * b[1]);
}

__device__ __forceinline__ void face_frame(const float *v0, const float *v1, const float *v2, Frame &f)
{
#pragma unroll
    for (int k = 0; k < 3; ++k) { f.e1[k] = v1[k] - v0[k]; f.e2[k] = v2[k] - v0[k]; }
    const float len = sqrtf(dot3(f.e1[0], f.e1[0], f.e1[1], f.e1[1], f.e1[2], f.e1[2]));
    f.l = fmaxf(len, 1e-12f);
#pragma unroll
    for (int k = 0; k < 3; ++k) f.xh[k] = f.e1[k] / f.l;
    float n[3], yv[3];
    cross3(f.e1, f.e2, n);
    const float nl = fmaxf(sqrtf(dot3(n[0], n[0], n[1], n[1], n[2], n[2])), 1e-12f);
#pragma unroll
    for (int k = 0; k < 3; ++k) f.nh[k] = n[k] / nl;
    cross3(f.nh, f.xh, yv);
    const float yl = fmaxf(sqrtf(dot3(yv[0], yv[0], yv[1], yv[1], yv[2], yv[2])), 1e-12f);
#pragma unroll
    for (int k = 0; k < 3; ++k) f.yh[k] = yv[k] / yl;
}

__device__ __forceinline__ void tri_factor(const Frame &f, float rad_base, int adaptive, float &L00, float &L01, float &L11)
{
    const float s = f.l * rad_base;
    L00 = s; L01 = 0.0f; L11 = s;
    if (adaptive) {
        const float Ax = dot3(f.e2[0], f.xh[0], f.e2[1], f.xh[1], f.e2[2], f.xh[2]);
        const float Ay = dot3(f.e2[0], f.yh[0], f.e2[1], f.yh[1], f.e2[2], f.yh[2]);
        const float Ex = f.l * 0.5f, Ey = f.l * 0.8660254037844386f;
        L01 = ((Ax - Ex) / Ey) * s;
        L11 = (Ay / Ey) * s;
    }
}

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
    const float l00 = L00 * g, l01 = L01 * g, l11 = L11 * g, l22 = thin_z * g;
    float N[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        N[a][0] = R[a][0] * l00;
        N[a][1] = fma_(R[a][1], l11, R[a][0] * l01);
        N[a][2] = R[a][2] * l22;
    }
    float c6[6];
    c6[0] = dot3(N[0][0], N[0][0], N[0][1], N[0][1], N[0][2], N[0][2]);
    c6[1] = dot3(N[0][0], N[1][0], N[0][1], N[1][1], N[0][2], N[1][2]);
    c6[2] = dot3(N[0][0], N[2][0], N[0][1], N[2][1], N[0][2], N[2][2]);
    c6[3] = dot3(N[1][0], N[1][0], N[1][1], N[1][1], N[1][2], N[1][2]);
    c6[4] = dot3(N[1][0], N[2][0], N[1][1], N[2][1], N[1][2], N[2][2]);
    c6[5] = dot3(N[2][0], N[2][0], N[2][1], N[2][1], N[2][2], N[2][2]);
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

// y = v / max(|v|, eps):  dv = (dy - y (y . dy)) / max(|v|, eps)
__device__ __forceinline__ void normalize_bwd(const float *v, const float *dy, float *dv)
{
    const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    if (n < 1e-12f) {
#pragma unroll
        for (int k = 0; k < 3; ++k) dv[k] = dy[k] / 1e-12f;
        return;
    }
    const float inv = 1.0f / n;
    const float y[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
    const float yd = y[0] * dy[0] + y[1] * dy[1] + y[2] * dy[2];
#pragma unroll
    for (int k = 0; k < 3; ++k) dv[k] = (dy[k] - y[k] * yd) * inv;
}
__device__ __forceinline__ void cross_plain(const float *a, const float *b, float *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
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
            const float R[3][3] = {{fr.xh[0], fr.yh[0], fr.nh[0]}, {fr.xh[1], fr.yh[1], fr.nh[1]}, {fr.xh[2], fr.yh[2], fr.nh[2]}};
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
                const float Lm[3][3] = {{L00, L01, 0}, {0, L11, 0}, {0, 0, thin_z}};
                float G6[6] = {0, 0, 0, 0, 0, 0};
                for (int j = 0; j < k; ++j)
#pragma unroll
                    for (int c = 0; c < 6; ++c) G6[c] += dL_dcov6[(f * k + j) * 6 + c];
                const float Gs[3][3] = {{2 * G6[0], G6[1], G6[2]}, {G6[1], 2 * G6[3], G6[4]}, {G6[2], G6[4], 2 * G6[5]}};
                float N[3][3], dN[3][3];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        float t = 0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) t += R[a][c] * (g * Lm[c][b]);
                        N[a][b] = t;
                    }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        float t = 0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) t += Gs[a][c] * N[c][b];
                        dN[a][b] = t;
                    }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        float t = 0, rt = 0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            t += dN[a][c] * (g * Lm[b][c]);
                            rt += R[c][a] * dN[c][b];
                        }
                        dR[a][b] += t;
                        dg_local += rt * Lm[a][b];
                    }
            }
            float dxh[3], dyh[3], dnh[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { dxh[a] = dR[a][0]; dyh[a] = dR[a][1]; dnh[a] = dR[a][2]; }
            float yv[3], dyv[3], t[3], n[3], dn[3], de1[3], de2[3], dx1[3];
            cross_plain(fr.nh, fr.xh, yv);
            normalize_bwd(yv, dyh, dyv);
            cross_plain(fr.xh, dyv, t);
#pragma unroll
            for (int a = 0; a < 3; ++a) dnh[a] += t[a];
            cross_plain(dyv, fr.nh, t);
#pragma unroll
            for (int a = 0; a < 3; ++a) dxh[a] += t[a];
            cross_plain(fr.e1, fr.e2, n);
            normalize_bwd(n, dnh, dn);
            cross_plain(fr.e2, dn, de1);
            cross_plain(dn, fr.e1, de2);
            normalize_bwd(fr.e1, dxh, dx1);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                de1[a] += dx1[a];
                dv[1][a] += de1[a];
                dv[2][a] += de2[a];
                dv[0][a] -= de1[a] + de2[a];
            }
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
