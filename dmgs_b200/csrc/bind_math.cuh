// bind_math.cuh -- the mesh-face -> Gaussian binding arithmetic shared by binding.cu (stand-alone
// binding kernels) and preprocess.cu (binding fused into the per-Gaussian forward / backward).
// Follows scene/gaussian_geo_model_mlp_flex.py:267-311 (frame, means, cov3D_L) and :370-385
// (get_covariance_dyn); the expressions are written once here so that the fused path produces the same
// bits as the stand-alone one (and as oracle/splat_oracle.c:orc_bind_forward).
#pragma once
#include "common.cuh"

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

// Sigma = (R g L)(R g L)^T as the 6-vector (xx, xy, xz, yy, yz, zz); R = [xh | yh | nh] (columns)
__device__ __forceinline__ void bind_cov6(const Frame &fr, float L00, float L01, float L11, float thin_z, float g, float *c6)
{
    const float R[3][3] = {{fr.xh[0], fr.yh[0], fr.nh[0]}, {fr.xh[1], fr.yh[1], fr.nh[1]}, {fr.xh[2], fr.yh[2], fr.nh[2]}};
    const float l00 = L00 * g, l01 = L01 * g, l11 = L11 * g, l22 = thin_z * g;
    float N[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        N[a][0] = R[a][0] * l00;
        N[a][1] = fma_(R[a][1], l11, R[a][0] * l01);
        N[a][2] = R[a][2] * l22;
    }
    c6[0] = dot3(N[0][0], N[0][0], N[0][1], N[0][1], N[0][2], N[0][2]);
    c6[1] = dot3(N[0][0], N[1][0], N[0][1], N[1][1], N[0][2], N[1][2]);
    c6[2] = dot3(N[0][0], N[2][0], N[0][1], N[2][1], N[0][2], N[2][2]);
    c6[3] = dot3(N[1][0], N[1][0], N[1][1], N[1][1], N[1][2], N[1][2]);
    c6[4] = dot3(N[1][0], N[2][0], N[1][1], N[2][1], N[1][2], N[2][2]);
    c6[5] = dot3(N[2][0], N[2][0], N[2][1], N[2][1], N[2][2], N[2][2]);
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

// The gradient the reference's autograd produces for Sigma (cov3D_L constant, mlp_flex.py:285): G6 = dL/dcov6
// (summed over the Gaussians it stands for) -> dR += dL/dN (g L)^T, dg += <R^T dL/dN, L>.
__device__ __forceinline__ void bind_cov_adjoint(const Frame &fr, float L00, float L01, float L11, float thin_z, float g,
                                                 const float *G6, float dR[3][3], float &dg)
{
    const float R[3][3] = {{fr.xh[0], fr.yh[0], fr.nh[0]}, {fr.xh[1], fr.yh[1], fr.nh[1]}, {fr.xh[2], fr.yh[2], fr.nh[2]}};
    const float Lm[3][3] = {{L00, L01, 0}, {0, L11, 0}, {0, 0, thin_z}};
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
            dg += rt * Lm[a][b];
        }
}

// dL/dR (columns = dL/dxh, dL/dyh, dL/dnh) through the normalisations / cross products of the frame to the three
// vertices: dv[m][c] += ...  (rows: v0, v1, v2)
__device__ __forceinline__ void frame_adjoint(const Frame &fr, const float dR[3][3], float dv[3][3])
{
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

// Source of the per-Gaussian geometry when the binding is fused into preprocess (Gaussian i = face i / k,
// barycentric row i % k).
struct BindSrc {
    const float *verts;
    const int64_t *faces;
    const float *bc;
    const float *g_ptr;  // device scalar tanh(scale_factor) * max_scale, or NULL (g = 1)
    float rad_base, thin_z;
    int k, adaptive;
};

}  // namespace dmgs
