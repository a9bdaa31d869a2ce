// stage3.cu -- stage-3 binding, per-Gaussian part: face frame x in-plane rotation -> world rotation,
// 2-D log-scales -> scales, and either the unit quaternion the rasteriser's (scales, rotations) mode
// consumes or the 6-vector covariance of its cov3D_precomp mode; forward and backward.
//
// Replaces scene/gaussian_geo_model_finetune.py:446-453 (get_scaling), :456-463 (get_rotation:
// get_rot_matrix + pytorch3d.transforms.matrix_to_quaternion + normalize), :465-482 (get_rot_matrix) and
// :501-516 (get_covariance) -- ~30 eager PyTorch kernels over up to 5.4 M Gaussians, the branchy
// torch.where quaternion conversion included -- with one pass per direction.  One thread per face walks
// its k Gaussians, so the gradient of the shared face frame (dL/drot_t2w, which dmgs_bind_backward then
// carries to the vertices) is summed in registers without atomics.
//
//   c = r / max(|r|, 1e-12), (a, b) = c;  R = rot_t2w [[a,-b,0],[b,a,0],[0,0,1]]   (columns: a x + b y, -b x + a y, n)
//   s = (exp(s0), exp(s1), thin_z);  Sigma = (R diag s)(R diag s)^T;  q = normalize(matrix_to_quaternion(R))
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

// pytorch3d.transforms.matrix_to_quaternion semantics: the candidate with the largest denominator.
// v is affine in the matrix entries (v[best] = q_abs[best]^2), q~ = v / (2 max(q_abs[best], 0.1)).
__device__ __forceinline__ int quat_candidate(const float m[3][3], float v[4], float &t)
{
    const float t2[4] = {1.0f + m[0][0] + m[1][1] + m[2][2], 1.0f + m[0][0] - m[1][1] - m[2][2],
                         1.0f - m[0][0] + m[1][1] - m[2][2], 1.0f - m[0][0] - m[1][1] + m[2][2]};
    int best = 0;
    float qa = t2[0] > 0.0f ? sqrtf(t2[0]) : 0.0f;
#pragma unroll
    for (int i = 1; i < 4; ++i) {
        const float q = t2[i] > 0.0f ? sqrtf(t2[i]) : 0.0f;
        if (q > qa) { qa = q; best = i; }
    }
    t = qa;
    const float sq = qa * qa;
    switch (best) {
    case 0: v[0] = sq; v[1] = m[2][1] - m[1][2]; v[2] = m[0][2] - m[2][0]; v[3] = m[1][0] - m[0][1]; break;
    case 1: v[0] = m[2][1] - m[1][2]; v[1] = sq; v[2] = m[1][0] + m[0][1]; v[3] = m[0][2] + m[2][0]; break;
    case 2: v[0] = m[0][2] - m[2][0]; v[1] = m[1][0] + m[0][1]; v[2] = sq; v[3] = m[1][2] + m[2][1]; break;
    default: v[0] = m[1][0] - m[0][1]; v[1] = m[2][0] + m[0][2]; v[2] = m[2][1] + m[1][2]; v[3] = sq; break;
    }
    return best;
}

// adjoint of v(m) for the chosen candidate: dm += A_best^T dv
__device__ __forceinline__ void quat_candidate_bwd(int best, const float dv[4], float dm[3][3])
{
    // the diagonal enters v[best] = 1 +- m00 +- m11 +- m22
    const float sg[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
#pragma unroll
    for (int i = 0; i < 3; ++i) dm[i][i] += sg[best][i] * dv[best];
    switch (best) {
    case 0:
        dm[2][1] += dv[1]; dm[1][2] -= dv[1]; dm[0][2] += dv[2]; dm[2][0] -= dv[2]; dm[1][0] += dv[3]; dm[0][1] -= dv[3]; break;
    case 1:
        dm[2][1] += dv[0]; dm[1][2] -= dv[0]; dm[1][0] += dv[2]; dm[0][1] += dv[2]; dm[0][2] += dv[3]; dm[2][0] += dv[3]; break;
    case 2:
        dm[0][2] += dv[0]; dm[2][0] -= dv[0]; dm[1][0] += dv[1]; dm[0][1] += dv[1]; dm[1][2] += dv[3]; dm[2][1] += dv[3]; break;
    default:
        dm[1][0] += dv[0]; dm[0][1] -= dv[0]; dm[2][0] += dv[1]; dm[0][2] += dv[1]; dm[2][1] += dv[2]; dm[1][2] += dv[2]; break;
    }
}

struct Stage3Local {
    float a, b, rn;      // normalised in-plane rotation and max(|r|, 1e-12)
    float R[3][3];
    float s[3];
};

__device__ __forceinline__ void stage3_local(const float *rot, const float *r2, const float *s2, float thin_z, Stage3Local &o)
{
    const float n = sqrtf(fma_(r2[1], r2[1], r2[0] * r2[0]));
    o.rn = fmaxf(n, 1e-12f);
    o.a = r2[0] / o.rn;
    o.b = r2[1] / o.rn;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        o.R[r][0] = fma_(rot[3 * r + 1], o.b, rot[3 * r] * o.a);
        o.R[r][1] = fma_(rot[3 * r + 1], o.a, -(rot[3 * r] * o.b));
        o.R[r][2] = rot[3 * r + 2];
    }
    o.s[0] = expf(s2[0]);
    o.s[1] = expf(s2[1]);
    o.s[2] = thin_z;
}

__global__ void __launch_bounds__(256)
stage3_fwd_kernel(int64_t F, int k, const float *__restrict__ rot_t2w, const float *__restrict__ rotation2d,
                  const float *__restrict__ scaling2d, float thin_z, float *__restrict__ scales, float *__restrict__ quats,
                  float *__restrict__ cov6)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float rot[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) rot[i] = rot_t2w[9 * f + i];
    for (int j = 0; j < k; ++j) {
        const int64_t g = f * k + j;
        Stage3Local L;
        stage3_local(rot, rotation2d + 2 * g, scaling2d + 2 * g, thin_z, L);
        if (scales) { scales[3 * g] = L.s[0]; scales[3 * g + 1] = L.s[1]; scales[3 * g + 2] = L.s[2]; }
        if (quats) {
            float v[4], t;
            quat_candidate(L.R, v, t);
            const float den = 2.0f * fmaxf(t, 0.1f);
            float q[4] = {v[0] / den, v[1] / den, v[2] / den, v[3] / den};
            const float qn = fmaxf(sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), 1e-12f);
            *reinterpret_cast<float4 *>(quats + 4 * g) = make_float4(q[0] / qn, q[1] / qn, q[2] / qn, q[3] / qn);
        }
        if (cov6) {
            float N[3][3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) N[r][c] = L.R[r][c] * L.s[c];
            float2 *dst = reinterpret_cast<float2 *>(cov6 + 6 * g);
            dst[0] = make_float2(dot3(N[0][0], N[0][0], N[0][1], N[0][1], N[0][2], N[0][2]),
                                 dot3(N[0][0], N[1][0], N[0][1], N[1][1], N[0][2], N[1][2]));
            dst[1] = make_float2(dot3(N[0][0], N[2][0], N[0][1], N[2][1], N[0][2], N[2][2]),
                                 dot3(N[1][0], N[1][0], N[1][1], N[1][1], N[1][2], N[1][2]));
            dst[2] = make_float2(dot3(N[1][0], N[2][0], N[1][1], N[2][1], N[1][2], N[2][2]),
                                 dot3(N[2][0], N[2][0], N[2][1], N[2][1], N[2][2], N[2][2]));
        }
    }
}

__global__ void __launch_bounds__(256)
stage3_bwd_kernel(int64_t F, int k, const float *__restrict__ rot_t2w, const float *__restrict__ rotation2d,
                  const float *__restrict__ scaling2d, float thin_z, const float *__restrict__ dL_dscales,
                  const float *__restrict__ dL_dquats, const float *__restrict__ dL_dcov6, float *__restrict__ dL_drot,
                  float *__restrict__ dL_drotation2d, float *__restrict__ dL_dscaling2d)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    float rot[9], drot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 9; ++i) rot[i] = rot_t2w[9 * f + i];
    for (int j = 0; j < k; ++j) {
        const int64_t g = f * k + j;
        Stage3Local L;
        stage3_local(rot, rotation2d + 2 * g, scaling2d + 2 * g, thin_z, L);
        float dR[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, ds[3] = {0, 0, 0};
        if (dL_dscales) { ds[0] = dL_dscales[3 * g]; ds[1] = dL_dscales[3 * g + 1]; }
        if (dL_dquats) {
            float v[4], t;
            const int best = quat_candidate(L.R, v, t);
            const float tc = fmaxf(t, 0.1f), den = 2.0f * tc;
            const float qt[4] = {v[0] / den, v[1] / den, v[2] / den, v[3] / den};
            const float qn = fmaxf(sqrtf(qt[0] * qt[0] + qt[1] * qt[1] + qt[2] * qt[2] + qt[3] * qt[3]), 1e-12f);
            const float4 gq4 = *reinterpret_cast<const float4 *>(dL_dquats + 4 * g);
            const float gq[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
            float dot = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) dot += (qt[i] / qn) * gq[i];
            float gt[4], gv = 0.0f;  // gradient w.r.t. q~, and g~ . v
#pragma unroll
            for (int i = 0; i < 4; ++i) { gt[i] = (gq[i] - (qt[i] / qn) * dot) / qn; gv += gt[i] * v[i]; }
            float dv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) dv[i] = gt[i] / den;
            // t = sqrt(v[best]) enters the denominator (unless clamped at 0.1)
            if (t > 0.1f) dv[best] -= gv / (4.0f * t * t * t);
            quat_candidate_bwd(best, dv, dR);
        }
        if (dL_dcov6) {
            const float *G6 = dL_dcov6 + 6 * g;
            const float Gs[3][3] = {{2 * G6[0], G6[1], G6[2]}, {G6[1], 2 * G6[3], G6[4]}, {G6[2], G6[4], 2 * G6[5]}};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float gc[3], dotc = 0.0f;  // (G + G^T) R[:,c]
#pragma unroll
                for (int r = 0; r < 3; ++r) gc[r] = Gs[r][0] * L.R[0][c] + Gs[r][1] * L.R[1][c] + Gs[r][2] * L.R[2][c];
#pragma unroll
                for (int r = 0; r < 3; ++r) { dR[r][c] += L.s[c] * L.s[c] * gc[r]; dotc += L.R[r][c] * gc[r]; }
                ds[c] += L.s[c] * dotc;
            }
        }
        // scales -> log-scales (thin_z is a constant)
        if (dL_dscaling2d) { dL_dscaling2d[2 * g] = ds[0] * L.s[0]; dL_dscaling2d[2 * g + 1] = ds[1] * L.s[1]; }
        // R columns: a x + b y, -b x + a y, n  ->  (a, b), the frame
        float da = 0.0f, db = 0.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float x = rot[3 * r], y = rot[3 * r + 1];
            da += dR[r][0] * x + dR[r][1] * y;
            db += dR[r][0] * y - dR[r][1] * x;
            drot[3 * r] += L.a * dR[r][0] - L.b * dR[r][1];
            drot[3 * r + 1] += L.b * dR[r][0] + L.a * dR[r][1];
            drot[3 * r + 2] += dR[r][2];
        }
        if (dL_drotation2d) {
            // c = r / max(|r|, eps): dr = (dc - c (c . dc)) / |r|  (the clamped branch is a plain division)
            const float r0 = rotation2d[2 * g], r1 = rotation2d[2 * g + 1];
            const bool clamped = sqrtf(fma_(r1, r1, r0 * r0)) < 1e-12f;
            const float cd = clamped ? 0.0f : L.a * da + L.b * db;
            dL_drotation2d[2 * g] = (da - L.a * cd) / L.rn;
            dL_drotation2d[2 * g + 1] = (db - L.b * cd) / L.rn;
        }
    }
    if (dL_drot) {
#pragma unroll
        for (int i = 0; i < 9; ++i) dL_drot[9 * f + i] = drot[i];
    }
}

int launch_stage3_fwd(int64_t F, int k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                      float thin_z, float *scales, float *quats, float *cov6, cudaStream_t s)
{
    if (F <= 0) return 0;
    stage3_fwd_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(F, k, rot_t2w, rotation2d, scaling2d, thin_z, scales,
                                                                   quats, cov6);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

int launch_stage3_bwd(int64_t F, int k, const float *rot_t2w, const float *rotation2d, const float *scaling2d,
                      float thin_z, const float *dL_dscales, const float *dL_dquats, const float *dL_dcov6, float *dL_drot,
                      float *dL_drotation2d, float *dL_dscaling2d, cudaStream_t s)
{
    if (F <= 0) return 0;
    stage3_bwd_kernel<<<(unsigned)((F + 255) / 256), 256, 0, s>>>(F, k, rot_t2w, rotation2d, scaling2d, thin_z, dL_dscales,
                                                                   dL_dquats, dL_dcov6, dL_drot, dL_drotation2d,
                                                                   dL_dscaling2d);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
