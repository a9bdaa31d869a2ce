/*
 * oracle/upstream_arith.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (nothing under dmgs_b200/ loads it).
 *
 * Purpose: MEASURE how far the product's private arithmetic contract (DESIGN.md section 3:
 * -fmad=false, explicit __fmaf_rn, dmgs_exp) is from the arithmetic the upstream rasteriser
 * (graphdeco-inria/diff-gaussian-rasterization, .gitmodules:4-6 -- sources absent from
 * /root/reference) would produce when nvcc compiles it with its DEFAULT flags: implicit
 * contraction (-fmad=true), CUDA's libm expf/sqrtf, and upstream's own expression grouping
 * (GLM-style column-major 3x3 products evaluated left to right,
 * `ndc2Pix` with double literals, `power = -0.5f * (A dx dx + C dy dy) - B dx dy`).
 *
 * The forward decisions (radius, tile rectangle, alpha >= 1/255, T < 1e-4) are restated here in
 * that grouping, as naive one-thread-per-item kernels, and compiled WITHOUT -fmad=false.
 * scripts/arith_divergence.py runs both on identical inputs and counts the Gaussians whose
 * radius / rectangle differ and the pixels whose contributor count differs
 * (profiles/r2_arith_divergence.json).  This is NOT the reference and is never timed or
 * reported as such: it is the algorithm of Appendix A typed from its published description.
 */
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// GLM-like column-major 3x3: m.c[col][row]; product and transpose evaluated the way GLM's
// operator* writes them (three products summed left to right per element).
struct M3 { float c[3][3]; };

__device__ __forceinline__ M3 mul(const M3 &A, const M3 &B)
{
    M3 R;
#pragma unroll
    for (int col = 0; col < 3; ++col)
#pragma unroll
        for (int row = 0; row < 3; ++row)
            R.c[col][row] = A.c[0][row] * B.c[col][0] + A.c[1][row] * B.c[col][1] + A.c[2][row] * B.c[col][2];
    return R;
}
__device__ __forceinline__ M3 transpose(const M3 &A)
{
    M3 R;
#pragma unroll
    for (int col = 0; col < 3; ++col)
#pragma unroll
        for (int row = 0; row < 3; ++row) R.c[col][row] = A.c[row][col];
    return R;
}

__device__ __forceinline__ float ndc2pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

struct UaParams {
    int P, W, H;
    float tanfovx, tanfovy, mod;
    float V[16], PV[16];
};

__global__ void ua_preprocess_kernel(const __grid_constant__ UaParams a, const float *__restrict__ means3D,
                                     const float *__restrict__ scales, const float *__restrict__ rots,
                                     const float *__restrict__ cov3D_precomp, const float *__restrict__ opacities,
                                     float *depths, int *radii, float *xy, float *conic_opacity, int *rect,
                                     unsigned *tiles_touched)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.P) return;
    radii[i] = 0;
    tiles_touched[i] = 0;
    depths[i] = 0.0f;
    xy[2 * i] = xy[2 * i + 1] = 0.0f;
    for (int k = 0; k < 4; ++k) { conic_opacity[4 * i + k] = 0.0f; rect[4 * i + k] = 0; }
    const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
    const float *V = a.V, *PV = a.PV;
    // transformPoint4x3 / 4x4
    const float vx = V[0] * px + V[4] * py + V[8] * pz + V[12];
    const float vy = V[1] * px + V[5] * py + V[9] * pz + V[13];
    const float vz = V[2] * px + V[6] * py + V[10] * pz + V[14];
    if (vz <= 0.2f) return;
    const float hx = PV[0] * px + PV[4] * py + PV[8] * pz + PV[12];
    const float hy = PV[1] * px + PV[5] * py + PV[9] * pz + PV[13];
    const float hw = PV[3] * px + PV[7] * py + PV[11] * pz + PV[15];
    const float p_w = 1.0f / (hw + 0.0000001f);
    const float projx = hx * p_w, projy = hy * p_w;

    float c6[6];
    if (cov3D_precomp) {
        for (int k = 0; k < 6; ++k) c6[k] = cov3D_precomp[6 * i + k];
    } else {
        M3 S = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
        S.c[0][0] = a.mod * scales[3 * i];
        S.c[1][1] = a.mod * scales[3 * i + 1];
        S.c[2][2] = a.mod * scales[3 * i + 2];
        const float r = rots[4 * i], x = rots[4 * i + 1], y = rots[4 * i + 2], z = rots[4 * i + 3];
        M3 R;  // constructor arguments are column by column
        R.c[0][0] = 1.f - 2.f * (y * y + z * z); R.c[0][1] = 2.f * (x * y - r * z); R.c[0][2] = 2.f * (x * z + r * y);
        R.c[1][0] = 2.f * (x * y + r * z); R.c[1][1] = 1.f - 2.f * (x * x + z * z); R.c[1][2] = 2.f * (y * z - r * x);
        R.c[2][0] = 2.f * (x * z - r * y); R.c[2][1] = 2.f * (y * z + r * x); R.c[2][2] = 1.f - 2.f * (x * x + y * y);
        const M3 M = mul(S, R);
        const M3 Sg = mul(transpose(M), M);
        c6[0] = Sg.c[0][0]; c6[1] = Sg.c[0][1]; c6[2] = Sg.c[0][2];
        c6[3] = Sg.c[1][1]; c6[4] = Sg.c[1][2]; c6[5] = Sg.c[2][2];
    }
    // computeCov2D
    const float fx = a.W / (2.0f * a.tanfovx), fy = a.H / (2.0f * a.tanfovy);
    float tx = vx, ty = vy;
    const float tz = vz;
    const float limx = 1.3f * a.tanfovx, limy = 1.3f * a.tanfovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = fminf(limx, fmaxf(-limx, txtz)) * tz;
    ty = fminf(limy, fmaxf(-limy, tytz)) * tz;
    M3 J;
    J.c[0][0] = fx / tz; J.c[0][1] = 0.0f; J.c[0][2] = -(fx * tx) / (tz * tz);
    J.c[1][0] = 0.0f; J.c[1][1] = fy / tz; J.c[1][2] = -(fy * ty) / (tz * tz);
    J.c[2][0] = 0; J.c[2][1] = 0; J.c[2][2] = 0;
    M3 Wm;
    Wm.c[0][0] = V[0]; Wm.c[0][1] = V[4]; Wm.c[0][2] = V[8];
    Wm.c[1][0] = V[1]; Wm.c[1][1] = V[5]; Wm.c[1][2] = V[9];
    Wm.c[2][0] = V[2]; Wm.c[2][1] = V[6]; Wm.c[2][2] = V[10];
    const M3 T = mul(Wm, J);
    M3 Vrk;
    Vrk.c[0][0] = c6[0]; Vrk.c[0][1] = c6[1]; Vrk.c[0][2] = c6[2];
    Vrk.c[1][0] = c6[1]; Vrk.c[1][1] = c6[3]; Vrk.c[1][2] = c6[4];
    Vrk.c[2][0] = c6[2]; Vrk.c[2][1] = c6[4]; Vrk.c[2][2] = c6[5];
    M3 cov = mul(mul(transpose(T), transpose(Vrk)), T);
    cov.c[0][0] += 0.3f;
    cov.c[1][1] += 0.3f;
    const float cx = cov.c[0][0], cy = cov.c[0][1], cz = cov.c[1][1];
    const float det = (cx * cz - cy * cy);
    if (det == 0.0f) return;
    const float det_inv = 1.f / det;
    const float conx = cz * det_inv, cony = -cy * det_inv, conz = cx * det_inv;
    const float mid = 0.5f * (cx + cz);
    const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
    const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
    const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
    const float pix_x = ndc2pix(projx, a.W), pix_y = ndc2pix(projy, a.H);
    const int gx = (a.W + 15) / 16, gy = (a.H + 15) / 16;
    const int max_radius = (int)my_radius;
    const int x0 = min(gx, max(0, (int)((pix_x - max_radius) / 16)));
    const int y0 = min(gy, max(0, (int)((pix_y - max_radius) / 16)));
    const int x1 = min(gx, max(0, (int)((pix_x + max_radius + 16 - 1) / 16)));
    const int y1 = min(gy, max(0, (int)((pix_y + max_radius + 16 - 1) / 16)));
    if ((x1 - x0) * (y1 - y0) == 0) return;
    depths[i] = vz;
    radii[i] = (int)my_radius;
    xy[2 * i] = pix_x;
    xy[2 * i + 1] = pix_y;
    conic_opacity[4 * i] = conx; conic_opacity[4 * i + 1] = cony; conic_opacity[4 * i + 2] = conz;
    conic_opacity[4 * i + 3] = opacities[i];
    rect[4 * i] = x0; rect[4 * i + 1] = x1; rect[4 * i + 2] = y0; rect[4 * i + 3] = y1;
    tiles_touched[i] = (unsigned)((y1 - y0) * (x1 - x0));
}

// one thread per pixel walks its tile's depth-ordered list front to back
__global__ void ua_blend_kernel(int W, int H, float bg0, float bg1, float bg2, const int2 *__restrict__ ranges,
                                const int *__restrict__ gidx, const float2 *__restrict__ xy,
                                const float4 *__restrict__ conic_opacity, const float *__restrict__ rgb, int rgb_stride,
                                float *out_color, float *final_T, unsigned *n_contrib)
{
    const int px = blockIdx.x * 16 + threadIdx.x, py = blockIdx.y * 16 + threadIdx.y;
    if (px >= W || py >= H) return;
    const int gx = (W + 15) / 16;
    const int2 rng = ranges[blockIdx.y * gx + blockIdx.x];
    const float pixfx = (float)px, pixfy = (float)py;
    float T = 1.0f, C[3] = {0, 0, 0};
    unsigned contributor = 0, last_contributor = 0;
    for (int j = rng.x; j < rng.y; ++j) {
        contributor++;
        const int id = gidx[j];
        const float2 p = xy[id];
        const float dx = p.x - pixfx, dy = p.y - pixfy;
        const float4 con_o = conic_opacity[id];
        const float power = -0.5f * (con_o.x * dx * dx + con_o.z * dy * dy) - con_o.y * dx * dy;
        if (power > 0.0f) continue;
        const float alpha = fminf(0.99f, con_o.w * expf(power));
        if (alpha < 1.0f / 255.0f) continue;
        const float test_T = T * (1 - alpha);
        if (test_T < 0.0001f) break;
        for (int ch = 0; ch < 3; ++ch) C[ch] += rgb[(size_t)id * rgb_stride + ch] * alpha * T;
        T = test_T;
        last_contributor = contributor;
    }
    const size_t pix = (size_t)py * W + px, HW = (size_t)H * W;
    final_T[pix] = T;
    n_contrib[pix] = last_contributor;
    out_color[pix] = C[0] + T * bg0;
    out_color[HW + pix] = C[1] + T * bg1;
    out_color[2 * HW + pix] = C[2] + T * bg2;
}

}  // namespace

extern "C" {

/* view16 / proj16: HOST floats, the flat row-major torch layout of world_view_transform / full_proj_transform.
 * Device outputs: depths f32[P], radii i32[P], xy f32[P,2], conic_opacity f32[P,4], rect i32[P,4]={x0,x1,y0,y1},
 * tiles_touched u32[P].  scales/rots or cov3D_precomp (the other NULL). */
int ua_preprocess(int P, int W, int H, float tanfovx, float tanfovy, float scale_modifier, const float *view16,
                  const float *proj16, const float *means3D, const float *scales, const float *rots,
                  const float *cov3D_precomp, const float *opacities, float *depths, int *radii, float *xy,
                  float *conic_opacity, int *rect, unsigned *tiles_touched, void *stream)
{
    if (P <= 0) return 0;
    UaParams a;
    a.P = P; a.W = W; a.H = H; a.tanfovx = tanfovx; a.tanfovy = tanfovy; a.mod = scale_modifier;
    for (int k = 0; k < 16; ++k) { a.V[k] = view16[k]; a.PV[k] = proj16[k]; }
    ua_preprocess_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a, means3D, scales, rots, cov3D_precomp,
                                                                            opacities, depths, radii, xy, conic_opacity,
                                                                            rect, tiles_touched);
    return (int)cudaGetLastError();
}

/* ranges i32[T,2], gidx i32[R]: tile-major depth-ordered instance list; rgb f32[P,rgb_stride]. */
int ua_blend(int W, int H, const float *bg3_host, const int *ranges, const int *gidx, const float *xy,
             const float *conic_opacity, const float *rgb, int rgb_stride, float *out_color, float *final_T,
             unsigned *n_contrib, void *stream)
{
    dim3 grid((W + 15) / 16, (H + 15) / 16), block(16, 16);
    ua_blend_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(W, H, bg3_host[0], bg3_host[1], bg3_host[2],
                                                               (const int2 *)ranges, gidx, (const float2 *)xy,
                                                               (const float4 *)conic_opacity, rgb, rgb_stride, out_color,
                                                               final_T, n_contrib);
    return (int)cudaGetLastError();
}

}  // extern "C"
