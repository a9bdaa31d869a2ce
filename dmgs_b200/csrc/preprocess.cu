// preprocess.cu -- per-Gaussian forward (EWA projection, SH colour, tile rectangle, depth key)
// and the fused per-Gaussian backward (cov2D adjoint + projection + SH + scale/quaternion).
// Replaces upstream preprocessCUDA fwd/bwd + computeCov2DCUDA; the SH
// colour path also covers DMGS's python eval_sh + sigmoid (utils/sh_utils.py:41-99,
// gaussian_renderer/__init__.py:74-78).  One thread per Gaussian; HBM-bound.
#include "common.cuh"
#include "kernels.cuh"
#include "sh_stage.cuh"
#include "bind_math.cuh"

namespace dmgs {

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2_0 1.0925484305920792f
#define SH_C2_1 -1.0925484305920792f
#define SH_C2_2 0.31539156525252005f
#define SH_C2_3 -1.0925484305920792f
#define SH_C2_4 0.5462742152960396f
#define SH_C3_0 -0.5900435899266435f
#define SH_C3_1 2.890611442640554f
#define SH_C3_2 -0.4570457994644658f
#define SH_C3_3 0.3731763325901154f
#define SH_C3_4 -0.4570457994644658f
#define SH_C3_5 1.445305721320277f
#define SH_C3_6 -0.5900435899266435f

DevParams make_dev_params(const dmgs_params *p)
{
    DevParams d;
    d.P = p->P; d.sh_degree = p->sh_degree; d.M = p->sh_coeffs; d.W = p->image_width; d.H = p->image_height;
    d.sh_layout = p->sh_layout; d.sh_act = p->sh_activation;
    d.gx = (d.W + DMGS_TILE - 1) / DMGS_TILE; d.gy = (d.H + DMGS_TILE - 1) / DMGS_TILE;
    d.tanfovx = p->tanfovx; d.tanfovy = p->tanfovy;
    d.fx = (float)d.W / (2.0f * p->tanfovx); d.fy = (float)d.H / (2.0f * p->tanfovy);
    d.limx = 1.3f * p->tanfovx; d.limy = 1.3f * p->tanfovy;
    d.mod = p->scale_modifier;
    for (int i = 0; i < 3; ++i) { d.bg[i] = p->bg[i]; d.cam[i] = p->campos[i]; }
    for (int i = 0; i < 16; ++i) { d.V[i] = p->viewmatrix[i]; d.PV[i] = p->projmatrix[i]; }
    return d;
}

__device__ __forceinline__ int sh_basis(int deg, float x, float y, float z, float *b)
{
    b[0] = SH_C0;
    if (deg < 1) return 1;
    b[1] = -(SH_C1 * y);
    b[2] = SH_C1 * z;
    b[3] = -(SH_C1 * x);
    if (deg < 2) return 4;
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = SH_C2_0 * xy;
    b[5] = SH_C2_1 * yz;
    b[6] = SH_C2_2 * (fma_(2.0f, zz, -xx) - yy);
    b[7] = SH_C2_3 * xz;
    b[8] = SH_C2_4 * (xx - yy);
    if (deg < 3) return 9;
    const float t4 = fma_(4.0f, zz, -xx) - yy;
    b[9] = (SH_C3_0 * y) * fma_(3.0f, xx, -yy);
    b[10] = (SH_C3_1 * xy) * z;
    b[11] = (SH_C3_2 * y) * t4;
    b[12] = (SH_C3_3 * z) * fma_(-3.0f, yy, fma_(-3.0f, xx, 2.0f * zz));
    b[13] = (SH_C3_4 * x) * t4;
    b[14] = (SH_C3_5 * z) * (xx - yy);
    b[15] = (SH_C3_6 * x) * fma_(-3.0f, yy, xx);
    return 16;
}

__device__ __forceinline__ void quat_to_rot(const float *q, float R[3][3])
{
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    // rounding order = what nvcc (default contraction + CSE of the shared products) emits for upstream's
    // 1 - 2(yy + zz), 2(xy -+ rz), ... : see oracle/upstream_arith.cu
    R[0][0] = 1.0f - 2.0f * (y * y + z * z);
    R[0][1] = 2.0f * fma_(x, y, -(r * z));
    R[0][2] = 2.0f * fma_(r, y, x * z);
    R[1][0] = 2.0f * fma_(x, y, r * z);
    R[1][1] = 1.0f - 2.0f * fma_(x, x, z * z);
    R[1][2] = 2.0f * fma_(y, z, -(r * x));
    R[2][0] = 2.0f * fma_(-r, y, x * z);
    R[2][1] = 2.0f * fma_(y, z, r * x);
    R[2][2] = 1.0f - 2.0f * fma_(x, x, y * y);
}

// Shared between forward and backward: T = J * W (2x3), u = Sigma * T^T, cov2D (a,b,c incl. +0.3)
struct Ewa {
    float tx, ty, tz, cx, cy, txtz, tytz;
    float T0[3], T1[3], u0[3], u1[3];
    float a, b, c;
};
__device__ __forceinline__ void ewa_project(const DevParams &pr, float x, float y, float z, const float *c6, Ewa &e)
{
    e.tx = affine3(pr.V, 0, x, y, z);
    e.ty = affine3(pr.V, 1, x, y, z);
    e.tz = affine3(pr.V, 2, x, y, z);
    e.txtz = e.tx / e.tz;
    e.tytz = e.ty / e.tz;
    e.cx = fminf(pr.limx, fmaxf(-pr.limx, e.txtz)) * e.tz;
    e.cy = fminf(pr.limy, fmaxf(-pr.limy, e.tytz)) * e.tz;
    const float J00 = pr.fx / e.tz, J02 = -(pr.fx * e.cx) / (e.tz * e.tz);
    const float J11 = pr.fy / e.tz, J12 = -(pr.fy * e.cy) / (e.tz * e.tz);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        e.T0[j] = fma_(J02, pr.V[4 * j + 2], J00 * pr.V[4 * j + 0]);
        e.T1[j] = fma_(J12, pr.V[4 * j + 2], J11 * pr.V[4 * j + 1]);
    }
    const float S[3][3] = {{c6[0], c6[1], c6[2]}, {c6[1], c6[3], c6[4]}, {c6[2], c6[4], c6[5]}};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        e.u0[j] = dot3(S[j][0], e.T0[0], S[j][1], e.T0[1], S[j][2], e.T0[2]);
        e.u1[j] = dot3(S[j][0], e.T1[0], S[j][1], e.T1[1], S[j][2], e.T1[2]);
    }
    e.a = dot3(e.T0[0], e.u0[0], e.T0[1], e.u0[1], e.T0[2], e.u0[2]) + 0.3f;
    e.b = dot3(e.T0[0], e.u1[0], e.T0[1], e.u1[1], e.T0[2], e.u1[2]);  // cov[0][1] of T^T Vrk^T T
    e.c = dot3(e.T1[0], e.u1[0], e.T1[1], e.u1[1], e.T1[2], e.u1[2]) + 0.3f;
}

// SHMODE 0: coefficients read straight from global memory (any M);
// SHMODE 1 / 2: M == 16 rows staged through shared memory by TMA, layout [P,16,3] / [P,3,16].
template <int SHMODE>
__device__ __forceinline__ float sh_get(const float *__restrict__ shs, const DevParams &pr, const float *r, int i, int k,
                                        int ch)
{
    if (SHMODE == 1) return r[k * 3 + ch];
    if (SHMODE == 2) return r[ch * 16 + k];
    return pr.sh_layout == 0 ? __ldg(shs + ((size_t)i * pr.M + k) * 3 + ch) : __ldg(shs + ((size_t)i * 3 + ch) * pr.M + k);
}

// BOUND: the Gaussian's mean and covariance are not read from means3D / cov3D_precomp but built in registers
// from its mesh face (Gaussian i = face i / k, barycentric row i % k; bind_math.cuh) -- the binding fused into
// preprocess: no xyz[P,3] / cov6[P,6] round trip through HBM (xyz_out, optional, serves the texture MLP).
template <int SHMODE, bool BOUND>
__global__ void __launch_bounds__(PRE_BLK, (BOUND ? 3 : 4) * (256 / PRE_BLK))
preprocess_fwd_kernel(const __grid_constant__ DevParams pr, const __grid_constant__ BindSrc bs, float *__restrict__ xyz_out,
                      const float *__restrict__ means3D, const float *__restrict__ scales,
                      const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
                      const float *__restrict__ opacities, const float *__restrict__ shs,
                      const float *__restrict__ colors_precomp, int32_t *__restrict__ radii,
                      float *__restrict__ depths, float4 *__restrict__ rec, float4 *__restrict__ rgb4,
                      uint8_t *__restrict__ clamped, float *__restrict__ cov3D, uint32_t *__restrict__ tiles,
                      uint2 *__restrict__ rect, uint32_t *__restrict__ sort_key, uint32_t *__restrict__ sort_val,
                      uint32_t *__restrict__ total_instances, uint32_t *__restrict__ key_stat)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inb = i < pr.P;
    ShStage stage;
    if (SHMODE) stage.init(dsm);

    // invisible defaults
    int rad = 0;
    uint32_t ntiles = 0, key = 0xFFFFFFFFu;
    float depth = 0.0f;
    float4 ra = make_float4(0, 0, 0, 0), rb = make_float4(0, 0, 0, 0), col = make_float4(0, 0, 0, 0);
    uint2 rc = make_uint2(0, 0);
    uint8_t clampbits = 0;
    float c6[6] = {0, 0, 0, 0, 0, 0};
    float x = 0, y = 0, z = 0, tz = 0;
    float v0[3], v1[3], v2[3];
    if (inb) {
        if (BOUND) {
            const int64_t f = i / bs.k;
            const int j = i - (int)f * bs.k;
            const int64_t i0 = bs.faces[3 * f], i1 = bs.faces[3 * f + 1], i2 = bs.faces[3 * f + 2];
#pragma unroll
            for (int c = 0; c < 3; ++c) { v0[c] = bs.verts[3 * i0 + c]; v1[c] = bs.verts[3 * i1 + c]; v2[c] = bs.verts[3 * i2 + c]; }
            const float b0 = __ldg(bs.bc + 3 * j), b1 = __ldg(bs.bc + 3 * j + 1), b2 = __ldg(bs.bc + 3 * j + 2);
            x = dot3(b0, v0[0], b1, v1[0], b2, v2[0]);
            y = dot3(b0, v0[1], b1, v1[1], b2, v2[1]);
            z = dot3(b0, v0[2], b1, v1[2], b2, v2[2]);
            if (xyz_out) { xyz_out[3 * (size_t)i] = x; xyz_out[3 * (size_t)i + 1] = y; xyz_out[3 * (size_t)i + 2] = z; }
        } else {
            x = means3D[3 * i]; y = means3D[3 * i + 1]; z = means3D[3 * i + 2];
        }
        tz = affine3(pr.V, 2, x, y, z);
    }
    const bool near_ok = inb && tz > DMGS_NEAR;
    // fetch the SH row while the projection math runs (rows of near-culled Gaussians are skipped)
    if (SHMODE) stage.load(near_ok, shs + (size_t)i * SH_ROW_FLOATS);

    if (near_ok) {
        const float hx = affine3(pr.PV, 0, x, y, z), hy = affine3(pr.PV, 1, x, y, z), hw = affine3(pr.PV, 3, x, y, z);
        const float pw = 1.0f / (hw + 1e-7f);
        const float ppx = hx * pw, ppy = hy * pw;
        if (BOUND) {
            Frame fr;
            face_frame(v0, v1, v2, fr);
            float L00, L01, L11;
            tri_factor(fr, bs.rad_base, bs.adaptive, L00, L01, L11);
            bind_cov6(fr, L00, L01, L11, bs.thin_z, bs.g_ptr ? __ldg(bs.g_ptr) : 1.0f, c6);
        } else if (cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; ++k) c6[k] = cov3D_precomp[6 * (size_t)i + k];
        } else {
            const float s[3] = {pr.mod * scales[3 * i], pr.mod * scales[3 * i + 1], pr.mod * scales[3 * i + 2]};
            const float4 q4 = reinterpret_cast<const float4 *>(rotations)[i];
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            float R[3][3], Mm[3][3];
            quat_to_rot(q, R);
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int j = 0; j < 3; ++j) Mm[k][j] = s[k] * R[j][k];
            c6[0] = dot3(Mm[0][0], Mm[0][0], Mm[1][0], Mm[1][0], Mm[2][0], Mm[2][0]);
            c6[1] = dot3(Mm[0][0], Mm[0][1], Mm[1][0], Mm[1][1], Mm[2][0], Mm[2][1]);
            c6[2] = dot3(Mm[0][0], Mm[0][2], Mm[1][0], Mm[1][2], Mm[2][0], Mm[2][2]);
            c6[3] = dot3(Mm[0][1], Mm[0][1], Mm[1][1], Mm[1][1], Mm[2][1], Mm[2][1]);
            c6[4] = dot3(Mm[0][1], Mm[0][2], Mm[1][1], Mm[1][2], Mm[2][1], Mm[2][2]);
            c6[5] = dot3(Mm[0][2], Mm[0][2], Mm[1][2], Mm[1][2], Mm[2][2], Mm[2][2]);
        }
        Ewa e;
        ewa_project(pr, x, y, z, c6, e);
        const float det = fma_(e.a, e.c, -(e.b * e.b));
        if (det != 0.0f) {
            const float det_inv = 1.0f / det;
            const float mid = 0.5f * (e.a + e.c);
            const float sq = sqrtf(fmaxf(0.1f, fma_(mid, mid, -det)));
            const float rf = fminf(ceilf(3.0f * sqrtf(fmaxf(mid + sq, mid - sq))), 1.0e9f);
            const float px = ndc2pix(ppx, pr.W), py = ndc2pix(ppy, pr.H);
            const int x0 = clampi_f((px - rf) * 0.0625f, pr.gx), x1 = clampi_f((px + rf + 15.0f) * 0.0625f, pr.gx);
            const int y0 = clampi_f((py - rf) * 0.0625f, pr.gy), y1 = clampi_f((py + rf + 15.0f) * 0.0625f, pr.gy);
            const int nt = (x1 - x0) * (y1 - y0);
            if (nt > 0) {
                const float op = opacities[i];
                // conservative blend cut-off: power < -cut  =>  opacity*exp(power) < 1/255 for sure
                // (+inf, i.e. never culled, when a caller-supplied covariance makes the conic indefinite)
                const bool pd = det > 0.0f && e.a > 0.0f && e.c > 0.0f;
                const float cut = !pd ? __int_as_float(0x7f800000) : (op > 0.0f ? logf(255.0f * op) + 1.0e-3f : -1.0f);
                rad = (int)rf;
                ntiles = (uint32_t)nt;
                depth = e.tz;
                key = __float_as_uint(depth);
                ra = make_float4(px, py, e.c * det_inv, -e.b * det_inv);
                rb = make_float4(e.a * det_inv, op, cut, 0.0f);
                rc = make_uint2((uint32_t)x0 | ((uint32_t)x1 << 16), (uint32_t)y0 | ((uint32_t)y1 << 16));
            }
        }
    }
    if (SHMODE) stage.wait();  // warp-uniform: every issued row has landed (also required before exit)
    if (ntiles) {
        if (colors_precomp) {
            col = make_float4(colors_precomp[3 * i], colors_precomp[3 * i + 1], colors_precomp[3 * i + 2], 0.0f);
        } else {
            const float dx = x - pr.cam[0], dy = y - pr.cam[1], dz = z - pr.cam[2];
            const float len = sqrtf(dot3(dx, dx, dy, dy, dz, dz));
            float bas[16];
            const int nb = sh_basis(pr.sh_degree, dx / len, dy / len, dz / len, bas);
            float v[3] = {0.0f, 0.0f, 0.0f};
            if (SHMODE) {
                // stream the staged row (12 x LDS.128); per channel the terms arrive in increasing k,
                // the accumulation order of the arithmetic contract
                const float4 *row4 = reinterpret_cast<const float4 *>(stage.row);
#pragma unroll
                for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) {
                    const float4 q4 = row4[j];
                    const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int f = 4 * j + c;
                        const int k = SHMODE == 1 ? f / 3 : f % 16, ch = SHMODE == 1 ? f % 3 : f / 16;
                        if (k == 0) v[ch] = bas[0] * qv[c];
                        else if (k < nb) v[ch] = fma_(bas[k], qv[c], v[ch]);
                    }
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float acc = bas[0] * sh_get<0>(shs, pr, nullptr, i, 0, ch);
#pragma unroll
                    for (int k = 1; k < 16; ++k)
                        if (k < nb) acc = fma_(bas[k], sh_get<0>(shs, pr, nullptr, i, k, ch), acc);
                    v[ch] = acc;
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float acc = v[ch];
                if (pr.sh_act == 0) {
                    acc = acc + 0.5f;
                    if (acc < 0.0f) clampbits |= (uint8_t)(1u << ch);
                    acc = fmaxf(acc, 0.0f);
                } else {
                    acc = 1.0f / (1.0f + dmgs_exp(-acc));
                }
                v[ch] = acc;
            }
            col = make_float4(v[0], v[1], v[2], 0.0f);
        }
    }
    // per block: number of (Gaussian, tile) instances, and the range of the live depth keys as
    // {max of ~key, max of key} (adaptive depth sort, sort.cu; both zero-initialised).  Culled Gaussians
    // never reach a tile list, so where the sort puts them is irrelevant.
    if (total_instances) {
        __shared__ uint32_t s_red[3];
        if (threadIdx.x < 3) s_red[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t wsum = __reduce_add_sync(0xffffffffu, ntiles);
        const uint32_t k0 = __reduce_max_sync(0xffffffffu, ntiles ? ~key : 0u);
        const uint32_t k1 = __reduce_max_sync(0xffffffffu, ntiles ? key : 0u);
        if ((threadIdx.x & 31) == 0 && wsum) {
            atomicAdd(&s_red[0], wsum);
            atomicMax(&s_red[1], k0);
            atomicMax(&s_red[2], k1);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_red[0]) {
            atomicAdd(total_instances, s_red[0]);
            atomicMax(key_stat, s_red[1]);
            atomicMax(key_stat + 1, s_red[2]);
        }
    }
    if (!inb) return;
    radii[i] = rad;
    depths[i] = depth;
    rec[2 * (size_t)i] = ra;
    rec[2 * (size_t)i + 1] = rb;
    rgb4[i] = col;
    clamped[i] = clampbits;
    tiles[i] = ntiles;
    rect[i] = rc;
    sort_key[i] = key;
    sort_val[i] = (uint32_t)i;
    if (!cov3D_precomp && !BOUND) {
#pragma unroll
        for (int k = 0; k < 6; ++k) cov3D[6 * (size_t)i + k] = ntiles ? c6[k] : 0.0f;
    }
}

// staged SH rows need 16 coefficients per channel and a 16-byte aligned tensor
static int sh_mode(const dmgs_params *prm, const float *shs)
{
    if (!shs || prm->sh_coeffs != 16 || (reinterpret_cast<uintptr_t>(shs) & 15)) return 0;
    return prm->sh_layout == 0 ? 1 : 2;
}

int launch_preprocess_fwd(const dmgs_params *prm, const BindSrc *bind, float *xyz_out, const float *means3D,
                          const float *scales, const float *rotations,
                          const float *cov3D_precomp, const float *opacities, const float *shs,
                          const float *colors_precomp, int32_t *radii, void *geom, const GeomLayout &L,
                          uint32_t *total_instances, uint32_t *key_stat, cudaStream_t s)
{
    const int P = prm->P;
    if (P <= 0) return 0;
    if (once_per_device(ONCE_PREPROCESS_FWD)) {
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_fwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
    }
    const int mode = sh_mode(prm, shs);
    const DevParams dp = make_dev_params(prm);
    BindSrc bs;
    memset(&bs, 0, sizeof(bs));
    if (bind) bs = *bind;
#define DMGS_FWD_ARGS                                                                                               \
    dp, bs, xyz_out, means3D, scales, rotations, cov3D_precomp, opacities, shs, colors_precomp, radii, at<float>(geom, L.depths), \
        at<float4>(geom, L.rec), at<float4>(geom, L.rgb), at<uint8_t>(geom, L.clamped), at<float>(geom, L.cov3D),    \
        at<uint32_t>(geom, L.tiles), at<uint2>(geom, L.rect), at<uint32_t>(geom, L.keys_a), at<uint32_t>(geom, L.order), \
        total_instances, key_stat
    const int grid = (P + PRE_BLK - 1) / PRE_BLK;
    if (bind) {
        if (mode == 1) preprocess_fwd_kernel<1, true><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_FWD_ARGS);
        else if (mode == 2) preprocess_fwd_kernel<2, true><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_FWD_ARGS);
        else preprocess_fwd_kernel<0, true><<<grid, PRE_BLK, 0, s>>>(DMGS_FWD_ARGS);
    } else {
        if (mode == 1) preprocess_fwd_kernel<1, false><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_FWD_ARGS);
        else if (mode == 2) preprocess_fwd_kernel<2, false><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_FWD_ARGS);
        else preprocess_fwd_kernel<0, false><<<grid, PRE_BLK, 0, s>>>(DMGS_FWD_ARGS);
    }
#undef DMGS_FWD_ARGS
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

// View-direction term of dL/dmean: the colour depends on the mean through dir = normalize(mean - campos).
// h[k] = sum_ch sh[k][ch] * g[ch] (g = dL/dcolour after the activation); (ox,oy,oz) = mean - campos, s2 = |o|^2,
// (dxn,dyn,dzn) = o / |o|.  Adds the gradient to gm.
__device__ __forceinline__ void sh_dir_grad(int deg, const float *h, float ox, float oy, float oz, float s2, float dxn,
                                            float dyn, float dzn, float *gm)
{
    if (deg > 0) {
        const float xx = dxn * dxn, yy = dyn * dyn, zz = dzn * dzn, xy_ = dxn * dyn, yz = dyn * dzn, xz = dxn * dzn;
        float ddx = -SH_C1 * h[3], ddy = -SH_C1 * h[1], ddz = SH_C1 * h[2];
        if (deg > 1) {
            ddx += SH_C2_0 * dyn * h[4] + SH_C2_2 * 2.0f * -dxn * h[6] + SH_C2_3 * dzn * h[7] + SH_C2_4 * 2.0f * dxn * h[8];
            ddy += SH_C2_0 * dxn * h[4] + SH_C2_1 * dzn * h[5] + SH_C2_2 * 2.0f * -dyn * h[6] + SH_C2_4 * 2.0f * -dyn * h[8];
            ddz += SH_C2_1 * dyn * h[5] + SH_C2_2 * 2.0f * 2.0f * dzn * h[6] + SH_C2_3 * dxn * h[7];
            if (deg > 2) {
                ddx += SH_C3_0 * h[9] * 3.0f * 2.0f * xy_ + SH_C3_1 * h[10] * yz + SH_C3_2 * h[11] * -2.0f * xy_ +
                       SH_C3_3 * h[12] * -3.0f * 2.0f * xz + SH_C3_4 * h[13] * (-3.0f * xx + 4.0f * zz - yy) +
                       SH_C3_5 * h[14] * 2.0f * xz + SH_C3_6 * h[15] * 3.0f * (xx - yy);
                ddy += SH_C3_0 * h[9] * 3.0f * (xx - yy) + SH_C3_1 * h[10] * xz +
                       SH_C3_2 * h[11] * (-3.0f * yy + 4.0f * zz - xx) + SH_C3_3 * h[12] * -3.0f * 2.0f * yz +
                       SH_C3_4 * h[13] * -2.0f * xy_ + SH_C3_5 * h[14] * -2.0f * yz + SH_C3_6 * h[15] * -3.0f * 2.0f * xy_;
                ddz += SH_C3_1 * h[10] * xy_ + SH_C3_2 * h[11] * 4.0f * 2.0f * yz +
                       SH_C3_3 * h[12] * 3.0f * (2.0f * zz - xx - yy) + SH_C3_4 * h[13] * 4.0f * 2.0f * xz +
                       SH_C3_5 * h[14] * (xx - yy);
            }
        }
        const float inv3 = 1.0f / sqrtf(s2 * s2 * s2);
        gm[0] += ((s2 - ox * ox) * ddx - oy * ox * ddy - oz * ox * ddz) * inv3;
        gm[1] += (-ox * oy * ddx + (s2 - oy * oy) * ddy - oz * oy * ddz) * inv3;
        gm[2] += (-ox * oz * ddx - oy * oz * ddy + (s2 - oz * oz) * ddz) * inv3;
    }
}

// ------------------------------------------------------------------------------ backward
// grad_blend: per Gaussian 12 floats {dmean2D.x, dmean2D.y, dconic.a, dconic.b(half), dconic.c,
// dopacity, dcolor.r, dcolor.g, dcolor.b, pad x3} accumulated by the blend backward.
// SHMODE 3 = deferred SH gradient (accumulate == 2): no SH row is read or written, only the 16-byte record
// BOUND: mean and covariance are rebuilt from the mesh face, and dL/dmean, dL/dSigma go straight through the
// binding adjoint (the reference's truncated gradient: cov3D_L constant) to dverts / dg with atomics -- no
// dL/dxyz[P,3] / dL/dcov6[P,6] is materialised.  dL_dmeans3D = dverts [V,3], dL_dcov3D = dg [1] in that mode.
template <int SHMODE, bool BOUND>
__global__ void __launch_bounds__(PRE_BLK, (BOUND ? 2 : (SHMODE == 3 ? 4 : 3)) * (256 / PRE_BLK))
preprocess_bwd_kernel(const __grid_constant__ DevParams pr, const __grid_constant__ BindSrc bs,
                      const float *__restrict__ means3D, const float *__restrict__ scales,
                      const float *__restrict__ rotations, const float *__restrict__ cov3D_precomp,
                      const float *__restrict__ shs, const int32_t *__restrict__ radii,
                      const float *__restrict__ cov3D_state, const uint8_t *__restrict__ clamped,
                      const float4 *__restrict__ rgb4, const float4 *__restrict__ grad_blend,
                      float *__restrict__ dL_dmeans3D, float *__restrict__ dL_dmeans2D,
                      float *__restrict__ dL_dopacity, float *__restrict__ dL_dcolprec, float *__restrict__ dL_dshs,
                      float *__restrict__ dL_dscales, float *__restrict__ dL_drots, float *__restrict__ dL_dcov3D,
                      const int accumulate)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    constexpr bool STAGED = SHMODE == 1 || SHMODE == 2, DEFER = SHMODE == 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inb = i < pr.P;
    const bool vis = inb && radii[i] > 0;
    const bool do_sh = shs && dL_dshs;
    ShStage stage;
    if (STAGED) {
        stage.init(dsm);
        stage.load(vis && do_sh, shs + (size_t)i * SH_ROW_FLOATS);
    }
    float gm[3] = {0, 0, 0}, g6[6] = {0, 0, 0, 0, 0, 0};
    float gs[3] = {0, 0, 0}, gq[4] = {0, 0, 0, 0};
    float d2x = 0, d2y = 0, dop = 0, dcol[3] = {0, 0, 0};
    float x = 0, y = 0, z = 0;
    const int M = pr.M;
    // BOUND: the face of this Gaussian
    float v0[3], v1[3], v2[3], bcw[3] = {0, 0, 0}, L00 = 0, L01 = 0, L11 = 0, gscale = 1.0f;
    int64_t vi[3] = {0, 0, 0};
    Frame fr;

    if (vis) {
        const float4 ga = grad_blend[3 * (size_t)i], gb = grad_blend[3 * (size_t)i + 1], gc = grad_blend[3 * (size_t)i + 2];
        d2x = ga.x; d2y = ga.y;
        const float gcx = ga.z, gcy = ga.w, gcz = gb.x;
        dop = gb.y;
        dcol[0] = gb.z; dcol[1] = gb.w; dcol[2] = gc.x;

        float c6[6];
        if (BOUND) {
            const int64_t f = i / bs.k;
            const int j = i - (int)f * bs.k;
#pragma unroll
            for (int c = 0; c < 3; ++c) { vi[c] = bs.faces[3 * f + c]; bcw[c] = __ldg(bs.bc + 3 * j + c); }
#pragma unroll
            for (int c = 0; c < 3; ++c) { v0[c] = bs.verts[3 * vi[0] + c]; v1[c] = bs.verts[3 * vi[1] + c]; v2[c] = bs.verts[3 * vi[2] + c]; }
            x = dot3(bcw[0], v0[0], bcw[1], v1[0], bcw[2], v2[0]);
            y = dot3(bcw[0], v0[1], bcw[1], v1[1], bcw[2], v2[1]);
            z = dot3(bcw[0], v0[2], bcw[1], v1[2], bcw[2], v2[2]);
            face_frame(v0, v1, v2, fr);
            tri_factor(fr, bs.rad_base, bs.adaptive, L00, L01, L11);
            gscale = bs.g_ptr ? __ldg(bs.g_ptr) : 1.0f;
            bind_cov6(fr, L00, L01, L11, bs.thin_z, gscale, c6);
        } else {
            x = means3D[3 * i]; y = means3D[3 * i + 1]; z = means3D[3 * i + 2];
            const float *csrc = cov3D_precomp ? cov3D_precomp : cov3D_state;
#pragma unroll
            for (int k = 0; k < 6; ++k) c6[k] = csrc[6 * (size_t)i + k];
        }
        Ewa e;
        ewa_project(pr, x, y, z, c6, e);
        const float a = e.a, b = e.b, c = e.c;
        const float denom = fma_(-b, b, a * c);
        const float d2inv = 1.0f / fma_(denom, denom, 1e-7f);
        const float dL_da = d2inv * fma_(-(b * b), gcz, fma_(2.0f * b * c, gcy, -(c * c) * gcx));
        const float dL_dc = d2inv * fma_(-(b * b), gcx, fma_(2.0f * a * b, gcy, -(a * a) * gcz));
        const float dL_db = d2inv * 2.0f * fma_(-fma_(2.0f * b, b, denom), gcy, fma_(a * b, gcz, (b * c) * gcx));
        const float *T0 = e.T0, *T1 = e.T1;
        g6[0] = fma_(T1[0] * T1[0], dL_dc, fma_(T0[0] * T1[0], dL_db, (T0[0] * T0[0]) * dL_da));
        g6[3] = fma_(T1[1] * T1[1], dL_dc, fma_(T0[1] * T1[1], dL_db, (T0[1] * T0[1]) * dL_da));
        g6[5] = fma_(T1[2] * T1[2], dL_dc, fma_(T0[2] * T1[2], dL_db, (T0[2] * T0[2]) * dL_da));
        g6[1] = fma_(2.0f * T1[0] * T1[1], dL_dc, fma_(fma_(T0[1], T1[0], T0[0] * T1[1]), dL_db, (2.0f * T0[0] * T0[1]) * dL_da));
        g6[2] = fma_(2.0f * T1[0] * T1[2], dL_dc, fma_(fma_(T0[2], T1[0], T0[0] * T1[2]), dL_db, (2.0f * T0[0] * T0[2]) * dL_da));
        g6[4] = fma_(2.0f * T1[1] * T1[2], dL_dc, fma_(fma_(T0[2], T1[1], T0[1] * T1[2]), dL_db, (2.0f * T0[1] * T0[2]) * dL_da));
        float dT0[3], dT1[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            dT0[j] = fma_(e.u1[j], dL_db, 2.0f * e.u0[j] * dL_da);
            dT1[j] = fma_(e.u0[j], dL_db, 2.0f * e.u1[j] * dL_dc);
        }
        const float *V = pr.V, *PV = pr.PV;
        const float dJ00 = dot3(V[0], dT0[0], V[4], dT0[1], V[8], dT0[2]);
        const float dJ02 = dot3(V[2], dT0[0], V[6], dT0[1], V[10], dT0[2]);
        const float dJ11 = dot3(V[1], dT1[0], V[5], dT1[1], V[9], dT1[2]);
        const float dJ12 = dot3(V[2], dT1[0], V[6], dT1[1], V[10], dT1[2]);
        const float xgm = (e.txtz < -pr.limx || e.txtz > pr.limx) ? 0.0f : 1.0f;
        const float ygm = (e.tytz < -pr.limy || e.tytz > pr.limy) ? 0.0f : 1.0f;
        const float tzi = 1.0f / e.tz, tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = xgm * (-pr.fx * tz2) * dJ02;
        const float dty = ygm * (-pr.fy * tz2) * dJ12;
        const float dtz = fma_((2.0f * pr.fy * e.cy) * tz3, dJ12,
                               fma_((2.0f * pr.fx * e.cx) * tz3, dJ02, fma_(-pr.fy * tz2, dJ11, (-pr.fx * tz2) * dJ00)));
#pragma unroll
        for (int k = 0; k < 3; ++k) gm[k] = dot3(V[4 * k], dtx, V[4 * k + 1], dty, V[4 * k + 2], dtz);

        const float hx = affine3(PV, 0, x, y, z), hy = affine3(PV, 1, x, y, z), hw = affine3(PV, 3, x, y, z);
        const float mw = 1.0f / (hw + 1e-7f);
        const float mul1 = hx * mw * mw, mul2 = hy * mw * mw;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float ax = fma_(-PV[4 * k + 3], mul1, PV[4 * k + 0] * mw);
            const float ay = fma_(-PV[4 * k + 3], mul2, PV[4 * k + 1] * mw);
            gm[k] += fma_(ay, d2y, ax * d2x);
        }

        if (scales && rotations && dL_dscales && dL_drots) {
            const float s[3] = {pr.mod * scales[3 * i], pr.mod * scales[3 * i + 1], pr.mod * scales[3 * i + 2]};
            const float4 q4 = reinterpret_cast<const float4 *>(rotations)[i];
            const float q[4] = {q4.x, q4.y, q4.z, q4.w};
            float R[3][3];
            quat_to_rot(q, R);
            const float Gs[3][3] = {{g6[0], 0.5f * g6[1], 0.5f * g6[2]}, {0.5f * g6[1], g6[3], 0.5f * g6[4]}, {0.5f * g6[2], 0.5f * g6[4], g6[5]}};
            float dR[3][3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float dMk[3];
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    dMk[j] = 2.0f * dot3(s[k] * R[0][k], Gs[0][j], s[k] * R[1][k], Gs[1][j], s[k] * R[2][k], Gs[2][j]);
                gs[k] = pr.mod * dot3(R[0][k], dMk[0], R[1][k], dMk[1], R[2][k], dMk[2]);
#pragma unroll
                for (int j = 0; j < 3; ++j) dR[j][k] = s[k] * dMk[j];
            }
            const float r = q[0], qx = q[1], qy = q[2], qz = q[3];
            gq[0] = 2.0f * (qz * (dR[1][0] - dR[0][1]) + qy * (dR[0][2] - dR[2][0]) + qx * (dR[2][1] - dR[1][2]));
            gq[1] = 2.0f * (qy * (dR[0][1] + dR[1][0]) + qz * (dR[0][2] + dR[2][0]) + r * (dR[2][1] - dR[1][2])) - 4.0f * qx * (dR[1][1] + dR[2][2]);
            gq[2] = 2.0f * (qx * (dR[0][1] + dR[1][0]) + r * (dR[0][2] - dR[2][0]) + qz * (dR[1][2] + dR[2][1])) - 4.0f * qy * (dR[0][0] + dR[2][2]);
            gq[3] = 2.0f * (r * (dR[1][0] - dR[0][1]) + qx * (dR[0][2] + dR[2][0]) + qy * (dR[1][2] + dR[2][1])) - 4.0f * qz * (dR[0][0] + dR[1][1]);
        }
    }

    // ---- SH backward: dL/dsh rows and the view-direction term of dL/dmean
    if (STAGED) stage.wait();  // warp-uniform
    if (do_sh) {
        if (vis) {
            const float ox = x - pr.cam[0], oy = y - pr.cam[1], oz = z - pr.cam[2];
            const float s2 = dot3(ox, ox, oy, oy, oz, oz);
            const float len = sqrtf(s2);
            const float dxn = ox / len, dyn = oy / len, dzn = oz / len;
            float bas[16];
            const int nb = sh_basis(pr.sh_degree, dxn, dyn, dzn, bas);
            const uint8_t cb = clamped[i];
            const float4 rgbv = rgb4[i];
            const float sg[3] = {rgbv.x, rgbv.y, rgbv.z};
            float g[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                if (pr.sh_act == 0) g[ch] = (cb >> ch) & 1 ? 0.0f : dcol[ch];
                else g[ch] = dcol[ch] * (sg[ch] * (1.0f - sg[ch]));
            }
            // deferred mode: this view's contribution to dL/dsh is rank one per channel (basis(dir) x g) and its
            // view-direction term of dL/dmean is linear in g as well, so only g is recorded -- 16 bytes, and the
            // SH row is not even read -- and sh_grad_expand_kernel forms both once per step from all views' records
            if (DEFER) reinterpret_cast<float4 *>(dL_dshs)[i] = make_float4(g[0], g[1], g[2], 1.0f);
            if (!DEFER) {
            // h[k] = sum_ch sh[k][ch] * g[ch]: the view-direction gradient is linear in it, so the basis
            // Jacobian below is applied once instead of once per channel
            float h[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) h[k] = 0.0f;
            if (STAGED) {
                // stream the staged row in place: read 4 coefficients, write their 4 gradients back
                float4 *row4 = reinterpret_cast<float4 *>(stage.row);
#pragma unroll
                for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) {
                    const float4 q4 = row4[j];
                    const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
                    float ov[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int f = 4 * j + c;
                        const int k = SHMODE == 1 ? f / 3 : f % 16, ch = SHMODE == 1 ? f % 3 : f / 16;
                        const bool on = k < nb;
                        h[k] = on ? fma_(qv[c], g[ch], h[k]) : h[k];
                        ov[c] = on ? bas[k] * g[ch] : 0.0f;
                    }
                    row4[j] = make_float4(ov[0], ov[1], ov[2], ov[3]);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const size_t idx = pr.sh_layout == 0 ? ((size_t)i * M + k) * 3 + ch : ((size_t)i * 3 + ch) * M + k;
                        if (k < nb) {
                            h[k] = fma_(__ldg(shs + idx), g[ch], h[k]);
                            if (accumulate) atomicAdd(&dL_dshs[idx], bas[k] * g[ch]);
                            else dL_dshs[idx] = bas[k] * g[ch];
                        } else if (k < M && !accumulate) {
                            dL_dshs[idx] = 0.0f;
                        }
                    }
                    for (int k = 16; k < M && !accumulate; ++k) {
                        const size_t idx = pr.sh_layout == 0 ? ((size_t)i * M + k) * 3 + ch : ((size_t)i * 3 + ch) * M + k;
                        dL_dshs[idx] = 0.0f;
                    }
                }
            }
            sh_dir_grad(pr.sh_degree, h, ox, oy, oz, s2, dxn, dyn, dzn, gm);
            }
        } else if (inb && DEFER) {
            reinterpret_cast<float4 *>(dL_dshs)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);  // not seen by this view
        } else if (inb && !accumulate) {
            if (STAGED) {  // culled Gaussian: a row of zeros
                float4 *row4 = reinterpret_cast<float4 *>(stage.row);
#pragma unroll
                for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) row4[j] = make_float4(0, 0, 0, 0);
            } else {
                for (int k = 0; k < 3 * M; ++k) dL_dshs[(size_t)i * 3 * M + k] = 0.0f;
            }
        }
        if (STAGED) {
            // store: every in-range row (zeros for culled Gaussians); accumulate: only rows with a gradient
            if (inb && (vis || !accumulate))
                stage.flush_row(dL_dshs + (size_t)i * SH_ROW_FLOATS, accumulate != 0);
        }
    }

    if (BOUND) {
        // binding adjoint of this Gaussian: dL/dmean through its barycentric row, dL/dSigma through R (and g),
        // R through the frame to the three vertices; nine reductions into dverts, dg block-reduced
        float dg_local = 0.0f;
        if (vis) {
            float dv[3][3], dR[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) dv[m][c] = bcw[m] * gm[c];
            bind_cov_adjoint(fr, L00, L01, L11, bs.thin_z, gscale, g6, dR, dg_local);
            frame_adjoint(fr, dR, dv);
#pragma unroll
            for (int m = 0; m < 3; ++m)
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(dL_dmeans3D + 3 * vi[m] + c, dv[m][c]);
        }
        if (dL_dcov3D) {
            __shared__ float red[8];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) dg_local += __shfl_xor_sync(0xffffffffu, dg_local, d);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dg_local;
            __syncthreads();
            if (threadIdx.x == 0) {
                float t = 0;
#pragma unroll
                for (int w = 0; w < PRE_BLK / 32; ++w) t += red[w];
                if (t != 0.0f) atomicAdd(dL_dcov3D, t);
            }
        }
        if (inb) {
            if (accumulate) {
                if (vis) {
                    atomicAdd(&dL_dmeans2D[3 * (size_t)i], d2x);
                    atomicAdd(&dL_dmeans2D[3 * (size_t)i + 1], d2y);
                    atomicAdd(&dL_dopacity[i], dop);
                    if (dL_dcolprec) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) atomicAdd(&dL_dcolprec[3 * (size_t)i + k], dcol[k]);
                    }
                }
            } else {
                dL_dmeans2D[3 * (size_t)i] = d2x;
                dL_dmeans2D[3 * (size_t)i + 1] = d2y;
                dL_dmeans2D[3 * (size_t)i + 2] = 0.0f;
                dL_dopacity[i] = dop;
                if (dL_dcolprec) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) dL_dcolprec[3 * (size_t)i + k] = dcol[k];
                }
            }
        }
    } else if (inb) {
        if (accumulate) {
            if (vis) {
#pragma unroll
                for (int k = 0; k < 3; ++k) atomicAdd(&dL_dmeans3D[3 * (size_t)i + k], gm[k]);
                atomicAdd(&dL_dmeans2D[3 * (size_t)i], d2x);
                atomicAdd(&dL_dmeans2D[3 * (size_t)i + 1], d2y);
                atomicAdd(&dL_dopacity[i], dop);
                if (dL_dcolprec) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(&dL_dcolprec[3 * (size_t)i + k], dcol[k]);
                }
                if (dL_dcov3D) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) atomicAdd(&dL_dcov3D[6 * (size_t)i + k], g6[k]);
                }
                if (dL_dscales) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(&dL_dscales[3 * (size_t)i + k], gs[k]);
                }
                if (dL_drots) {
                    red_add_v4(dL_drots + 4 * (size_t)i, make_float4(gq[0], gq[1], gq[2], gq[3]));
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * (size_t)i + k] = gm[k];
            dL_dmeans2D[3 * (size_t)i] = d2x;
            dL_dmeans2D[3 * (size_t)i + 1] = d2y;
            dL_dmeans2D[3 * (size_t)i + 2] = 0.0f;
            dL_dopacity[i] = dop;
            if (dL_dcolprec) {
#pragma unroll
                for (int k = 0; k < 3; ++k) dL_dcolprec[3 * (size_t)i + k] = dcol[k];
            }
            if (dL_dcov3D) {
#pragma unroll
                for (int k = 0; k < 6; ++k) dL_dcov3D[6 * (size_t)i + k] = g6[k];
            }
            if (dL_dscales) {
#pragma unroll
                for (int k = 0; k < 3; ++k) dL_dscales[3 * (size_t)i + k] = gs[k];
            }
            if (dL_drots) reinterpret_cast<float4 *>(dL_drots)[i] = make_float4(gq[0], gq[1], gq[2], gq[3]);
        }
    }
    if (STAGED) stage.flush();  // shared rows must stay valid until the bulk stores have read them
}

int launch_preprocess_bwd(const dmgs_params *prm, const BindSrc *bind, const float *means3D, const float *scales,
                          const float *rotations,
                          const float *cov3D_precomp, const float *shs, const int32_t *radii, const void *geom,
                          const GeomLayout &L, const float *grad_blend, float *dL_dmeans3D, float *dL_dmeans2D,
                          float *dL_dopacity, float *dL_dcolprec, float *dL_dshs, float *dL_dscales, float *dL_drots,
                          float *dL_dcov3D, int accumulate, cudaStream_t s)
{
    const int P = prm->P;
    if (P <= 0) return 0;
    if (once_per_device(ONCE_PREPROCESS_BWD)) {
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(preprocess_bwd_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
    }
    int mode = (dL_dshs && !(reinterpret_cast<uintptr_t>(dL_dshs) & 15)) ? sh_mode(prm, shs) : 0;
    if (accumulate == 2) {
        if (!dL_dshs || !shs || (reinterpret_cast<uintptr_t>(dL_dshs) & 15)) { set_error("accumulate = 2 needs shs and a 16-byte aligned record array"); return -6; }
        if (bind) { set_error("the fused binding backward does not support the deferred SH gradient (accumulate = 2)"); return -6; }
        mode = 3;
    }
    const DevParams dp = make_dev_params(prm);
    BindSrc bs;
    memset(&bs, 0, sizeof(bs));
    if (bind) bs = *bind;
#define DMGS_BWD_ARGS                                                                                                  \
    dp, bs, means3D, scales, rotations, cov3D_precomp, shs, radii, at<float>(geom, L.cov3D), at<uint8_t>(geom, L.clamped), \
        at<float4>(geom, L.rgb), reinterpret_cast<const float4 *>(grad_blend), dL_dmeans3D, dL_dmeans2D, dL_dopacity,   \
        dL_dcolprec, dL_dshs, dL_dscales, dL_drots, dL_dcov3D, accumulate
    const int grid = (P + PRE_BLK - 1) / PRE_BLK;
    if (bind) {
        if (mode == 1) preprocess_bwd_kernel<1, true><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_BWD_ARGS);
        else if (mode == 2) preprocess_bwd_kernel<2, true><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_BWD_ARGS);
        else preprocess_bwd_kernel<0, true><<<grid, PRE_BLK, 0, s>>>(DMGS_BWD_ARGS);
    } else if (mode == 1) preprocess_bwd_kernel<1, false><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_BWD_ARGS);
    else if (mode == 2) preprocess_bwd_kernel<2, false><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(DMGS_BWD_ARGS);
    else if (mode == 3) preprocess_bwd_kernel<3, false><<<grid, PRE_BLK, 0, s>>>(DMGS_BWD_ARGS);
    else preprocess_bwd_kernel<0, false><<<grid, PRE_BLK, 0, s>>>(DMGS_BWD_ARGS);
#undef DMGS_BWD_ARGS
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------ deferred SH gradient
// dL/dsh[i] = sum over the step's views v that saw Gaussian i of basis(dir_v(i)) (x) g_v(i), and the
// view-direction term of dL/dmean[i] (linear in g_v as well), from the 16-byte records {g.r, g.g, g.b,
// seen} written by preprocess_bwd_kernel (accumulate == 2).  Per Gaussian and step this moves V * 16 B of
// records, one read of the coefficient row and one write of the gradient row instead of V reads of the
// coefficients and V read-modify-writes of the gradient row.
struct ExpandArgs {
    int P, V, sh_degree, M, layout, accumulate;
    long long view_stride;  // floats between two views' record arrays
    float cam[DMGS_MAX_STEP_VIEWS][3];
};

template <int SHMODE>
__global__ void __launch_bounds__(PRE_BLK, 2 * (256 / PRE_BLK))
sh_grad_expand_kernel(const __grid_constant__ ExpandArgs a, const float *__restrict__ means3D, const float *__restrict__ shs,
                      const float *__restrict__ records, float *__restrict__ dL_dshs, float *__restrict__ dL_dmeans3D)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool inb = i < a.P;
    ShStage stage;
    if (SHMODE) {
        stage.init(dsm);
        stage.load(inb, shs + (size_t)i * SH_ROW_FLOATS);  // the coefficients: needed for the direction term
    }
    float acc[16][3], gm[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k][0] = acc[k][1] = acc[k][2] = 0.0f;
    bool any = false;
    float x = 0, y = 0, z = 0;
    if (inb) { x = means3D[3 * (size_t)i]; y = means3D[3 * (size_t)i + 1]; z = means3D[3 * (size_t)i + 2]; }
    if (SHMODE) stage.wait();
    if (inb) {
        for (int v = 0; v < a.V; ++v) {
            const float4 g4 = reinterpret_cast<const float4 *>(records + (size_t)v * a.view_stride)[i];
            if (g4.w == 0.0f) continue;
            any = true;
            const float g[3] = {g4.x, g4.y, g4.z};
            const float ox = x - a.cam[v][0], oy = y - a.cam[v][1], oz = z - a.cam[v][2];
            const float s2 = dot3(ox, ox, oy, oy, oz, oz);
            // the gradient is held to 1e-4, not to bit parity: one MUFU.RSQ + a Newton step (< 1 ulp) replaces the IEEE
            // square root and the three IEEE divisions of the forward's direction
            float inv = rsqrtf(s2);
            inv = inv * fma_(-0.5f * s2 * inv, inv, 1.5f);
            const float dxn = ox * inv, dyn = oy * inv, dzn = oz * inv;
            float bas[16], h[16];
            const int nb = sh_basis(a.sh_degree, dxn, dyn, dzn, bas);
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                h[k] = 0.0f;
                if (k < nb) {
                    acc[k][0] = fma_(bas[k], g[0], acc[k][0]);
                    acc[k][1] = fma_(bas[k], g[1], acc[k][1]);
                    acc[k][2] = fma_(bas[k], g[2], acc[k][2]);
                }
            }
            if (a.sh_degree > 0) {
                // h[k] = sum_ch sh[k][ch] * g[ch], streamed from the staged row (or read directly)
                if (SHMODE) {
                    const float4 *row4 = reinterpret_cast<const float4 *>(stage.row);
#pragma unroll
                    for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) {
                        const float4 q4 = row4[j];
                        const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int f = 4 * j + c;
                            const int k = SHMODE == 1 ? f / 3 : f % 16, ch = SHMODE == 1 ? f % 3 : f / 16;
                            h[k] = k < nb ? fma_(qv[c], g[ch], h[k]) : h[k];
                        }
                    }
                } else {
                    for (int ch = 0; ch < 3; ++ch)
                        for (int k = 0; k < nb; ++k) {
                            const size_t idx = a.layout == 0 ? ((size_t)i * a.M + k) * 3 + ch : ((size_t)i * 3 + ch) * a.M + k;
                            h[k] = fma_(__ldg(shs + idx), g[ch], h[k]);
                        }
                }
                sh_dir_grad(a.sh_degree, h, ox, oy, oz, s2, dxn, dyn, dzn, gm);
            }
        }
        if (any && a.sh_degree > 0) {
#pragma unroll
            for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * (size_t)i + k] += gm[k];
        }
    }
    if (SHMODE) {
        // overwrite: every in-range row (zeros when no view saw the Gaussian); accumulate: rows with a gradient
        __syncwarp();
        if (inb && (any || !a.accumulate)) {
            float4 *row4 = reinterpret_cast<float4 *>(stage.row);
#pragma unroll
            for (int j = 0; j < SH_ROW_FLOATS / 4; ++j) {
                float ov[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int f = 4 * j + c;
                    const int k = SHMODE == 1 ? f / 3 : f % 16, ch = SHMODE == 1 ? f % 3 : f / 16;
                    ov[c] = acc[k][ch];
                }
                row4[j] = make_float4(ov[0], ov[1], ov[2], ov[3]);
            }
            stage.flush_row(dL_dshs + (size_t)i * SH_ROW_FLOATS, a.accumulate != 0);
        }
        stage.flush();
    } else if (inb && (any || !a.accumulate)) {
        for (int ch = 0; ch < 3; ++ch)
            for (int k = 0; k < a.M; ++k) {
                const size_t idx = a.layout == 0 ? ((size_t)i * a.M + k) * 3 + ch : ((size_t)i * 3 + ch) * a.M + k;
                const float v = k < 16 ? acc[k][ch] : 0.0f;
                dL_dshs[idx] = a.accumulate ? dL_dshs[idx] + v : v;
            }
    }
}

int launch_sh_grad_expand(int P, int sh_degree, int M, int layout, int V, const float *campos_host, const float *means3D,
                          const float *shs, const float *records, int64_t view_stride, float *dL_dshs,
                          float *dL_dmeans3D, int accumulate, cudaStream_t s)
{
    if (P <= 0) return 0;
    if (V < 0 || V > DMGS_MAX_STEP_VIEWS) { set_error("sh_grad_expand: 0..%d views per call, got %d", DMGS_MAX_STEP_VIEWS, V); return -14; }
    if (once_per_device(ONCE_SH_EXPAND)) {
        DMGS_CUDA(cudaFuncSetAttribute(sh_grad_expand_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
        DMGS_CUDA(cudaFuncSetAttribute(sh_grad_expand_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH_STAGE_SMEM));
    }
    ExpandArgs a;
    memset(&a, 0, sizeof(a));
    a.P = P; a.V = V; a.sh_degree = sh_degree; a.M = M; a.layout = layout; a.accumulate = accumulate;
    a.view_stride = view_stride;
    for (int v = 0; v < V; ++v)
        for (int c = 0; c < 3; ++c) a.cam[v][c] = campos_host[3 * v + c];
    const bool staged = M == 16 && !((reinterpret_cast<uintptr_t>(dL_dshs) | reinterpret_cast<uintptr_t>(shs)) & 15);
    const int grid = (P + PRE_BLK - 1) / PRE_BLK;
    if (staged && layout == 0) sh_grad_expand_kernel<1><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(a, means3D, shs, records, dL_dshs, dL_dmeans3D);
    else if (staged) sh_grad_expand_kernel<2><<<grid, PRE_BLK, SH_STAGE_SMEM, s>>>(a, means3D, shs, records, dL_dshs, dL_dmeans3D);
    else sh_grad_expand_kernel<0><<<grid, PRE_BLK, 0, s>>>(a, means3D, shs, records, dL_dshs, dL_dmeans3D);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

// ------------------------------------------------------------------------------ markVisible
__global__ void mark_visible_kernel(int P, const float *__restrict__ means3D, const float *__restrict__ view,
                                    uint8_t *__restrict__ visible)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float tz = affine3(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
    visible[i] = tz > DMGS_NEAR;
}

int launch_mark_visible(int P, const float *means3D, const float *view_dev, uint8_t *visible, cudaStream_t s)
{
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, view_dev, visible);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

__global__ void exp_array_kernel(const float *x, float *y, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = dmgs_exp(x[i]);
}
int launch_exp_array(const float *x, float *y, int64_t n, cudaStream_t s)
{
    if (n <= 0) return 0;
    exp_array_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, n);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
