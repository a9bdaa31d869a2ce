// adam.cu -- fused multi-tensor Adam step straight from the (all-reduced) flat per-Gaussian gradient
// buffer.
//
// Replaces `self.optimizer.step(); self.optimizer.zero_grad(set_to_none=True)` on the reference's
// torch.optim.Adam(l, lr=0.0, eps=1e-15) with one parameter group per tensor
// (scene/gaussian_geo_model_finetune.py:526-537, train_geo_stage3.py:164-166; same construction in
// scene/gaussian_model.py training_setup): per group torch runs ~10 element-wise passes (or their
// foreach forms) over param / grad / exp_avg / exp_avg_sq.  Here ONE launch covers every group:
// 16 B read + 12 B written per element, plus 4 B to clear the gradient for the next step (the
// accumulate-mode backward adds into the buffer, so zero_grad is folded in rather than a memset).
//
// Arithmetic = torch.optim.Adam (single-tensor path, no weight decay / amsgrad / maximize), fp32:
//   m  = m + (1 - b1) (g - m)                 (lerp)
//   v  = b2 v + (1 - b2) g g                  (mul, addcmul)
//   p  = p - (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// with the bias corrections computed on the host in double precision and rounded as torch does.
#include "adam_math.cuh"
#include "common.cuh"
#include "kernels.cuh"

namespace dmgs {

constexpr int ADAM_MAX_SEG = DMGS_ADAM_MAX_SEGMENTS;

struct AdamSeg {
    float *p, *g, *m, *v;
    long long begin, n;  // position of the segment in the concatenated element range
    float step_lo, step_hi;  // lr / bias_correction1 for elements with (i % period) < split, >= split
    int period, split;
};
struct AdamArgs {
    AdamSeg seg[ADAM_MAX_SEG];
    int nseg;
    long long total;
    AdamConsts c;
    int zero_grad;
};

__device__ __forceinline__ void adam_update(const AdamArgs &a, const AdamSeg &sg, long long i, int k, float &p, float g,
                                            float &m, float &v, uint32_t ph)
{
    const float st = (sg.period > 0 && (ph + k) % (uint32_t)sg.period >= (uint32_t)sg.split) ? sg.step_hi : sg.step_lo;
    adam_math(a.c, st, p, g, m, v);
}

__global__ void __launch_bounds__(256, 4)
adam_kernel(const __grid_constant__ AdamArgs a)
{
    // blockIdx.y = segment; every thread handles two float4 groups per iteration, all eight 16-byte loads
    // issued before the first store (arrays are 16-byte aligned: FlatGradBuffer fields are)
    const AdamSeg &sg = a.seg[blockIdx.y];
    const long long n4 = sg.n >> 2, stride = (long long)gridDim.x * blockDim.x;
    float4 *P = reinterpret_cast<float4 *>(sg.p), *G = reinterpret_cast<float4 *>(sg.g);
    float4 *M = reinterpret_cast<float4 *>(sg.m), *V = reinterpret_cast<float4 *>(sg.v);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += 2 * stride) {
        const long long q1 = q + stride;
        const bool two = q1 < n4;
        float4 p0 = P[q], g0 = G[q], m0 = M[q], v0 = V[q], p1, g1, m1, v1;
        if (two) { p1 = P[q1]; g1 = G[q1]; m1 = M[q1]; v1 = V[q1]; }
        const uint32_t ph0 = sg.period > 0 ? (uint32_t)((unsigned long long)(4 * q) % (unsigned)sg.period) : 0u;
        adam_update(a, sg, 4 * q, 0, p0.x, g0.x, m0.x, v0.x, ph0);
        adam_update(a, sg, 4 * q, 1, p0.y, g0.y, m0.y, v0.y, ph0);
        adam_update(a, sg, 4 * q, 2, p0.z, g0.z, m0.z, v0.z, ph0);
        adam_update(a, sg, 4 * q, 3, p0.w, g0.w, m0.w, v0.w, ph0);
        P[q] = p0; M[q] = m0; V[q] = v0;
        if (a.zero_grad) G[q] = make_float4(0, 0, 0, 0);
        if (two) {
            const uint32_t ph1 = sg.period > 0 ? (uint32_t)((unsigned long long)(4 * q1) % (unsigned)sg.period) : 0u;
            adam_update(a, sg, 4 * q1, 0, p1.x, g1.x, m1.x, v1.x, ph1);
            adam_update(a, sg, 4 * q1, 1, p1.y, g1.y, m1.y, v1.y, ph1);
            adam_update(a, sg, 4 * q1, 2, p1.z, g1.z, m1.z, v1.z, ph1);
            adam_update(a, sg, 4 * q1, 3, p1.w, g1.w, m1.w, v1.w, ph1);
            P[q1] = p1; M[q1] = m1; V[q1] = v1;
            if (a.zero_grad) G[q1] = make_float4(0, 0, 0, 0);
        }
    }
    // the last n % 4 elements
    if (blockIdx.x == 0 && threadIdx.x < (sg.n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const uint32_t ph = sg.period > 0 ? (uint32_t)((unsigned long long)i % (unsigned)sg.period) : 0u;
        float p = sg.p[i], m = sg.m[i], v = sg.v[i];
        adam_update(a, sg, i, 0, p, sg.g[i], m, v, ph);
        sg.p[i] = p; sg.m[i] = m; sg.v[i] = v;
        if (a.zero_grad) sg.g[i] = 0.0f;
    }
}

int launch_adam(int nseg, const dmgs_adam_segment *segs, double beta1, double beta2, double eps, int64_t step,
                float grad_scale, int zero_grad, cudaStream_t s)
{
    if (nseg < 1 || nseg > ADAM_MAX_SEG) { set_error("adam: 1..%d segments per call, got %d", ADAM_MAX_SEG, nseg); return -12; }
    if (step < 1) { set_error("adam: step counts from 1"); return -12; }
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    // torch: bias_correction = 1 - beta ** step (Python doubles); step_size = lr / bc1 (double, then the
    // fp32 op addcdiv_(value=-step_size)); bias_correction2_sqrt = sqrt(bc2) (double) dividing fp32 sqrt(v)
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    long long pos = 0, nmax = 0;
    for (int k = 0; k < nseg; ++k) {
        const dmgs_adam_segment &u = segs[k];
        if (!u.param || !u.grad || !u.exp_avg || !u.exp_avg_sq || u.n < 0) { set_error("adam: bad segment %d", k); return -12; }
        const uintptr_t al = (uintptr_t)u.param | (uintptr_t)u.grad | (uintptr_t)u.exp_avg | (uintptr_t)u.exp_avg_sq;
        if (al & 15) { set_error("adam: segment %d arrays must be 16-byte aligned", k); return -12; }
        AdamSeg &d = a.seg[k];
        d.p = u.param; d.g = u.grad; d.m = u.exp_avg; d.v = u.exp_avg_sq;
        d.begin = pos; d.n = u.n;
        d.step_lo = (float)(u.lr / bc1);
        d.step_hi = (float)((u.period > 0 ? u.lr_hi : u.lr) / bc1);
        d.period = u.period; d.split = u.split;
        pos += (u.n + 3) / 4 * 4;
        if (u.n > nmax) nmax = u.n;
    }
    a.nseg = nseg; a.total = pos;
    a.c.b2 = (float)beta2;
    a.c.one_minus_b1 = (float)(1.0 - beta1);
    a.c.one_minus_b2 = (float)(1.0 - beta2);
    a.c.sqrt_bc2 = (float)sqrt(bc2);
    a.c.eps = (float)eps; a.c.grad_scale = grad_scale; a.zero_grad = zero_grad;
    if (pos == 0) return 0;
    long long blocks = (nmax / 4 + 255) / 256;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<dim3((unsigned)blocks, (unsigned)nseg), 256, 0, s>>>(a);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
