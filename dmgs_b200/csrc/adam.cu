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
    float b1, b2, one_minus_b1, one_minus_b2, sqrt_bc2, eps, grad_scale;
    int zero_grad;
};

__global__ void __launch_bounds__(256, 4)
adam_kernel(const __grid_constant__ AdamArgs a)
{
    // blockIdx.y = segment; every thread handles 4 consecutive elements (arrays are 16-byte aligned:
    // FlatGradBuffer fields are)
    const AdamSeg &sg = a.seg[blockIdx.y];
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < sg.n;
         i += (long long)gridDim.x * blockDim.x * 4) {
        const int cnt = (int)min((long long)4, sg.n - i);
        float p[4], g[4], m[4], v[4];
        const uint32_t ph = sg.period > 0 ? (uint32_t)((unsigned long long)i % (unsigned)sg.period) : 0u;
        if (cnt == 4) {
            const float4 P4 = *reinterpret_cast<const float4 *>(sg.p + i), G4 = *reinterpret_cast<const float4 *>(sg.g + i);
            const float4 M4 = *reinterpret_cast<const float4 *>(sg.m + i), V4 = *reinterpret_cast<const float4 *>(sg.v + i);
            p[0] = P4.x; p[1] = P4.y; p[2] = P4.z; p[3] = P4.w;
            g[0] = G4.x; g[1] = G4.y; g[2] = G4.z; g[3] = G4.w;
            m[0] = M4.x; m[1] = M4.y; m[2] = M4.z; m[3] = M4.w;
            v[0] = V4.x; v[1] = V4.y; v[2] = V4.z; v[3] = V4.w;
        } else {
            for (int k = 0; k < cnt; ++k) { p[k] = sg.p[i + k]; g[k] = sg.g[i + k]; m[k] = sg.m[i + k]; v[k] = sg.v[i + k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k >= cnt) break;
            const float gk = g[k] * a.grad_scale;
            m[k] = fma_(a.one_minus_b1, gk - m[k], m[k]);
            v[k] = fma_(a.one_minus_b2 * gk, gk, v[k] * a.b2);
            const float denom = sqrtf(v[k]) / a.sqrt_bc2 + a.eps;
            const float st = (sg.period > 0 && (ph + k) % (uint32_t)sg.period >= (uint32_t)sg.split) ? sg.step_hi : sg.step_lo;
            p[k] = p[k] - st * (m[k] / denom);
        }
        if (cnt == 4) {
            *reinterpret_cast<float4 *>(sg.p + i) = make_float4(p[0], p[1], p[2], p[3]);
            *reinterpret_cast<float4 *>(sg.m + i) = make_float4(m[0], m[1], m[2], m[3]);
            *reinterpret_cast<float4 *>(sg.v + i) = make_float4(v[0], v[1], v[2], v[3]);
            if (a.zero_grad) *reinterpret_cast<float4 *>(sg.g + i) = make_float4(0, 0, 0, 0);
        } else {
            for (int k = 0; k < cnt; ++k) {
                sg.p[i + k] = p[k]; sg.m[i + k] = m[k]; sg.v[i + k] = v[k];
                if (a.zero_grad) sg.g[i + k] = 0.0f;
            }
        }
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
    a.b1 = (float)beta1; a.b2 = (float)beta2;
    a.one_minus_b1 = (float)(1.0 - beta1);
    a.one_minus_b2 = (float)(1.0 - beta2);
    a.sqrt_bc2 = (float)sqrt(bc2);
    a.eps = (float)eps; a.grad_scale = grad_scale; a.zero_grad = zero_grad;
    if (pos == 0) return 0;
    long long blocks = (nmax / 4 + 255) / 256;
    const long long cap = (long long)DMGS_NUM_SMS * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<dim3((unsigned)blocks, (unsigned)nseg), 256, 0, s>>>(a);
    DMGS_CUDA(cudaGetLastError());
    count_launches(1);
    return 0;
}

}  // namespace dmgs
