// adam_math.cuh -- the one Adam update both optimiser kernels run (adam.cu: replicated step; peer.cu: the step fused
// into the gradient exchange).  Arithmetic = torch.optim.Adam, single-tensor path, fp32 (see adam.cu).
#pragma once
#include "common.cuh"

namespace dmgs {

struct AdamConsts {
    float b2, one_minus_b1, one_minus_b2, sqrt_bc2, eps, grad_scale;
};

// st: lr / bias_correction1 of this element
__device__ __forceinline__ void adam_math(const AdamConsts &c, float st, float &p, float g, float &m, float &v)
{
    const float gk = g * c.grad_scale;
    m = fma_(c.one_minus_b1, gk - m, m);
    v = fma_(c.one_minus_b2 * gk, gk, v * c.b2);
    const float denom = sqrtf(v) / c.sqrt_bc2 + c.eps;
    p = p - st * (m / denom);
}

}  // namespace dmgs
