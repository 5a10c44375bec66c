// Fused multi-tensor Adam step with the reference's positional weight decay
// (NeRF/create_nerf.py:199-258 `f_custom_adam`, nerfplusplus/custom_optim.py:11-70): SURVEY.md §8 row f2.
// One launch updates up to ADAM_MAX_TENSORS parameter tensors (the reference loops ~50 tensors in Python,
// ~10 tiny kernels each).  HBM-bound: 16 B read + 12 B written per element (p, g, m, v -> p, m, v).
#pragma once
#include "common.cuh"

namespace scnerf {

constexpr int ADAM_MAX_TENSORS = 40;
struct AdamTensor {
  float* p; const float* g; float* m; float* v; float* vmax;   // vmax: amsgrad only
  int64_t n; int32_t first_block;
  float step_size, inv_sqrt_bc2, weight_decay;   // lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t), 0 when not decayed
};
struct AdamTable { AdamTensor t[ADAM_MAX_TENSORS]; int n; float beta1, beta2, eps; };

constexpr int ADAM_BLOCK = 256, ADAM_ILP = 4;
__global__ void __launch_bounds__(ADAM_BLOCK) adam_multi_kernel(const __grid_constant__ AdamTable tab) {
  int ti = 0;
  while (ti + 1 < tab.n && (int)blockIdx.x >= tab.t[ti + 1].first_block) ++ti;
  const AdamTensor& T = tab.t[ti];
  const int64_t base = (int64_t)(blockIdx.x - T.first_block) * (ADAM_BLOCK * ADAM_ILP);
#pragma unroll
  for (int k = 0; k < ADAM_ILP; ++k) {
    const int64_t i = base + k * ADAM_BLOCK + threadIdx.x;
    if (i >= T.n) break;
    float p = T.p[i], g = T.g[i], m = T.m[i], v = T.v[i];
    if (T.weight_decay != 0.f) g = fmaf(T.weight_decay, p, g);       // grad.add(param, alpha=wd)
    m = fmaf(1.f - tab.beta1, g, m * tab.beta1);                     // exp_avg.mul_(b1).add_(grad, alpha=1-b1)
    v = fmaf(1.f - tab.beta2, g * g, v * tab.beta2);                 // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
    float vv = v;
    if (T.vmax) { vv = fmaxf(T.vmax[i], v); T.vmax[i] = vv; }
    const float denom = sqrtf(vv) * T.inv_sqrt_bc2 + tab.eps;
    T.p[i] = p - T.step_size * (m / denom);                          // param.addcdiv_(exp_avg, denom, -step_size)
    T.m[i] = m; T.v[i] = v;
  }
}

}  // namespace scnerf
