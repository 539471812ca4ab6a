// adam.cu -- one-launch Adam step over all parameter tensors of the path.
// Replaces torch.optim.Adam.step() + zero_grad() of the reference's training loop (main.py:174-230, 350-352):
// ~30 small tensors (UNet 18, M-network 6, gamma, y0) -> the stock optimiser launches a handful of kernels per
// tensor; here one kernel walks a table of tensors passed by value.  Arithmetic = torch's single-tensor Adam
// (torch/optim/adam.py, amsgrad = False, maximize = False, no weight decay), in fp32:
//   m = m + (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "kernels.h"

namespace socm {

constexpr int kAdamMaxTensors = 48;
struct AdamTable {
  float* p[kAdamMaxTensors];
  float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  float lr[kAdamMaxTensors];
  int first_block[kAdamMaxTensors + 1];  // prefix sums of ceil(n / 1024)
  int n[kAdamMaxTensors];
  int count;
};

__global__ void __launch_bounds__(256) adam_step_kernel(AdamTable t, float beta2, float omb1, float omb2, float eps,
                                                        float bc1, float bc2_sqrt, int zero_grad) {
  // which tensor does this block belong to? (count <= 48: linear scan of the prefix table in the kernel arguments)
  int k = 0;
  while (k + 1 < t.count && (int)blockIdx.x >= t.first_block[k + 1]) ++k;
  const int base = ((int)blockIdx.x - t.first_block[k]) * 1024;
  float* __restrict__ p = t.p[k];
  float* __restrict__ g = t.g[k];
  float* __restrict__ m = t.m[k];
  float* __restrict__ v = t.v[k];
  const float step_size = t.lr[k] / bc1;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = base + u * 256 + (int)threadIdx.x;
    if (i < t.n[k]) {
      const float gi = g[i];
      const float mi = m[i] + (gi - m[i]) * omb1;               // exp_avg.lerp_(grad, 1 - beta1)
      const float vi = __fmaf_rn(gi * omb2, gi, v[i] * beta2);  // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      p[i] = p[i] - step_size * (mi / denom);                             // addcdiv_(exp_avg, denom, value=-step_size)
      m[i] = mi;
      v[i] = vi;
      if (zero_grad) g[i] = 0.f;
    }
  }
}

}  // namespace socm

using namespace socm;

extern "C" int socm_adam_step_f32(const socm_adam_tensor* tensors, int32_t n_tensors, double beta1, double beta2,
                                  double eps, int32_t step, int32_t zero_grad, void* stream_) {
  SOCM_CHECK_ARG(tensors != nullptr && n_tensors >= 0 && step >= 1, "bad arguments");
  SOCM_CHECK_ARG(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., "bad hyper-parameters");
  // every scalar is formed in double and rounded once, as torch does with its Python floats: 1 - 0.999f computed
  // in fp32 would already be off by 4.7e-5
  const float bc1 = (float)(1.0 - pow(beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  int k = 0;
  while (k < n_tensors) {
    AdamTable tab;
    tab.count = 0;
    tab.first_block[0] = 0;
    for (; k < n_tensors && tab.count < kAdamMaxTensors; ++k) {
      const socm_adam_tensor& a = tensors[k];
      SOCM_CHECK_ARG(a.n >= 0 && a.n < (1ll << 31), "tensor %d: bad size", k);
      if (a.n == 0) continue;
      SOCM_CHECK_ARG(a.param && a.grad && a.exp_avg && a.exp_avg_sq, "tensor %d: NULL pointer", k);
      const int c = tab.count++;
      tab.p[c] = a.param, tab.g[c] = a.grad, tab.m[c] = a.exp_avg, tab.v[c] = a.exp_avg_sq;
      tab.lr[c] = a.lr;
      tab.n[c] = (int)a.n;
      tab.first_block[c + 1] = tab.first_block[c] + (int)((a.n + 1023) / 1024);
    }
    if (tab.count == 0) continue;
    adam_step_kernel<<<tab.first_block[tab.count], 256, 0, static_cast<cudaStream_t>(stream_)>>>(
        tab, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, bc1, bc2_sqrt, zero_grad);
    SOCM_LAUNCH_CHECK();
  }
  return SOCM_OK;
}
