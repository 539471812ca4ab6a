// ema.cu -- the per-iteration training statistics of the reference's loop in one launch.
// main.py:325-345 walks the control network's gradients on the host: grad_norm_sqd = sum |g|^2, an exponential moving
// average of every gradient tensor (compute_EMA, utils.py:389-396), its squared norm, and -- after the optimiser step --
// the EMAs of loss / mean(w) / std(w) and the running normalisation constant (main.py:354-393): ~80 tiny torch kernels
// per iteration on 18 tensors.  Here one kernel walks a table of (grad, EMA grad) pairs, reduces the two norms with a
// warp-shuffle block reduction + one fp64 atomic per block, and the last block to finish updates the scalar EMAs.
// compute_EMA(value, ema, c, itr):  itr == 0: value;  itr <= floor(1 / c): (value + itr ema) / (itr + 1);  else
// c value + (1 - c) ema   -- evaluated in fp32 in the reference's operation order.
#include "kernels.h"

namespace socm {

constexpr int kEmaMaxTensors = 48;
struct EmaTable {
  const float* g[kEmaMaxTensors];
  float* e[kEmaMaxTensors];
  int first_block[kEmaMaxTensors + 1];  // prefix sums of ceil(n / 1024)
  int n[kEmaMaxTensors];
  int count;
};

__device__ __forceinline__ float ema_update(float value, float ema, int mode, float itr_f, float c, float omc) {
  if (mode == 0) return value;
  if (mode == 1) return __fdiv_rn(__fadd_rn(value, __fmul_rn(itr_f, ema)), __fadd_rn(itr_f, 1.0f));
  return __fadd_rn(__fmul_rn(c, value), __fmul_rn(omc, ema));
}

// scratch: double[2] partial sums (sum g^2, sum ema^2) + unsigned[2] (block ticket, pad); zero on entry, zero on exit
__global__ void __launch_bounds__(256) ema_stats_kernel(EmaTable t, const float* __restrict__ scalars,
                                                        float* __restrict__ stats, double* __restrict__ scratch, int mode,
                                                        float itr_f, float c, float omc, int mode_w, float cw, float omcw) {
  int k = 0;
  while (k + 1 < t.count && (int)blockIdx.x >= t.first_block[k + 1]) ++k;
  const int base = ((int)blockIdx.x - t.first_block[k]) * 1024;
  const float* __restrict__ g = t.g[k];
  float* __restrict__ e = t.e[k];
  double sg = 0.0, se = 0.0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = base + u * 256 + (int)threadIdx.x;
    if (i < t.n[k]) {
      const float gi = g[i];
      const float ei = ema_update(gi, e[i], mode, itr_f, c, omc);
      e[i] = ei;
      sg += (double)gi * (double)gi;
      se += (double)ei * (double)ei;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sg += __shfl_xor_sync(0xffffffffu, sg, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  __shared__ double red[2][8];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = sg;
    red[1][warp] = se;
  }
  __syncthreads();
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch + 2);
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    atomicAdd(scratch, a);
    atomicAdd(scratch + 1, b);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const float gn = (float)atomicAdd(scratch, 0.0), en = (float)atomicAdd(scratch + 1, 0.0);
    stats[0] = gn;                                                  // grad_norm_sqd            main.py:327-329
    stats[1] = ema_update(gn, stats[1], mode, itr_f, c, omc);       // EMA_grad_norm_sqd        main.py:339-341
    stats[2] = en;                                                  // sqd_norm_EMA_grad        main.py:343-345
    stats[3] = ema_update(scalars[0], stats[3], mode, itr_f, c, omc);   // EMA_loss             main.py:366-377
    stats[4] = ema_update(scalars[1], stats[4], mode, itr_f, c, omc);   // EMA_weight_mean
    stats[5] = ema_update(scalars[2], stats[5], mode, itr_f, c, omc);   // EMA_weight_std
    stats[6] = ema_update(scalars[1], stats[6], mode_w, itr_f, cw, omcw);  // normalization_const  main.py:354-359
    scratch[0] = 0.0;
    scratch[1] = 0.0;
    *ticket = 0u;
  }
}

}  // namespace socm

using namespace socm;

static int ema_mode(int itr, double coeff) {
  if (itr == 0) return 0;
  return itr <= (int)floor(1.0 / coeff) ? 1 : 2;
}

extern "C" int socm_ema_stats_f32(const socm_ema_tensor* tensors, int32_t n_tensors, const float* scalars, float* stats,
                                  void* scratch, int32_t itr, double ema_coeff, double ema_weight_mean_coeff,
                                  void* stream_) {
  SOCM_CHECK_ARG(tensors && scalars && stats && scratch && itr >= 0, "bad arguments");
  SOCM_CHECK_ARG(n_tensors >= 1 && n_tensors <= kEmaMaxTensors, "1..%d gradient tensors, got %d", kEmaMaxTensors, n_tensors);
  SOCM_CHECK_ARG(ema_coeff > 0. && ema_coeff <= 1. && ema_weight_mean_coeff > 0. && ema_weight_mean_coeff <= 1.,
                 "EMA coefficients must lie in (0, 1]");
  EmaTable tab;
  tab.count = 0;
  tab.first_block[0] = 0;
  for (int k = 0; k < n_tensors; ++k) {
    const socm_ema_tensor& a = tensors[k];
    SOCM_CHECK_ARG(a.n >= 0 && a.n < (1ll << 31), "tensor %d: bad size", k);
    if (a.n == 0) continue;
    SOCM_CHECK_ARG(a.grad && a.ema_grad, "tensor %d: NULL pointer", k);
    const int c = tab.count++;
    tab.g[c] = a.grad, tab.e[c] = a.ema_grad, tab.n[c] = (int)a.n;
    tab.first_block[c + 1] = tab.first_block[c] + (int)((a.n + 1023) / 1024);
  }
  SOCM_CHECK_ARG(tab.count > 0, "all gradient tensors are empty");
  ema_stats_kernel<<<tab.first_block[tab.count], 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      tab, scalars, stats, static_cast<double*>(scratch), ema_mode(itr, ema_coeff), (float)itr, (float)ema_coeff,
      (float)(1.0 - ema_coeff), ema_mode(itr, ema_weight_mean_coeff), (float)ema_weight_mean_coeff,
      (float)(1.0 - ema_weight_mean_coeff));
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
