// rollout_common.cuh -- argument block and the per-path step shared by all rollout kernels
// (generic warp-per-path, FFMA tile, tcgen05 tile).
#pragma once
#include "common.cuh"

namespace socm {

// ---------------------------------------------------------------- shared argument block
struct RolloutArgs {
  socm_setting st;
  const float* warmA;
  const float* warmc;
  const float* x0;
  const float* step_tab;
  const float* noise_in;
  uint64_t seed, path_offset;
  int B, K;
  float *states, *noises, *controls, *stop, *eff_dt, *lw_det, *lw_sto, *lw_term;
};

// eps[0..d) for (path m, step k): injected or Philox
__device__ __forceinline__ void draw_noise(const RolloutArgs& a, int m, int k, float* eps) {
  const int d = a.st.d;
  if (a.noise_in != nullptr) {
    const float* src = a.noise_in + ((size_t)k * a.B + m) * d;
    for (int j = 0; j < d; ++j) eps[j] = __ldg(src + j);
  } else {
    for (int blk = 0; blk * 4 < d; ++blk) {
      float z[4];
      philox_normal4(a.seed, a.path_offset + (uint64_t)m, (uint32_t)k, (uint32_t)blk, z);
      for (int j = 0; j < 4 && blk * 4 + j < d; ++j) eps[blk * 4 + j] = z[j];
    }
  }
}

// One path, one step with given noise, executed by ONE thread: advances x, writes the outputs.
__device__ __forceinline__ void path_step_eps(const RolloutArgs& a, int m, int k, float* x, int ldx,
                                              const float* gv, int ldv, const float* eps, PathAcc& acc) {
  const int d = a.st.d, K = a.K;
  float u[kMaxDim];
  const float dt = __ldg(a.step_tab + k), sq_ldt = __ldg(a.step_tab + K + k);
  const float dt_l = __ldg(a.step_tab + 2 * K + k), sq_dtl = __ldg(a.step_tab + 3 * K + k);
  const float* wA = a.warmA ? a.warmA + (size_t)k * d * d : nullptr;
  const float* wc = a.warmc ? a.warmc + (size_t)k * d : nullptr;
  const float eff = sde_step(a.st, wA, wc, x, ldx, gv, ldv, eps, u, dt, sq_ldt, dt_l, sq_dtl, acc);
  const size_t row = (size_t)k * a.B + m;
  if (a.states) {
    float* s = a.states + (row + a.B) * d;
    for (int j = 0; j < d; ++j) s[j] = x[j * ldx];
  }
  if (a.controls) {
    float* c = a.controls + row * d;
    for (int j = 0; j < d; ++j) c[j] = u[j];
  }
  if (a.noises && a.noise_in == nullptr) {
    float* n = a.noises + row * d;
    for (int j = 0; j < d; ++j) n[j] = eps[j];
  }
  if (a.stop) a.stop[row + a.B] = acc.alive;
  if (a.eff_dt) a.eff_dt[row] = eff;
}

// One path, one step, executed by ONE thread: draws noise, advances x, writes the outputs.
__device__ __forceinline__ void path_step(const RolloutArgs& a, int m, int k, float* x, int ldx,
                                          const float* gv, int ldv, PathAcc& acc) {
  float eps[kMaxDim];
  draw_noise(a, m, k, eps);
  path_step_eps(a, m, k, x, ldx, gv, ldv, eps, acc);
}

__device__ __forceinline__ void path_finish(const RolloutArgs& a, int m, const float* x, int ldx,
                                            const PathAcc& acc) {
  a.lw_det[m] = acc.lw_det;
  a.lw_sto[m] = acc.lw_sto;
  a.lw_term[m] = __fdiv_rn(-term_cost(a.st, x, ldx), a.st.lmbd);  // utils.py:101
}

}  // namespace socm
