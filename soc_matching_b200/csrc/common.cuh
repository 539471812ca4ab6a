// common.cuh -- shared device code: setting primitives, the per-path SDE step, Philox.
// Semantics follow SURVEY.md Appendix A.1 (a restatement of the reference's
// utils.py:17-128); op order in the state update is kept (no FMA contraction) because the
// stopping index of molecular_dynamics depends on the sign of |Phi| ~ 1e-7 quantities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/socm_b200.h"

namespace socm {

constexpr int kMaxDim = SOCM_MAX_DIM;

// ---------------------------------------------------------------- host-side error plumbing
void set_error(const char* fmt, ...);
#define SOCM_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::socm::set_error(__VA_ARGS__);      \
      return SOCM_ERR_INVALID;             \
    }                                      \
  } while (0)
#define SOCM_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e_ = (call);                                                         \
    if (e_ != cudaSuccess) {                                                         \
      ::socm::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return SOCM_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)
#define SOCM_LAUNCH_CHECK()                                                          \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      ::socm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return SOCM_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

int validate_setting(const socm_setting* st);
int validate_unet(const socm_unet* net, int d);
bool is_default_arch(const socm_unet* net);  // hdims == [256,128,64]
int f16_default();                           // SOCM_F16 environment switch: 1 / 0 force the fp16-split / 3xTF32 engine, -1 unset
int sm_count();

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

__device__ __forceinline__ float u32_to_unit(uint32_t r) {
  // (r + 0.5) * 2^-32 in (0, 1]
  return fmaf(__uint2float_rn(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

// Four N(0,1) draws for (path, step, block of 4 components).
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t path, uint32_t step,
                                               uint32_t blk, float z[4]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)path, (uint32_t)(path >> 32), step, blk),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float r0 = sqrtf(-2.0f * logf(u32_to_unit(r.x)));
  const float r1 = sqrtf(-2.0f * logf(u32_to_unit(r.z)));
  float s0, c0, s1, c1;
  sincospif(2.0f * u32_to_unit(r.y), &s0, &c0);
  sincospif(2.0f * u32_to_unit(r.w), &s1, &c1);
  z[0] = r0 * c0;
  z[1] = r0 * s0;
  z[2] = r1 * c1;
  z[3] = r1 * s1;
}

// ---------------------------------------------------------------- setting primitives
// x is accessed as x[j * ld] so that the same code serves a thread-private array (ld = 1)
// and a column of a feature-major shared-memory tile (ld = tile pitch).
__device__ __forceinline__ float dw_drift(float kap, float x) {
  // -2 * kappa * (x**2 - 1) * 2 * x   evaluated left to right (double_well.py:44-48)
  const float q = __fadd_rn(__fmul_rn(x, x), -1.0f);
  return __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(-2.0f, kap), q), 2.0f), x);
}

__device__ __forceinline__ float dw_drift_diag_grad(float kap, float x) {
  // -(8 kappa x^2 + 4 kappa (x^2 - 1))   (double_well.py:51-61)
  const float x2 = __fmul_rn(x, x);
  return -__fadd_rn(__fmul_rn(__fmul_rn(8.0f, kap), x2), __fmul_rn(__fmul_rn(4.0f, kap), __fadd_rn(x2, -1.0f)));
}

__device__ __forceinline__ void drift_vec(const socm_setting& st, const float* x, int ld, float* out) {
  const int d = st.d;
  if (st.kind == SOCM_OU_QUADRATIC || st.kind == SOCM_OU_LINEAR) {
    for (int i = 0; i < d; ++i) {
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(__ldg(st.A + i * d + j), x[j * ld], acc);
      out[i] = acc;
    }
  } else {
    for (int i = 0; i < d; ++i) out[i] = dw_drift(__ldg(st.kappa + i), x[i * ld]);
  }
}

__device__ __forceinline__ float run_cost(const socm_setting& st, const float* x, int ld) {
  if (st.kind == SOCM_OU_QUADRATIC) {
    const int d = st.d;
    float tot = 0.f;
    for (int i = 0; i < d; ++i) {
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(__ldg(st.P + i * d + j), x[j * ld], acc);
      tot = fmaf(x[i * ld], acc, tot);
    }
    return tot;
  }
  return st.kind == SOCM_MOLECULAR_DYNAMICS ? 1.0f : 0.0f;
}

__device__ __forceinline__ float term_cost(const socm_setting& st, const float* x, int ld) {
  const int d = st.d;
  float tot = 0.f;
  if (st.kind == SOCM_OU_QUADRATIC) {
    for (int i = 0; i < d; ++i) {
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(__ldg(st.Q + i * d + j), x[j * ld], acc);
      tot = fmaf(x[i * ld], acc, tot);
    }
  } else if (st.kind == SOCM_OU_LINEAR) {
    for (int i = 0; i < d; ++i) tot = fmaf(__ldg(st.omega + i), x[i * ld], tot);
  } else if (st.kind == SOCM_DOUBLE_WELL) {
    for (int i = 0; i < d; ++i) {
      const float q = __fadd_rn(__fmul_rn(x[i * ld], x[i * ld]), -1.0f);
      tot = __fadd_rn(tot, __fmul_rn(__ldg(st.nu + i), __fmul_rn(q, q)));
    }
  }
  return tot;
}

// grad_f(x), grad_g(x) into out[0..d)
__device__ __forceinline__ void grad_run_cost(const socm_setting& st, const float* x, int ld, float* out) {
  const int d = st.d;
  if (st.kind == SOCM_OU_QUADRATIC) {
    for (int i = 0; i < d; ++i) {
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(__ldg(st.P + i * d + j), x[j * ld], acc);
      out[i] = 2.0f * acc;
    }
  } else {
    for (int i = 0; i < d; ++i) out[i] = 0.f;
  }
}

__device__ __forceinline__ void grad_term_cost(const socm_setting& st, const float* x, int ld, float* out) {
  const int d = st.d;
  if (st.kind == SOCM_OU_QUADRATIC) {
    for (int i = 0; i < d; ++i) {
      float acc = 0.f;
      for (int j = 0; j < d; ++j) acc = fmaf(__ldg(st.Q + i * d + j), x[j * ld], acc);
      out[i] = 2.0f * acc;
    }
  } else if (st.kind == SOCM_OU_LINEAR) {
    for (int i = 0; i < d; ++i) out[i] = __ldg(st.omega + i);
  } else if (st.kind == SOCM_DOUBLE_WELL) {
    for (int i = 0; i < d; ++i) {
      const float xi = x[i * ld];
      const float q = __fadd_rn(__fmul_rn(xi, xi), -1.0f);
      out[i] = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.0f, __ldg(st.nu + i)), q), 2.0f), xi);
    }
  } else {
    for (int i = 0; i < d; ++i) out[i] = 0.f;
  }
}

// out = nabla_b(x) . c   with nabla_b contracted on its LAST index (method.py:614-624):
// OU settings return A^T (OU_quadratic.py:55-63) -> out_l = sum_n A[n][l] c_n; DW/MD: diagonal.
__device__ __forceinline__ void grad_drift_dot(const socm_setting& st, const float* x, int ld,
                                               const float* c, float* out) {
  const int d = st.d;
  if (st.kind == SOCM_OU_QUADRATIC || st.kind == SOCM_OU_LINEAR) {
    for (int l = 0; l < d; ++l) {
      float acc = 0.f;
      for (int n = 0; n < d; ++n) acc = fmaf(__ldg(st.A + n * d + l), c[n], acc);
      out[l] = acc;
    }
  } else {
    for (int l = 0; l < d; ++l) out[l] = dw_drift_diag_grad(__ldg(st.kappa + l), x[l * ld]) * c[l];
  }
}

// out_i = sum_j M[i][j] v[j]      (M row-major d x d)
__device__ __forceinline__ void matvec(const float* M, int d, const float* v, float* out) {
  for (int i = 0; i < d; ++i) {
    float acc = 0.f;
    for (int j = 0; j < d; ++j) acc = fmaf(__ldg(M + i * d + j), v[j], acc);
    out[i] = acc;
  }
}
// out_i = sum_j M[j][i] v[j]      (transpose product)
__device__ __forceinline__ void matvec_t(const float* M, int d, const float* v, float* out) {
  for (int i = 0; i < d; ++i) {
    float acc = 0.f;
    for (int j = 0; j < d; ++j) acc = fmaf(__ldg(M + j * d + i), v[j], acc);
    out[i] = acc;
  }
}

// Warm-start control  sigma^{-1}(c_k + A_k x - b(x))  added to u   (models.py:163-179).
__device__ __forceinline__ void add_warm_start(const socm_setting& st, const float* Ak, const float* ck,
                                               const float* x, int ld, float* u) {
  const int d = st.d;
  float aff[kMaxDim], bx[kMaxDim];
  drift_vec(st, x, ld, bx);
  for (int i = 0; i < d; ++i) {
    float acc = __ldg(ck + i);
    for (int j = 0; j < d; ++j) acc = fmaf(__ldg(Ak + i * d + j), x[j * ld], acc);
    aff[i] = acc - bx[i];
  }
  if (st.sigma_is_identity) {
    for (int i = 0; i < d; ++i) u[i] += aff[i];
  } else {
    float t[kMaxDim];
    matvec(st.sigma_inv, d, aff, t);
    for (int i = 0; i < d; ++i) u[i] += t[i];
  }
}

// ---------------------------------------------------------------- one Euler-Maruyama step of one path
struct PathAcc {
  float alive;   // 1 while Phi > 0 (utils.py:34, 74)
  float lw_det;  // log_path_weight_deterministic
  float lw_sto;  // log_path_weight_stochastic
};

// x[j*ldx]: state (updated in place);  gv[j*ldv]: UNet output nabla_V(t_k, x);
// eps[d], u[d]: thread-private, filled by the caller (eps) / here (u).
// Returns eff_dt (utils.py:70-78).  step constants: dt, sq_ldt = sqrt(lmbd*dt),
// dt_l = dt/lmbd, sq_dtl = sqrt(dt/lmbd).
__device__ __forceinline__ float sde_step(const socm_setting& st, const float* warmA, const float* warmc,
                                          float* x, int ldx, const float* gv, int ldv,
                                          const float* eps, float* u, float dt, float sq_ldt, float dt_l,
                                          float sq_dtl, PathAcc& acc) {
  const int d = st.d;
  float tmp[kMaxDim], se[kMaxDim];
  // u = -sigma^T nabla_V (method.py:68-72); gv == nullptr: u already holds the control (tabulated u, method.py:103-107)
  if (gv == nullptr) {
  } else if (st.sigma_is_identity) {
    for (int i = 0; i < d; ++i) u[i] = -gv[i * ldv];
  } else {
    for (int i = 0; i < d; ++i) {
      float a = 0.f;
      for (int j = 0; j < d; ++j) a = fmaf(__ldg(st.sigma + j * d + i), gv[j * ldv], a);
      u[i] = -a;
    }
  }
  if (warmA != nullptr) add_warm_start(st, warmA, warmc, x, ldx, u);

  // update = (b(x) + sigma u) dt + sqrt(lmbd dt) sigma eps   (utils.py:45-47)
  drift_vec(st, x, ldx, tmp);
  if (st.sigma_is_identity) {
    for (int i = 0; i < d; ++i) {
      tmp[i] = __fadd_rn(tmp[i], u[i]);
      se[i] = eps[i];
    }
  } else {
    float su[kMaxDim];
    matvec(st.sigma, d, u, su);
    matvec(st.sigma, d, eps, se);
    for (int i = 0; i < d; ++i) tmp[i] = __fadd_rn(tmp[i], su[i]);
  }
  for (int i = 0; i < d; ++i) tmp[i] = __fadd_rn(__fmul_rn(tmp[i], dt), __fmul_rn(sq_ldt, se[i]));

  float eff_dt = dt;
  float a_l = dt_l, sq_l = sq_dtl;
  if (st.kind == SOCM_MOLECULAR_DYNAMICS) {
    // stopping logic, utils.py:49-75 (strict inequalities; frac squared; two 1e-6 fudges)
    const float phi0 = -x[0];
    float xn0 = __fadd_rn(x[0], __fmul_rn(acc.alive, tmp[0]));
    const float phi1 = -xn0;
    const float still = (phi0 > 0.f && phi1 > 0.f) ? 1.f : 0.f;
    const float crossed = (phi0 > 0.f && phi1 < 0.f) ? 1.f : 0.f;
    const float frac =
        __fmul_rn(crossed, __fadd_rn(__fdiv_rn(phi0, __fadd_rn(__fadd_rn(phi0, -phi1), 1e-6f)), 1e-6f));
    const float fa = __fmul_rn(frac, acc.alive);
    for (int i = 0; i < d; ++i) {
      const float xb = x[i * ldx];
      const float xn = __fadd_rn(xb, __fmul_rn(acc.alive, tmp[i]));
      const float xf = __fadd_rn(xb, __fmul_rn(fa, tmp[i]));
      x[i * ldx] = __fadd_rn(__fmul_rn(crossed, xf), __fmul_rn(__fadd_rn(1.f, -crossed), xn));
    }
    eff_dt = __fadd_rn(__fmul_rn(__fmul_rn(crossed, __fmul_rn(frac, frac)), dt), __fmul_rn(still, dt));
    acc.alive = (-x[0] > 0.f) ? 1.f : 0.f;
    a_l = __fdiv_rn(eff_dt, st.lmbd);
    sq_l = __fsqrt_rn(a_l);
  } else {
    for (int i = 0; i < d; ++i) x[i * ldx] = __fadd_rn(x[i * ldx], __fmul_rn(acc.alive, tmp[i]));
  }

  // importance-weight accumulators; running cost at the POST-update state (utils.py:82-99)
  float uu = 0.f, ue = 0.f;
  for (int i = 0; i < d; ++i) {
    uu = __fadd_rn(uu, __fmul_rn(u[i], u[i]));
    ue = __fadd_rn(ue, __fmul_rn(u[i], eps[i]));
  }
  const float f = run_cost(st, x, ldx);
  acc.lw_det = __fadd_rn(acc.lw_det, __fmul_rn(a_l, __fadd_rn(-f, -__fmul_rn(0.5f, uu))));
  acc.lw_sto = __fadd_rn(acc.lw_sto, __fmul_rn(sq_l, -ue));
  return eff_dt;
}

// ---------------------------------------------------------------- register-resident variant of sde_step
// for the "diagonal" settings: sigma = identity, double-well type drift (kinds DOUBLE_WELL and
// MOLECULAR_DYNAMICS), no warm start.  Same operations in the same order as sde_step, but on DM
// compile-time lanes (DM >= d; lanes >= d must hold zeros in x, gv, eps, kap and stay zero), so
// every vector lives in registers.  kap[] = kappa padded with zeros.
__host__ __device__ inline bool diag_fast_path(const socm_setting& st, bool warm) {
  return st.sigma_is_identity && !warm && (st.kind == SOCM_DOUBLE_WELL || st.kind == SOCM_MOLECULAR_DYNAMICS);
}

template <int DM>
__device__ __forceinline__ float sde_step_diag(bool md, float lmbd, const float* kap, float* x, const float* gv,
                                               const float* eps, float* u, float dt, float sq_ldt, float dt_l,
                                               float sq_dtl, PathAcc& acc) {
  float tmp[DM];
#pragma unroll
  for (int i = 0; i < DM; ++i) {
    u[i] = -gv[i];
    tmp[i] = __fadd_rn(dw_drift(kap[i], x[i]), u[i]);
    tmp[i] = __fadd_rn(__fmul_rn(tmp[i], dt), __fmul_rn(sq_ldt, eps[i]));
  }
  float eff_dt = dt;
  float a_l = dt_l, sq_l = sq_dtl;
  if (md) {
    const float phi0 = -x[0];
    const float xn0 = __fadd_rn(x[0], __fmul_rn(acc.alive, tmp[0]));
    const float phi1 = -xn0;
    const float still = (phi0 > 0.f && phi1 > 0.f) ? 1.f : 0.f;
    const float crossed = (phi0 > 0.f && phi1 < 0.f) ? 1.f : 0.f;
    const float frac =
        __fmul_rn(crossed, __fadd_rn(__fdiv_rn(phi0, __fadd_rn(__fadd_rn(phi0, -phi1), 1e-6f)), 1e-6f));
    const float fa = __fmul_rn(frac, acc.alive);
#pragma unroll
    for (int i = 0; i < DM; ++i) {
      const float xb = x[i];
      const float xn = __fadd_rn(xb, __fmul_rn(acc.alive, tmp[i]));
      const float xf = __fadd_rn(xb, __fmul_rn(fa, tmp[i]));
      x[i] = __fadd_rn(__fmul_rn(crossed, xf), __fmul_rn(__fadd_rn(1.f, -crossed), xn));
    }
    eff_dt = __fadd_rn(__fmul_rn(__fmul_rn(crossed, __fmul_rn(frac, frac)), dt), __fmul_rn(still, dt));
    acc.alive = (-x[0] > 0.f) ? 1.f : 0.f;
    a_l = __fdiv_rn(eff_dt, lmbd);
    sq_l = __fsqrt_rn(a_l);
  } else {
#pragma unroll
    for (int i = 0; i < DM; ++i) x[i] = __fadd_rn(x[i], __fmul_rn(acc.alive, tmp[i]));
  }
  float uu = 0.f, ue = 0.f;
#pragma unroll
  for (int i = 0; i < DM; ++i) {
    uu = __fadd_rn(uu, __fmul_rn(u[i], u[i]));
    ue = __fadd_rn(ue, __fmul_rn(u[i], eps[i]));
  }
  const float f = md ? 1.0f : 0.0f;
  acc.lw_det = __fadd_rn(acc.lw_det, __fmul_rn(a_l, __fadd_rn(-f, -__fmul_rn(0.5f, uu))));
  acc.lw_sto = __fadd_rn(acc.lw_sto, __fmul_rn(sq_l, -ue));
  return eff_dt;
}

}  // namespace socm
