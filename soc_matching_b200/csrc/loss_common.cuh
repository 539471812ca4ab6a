// loss_common.cuh -- argument block and per-point loss shared by the generic and tiled K3 kernels.
#pragma once
#include "common.cuh"

namespace socm {

struct LossArgs {
  socm_setting st;
  const float* warmA;  // [K+1][d][d] or NULL
  const float* warmc;  // [K+1][d]
  const float* ts;     // [K+1]
  const float* states; // [K+1][B][d]
  const float* target; // [B][ldt]
  const float* w;      // [B]
  const float* stop;   // [K+1][B] or NULL
  float scale;
  int B, K, ldt;
  float* G;            // [B][ldt]
  double* loss_sums;
};

// Per-point loss and d loss / d nabla_V.  v[d] = UNet output; returns the loss term and writes
// dv[d]; G row gets -dv.  (method.py:280-287, 692-720)
__device__ __forceinline__ float point_loss(const LossArgs& a, int i, int m, const float* x, int ldx,
                                            const float* v, int ldv, float* dv) {
  const socm_setting& st = a.st;
  const int d = st.d;
  float diff[kMaxDim], r[kMaxDim];
  for (int j = 0; j < d; ++j) diff[j] = v[j * ldv];
  if (a.warmA != nullptr) {
    // nabla_V - sigma^{-T} u_ws(t_i, x),  u_ws = sigma^{-1}(c_i + A_i x - b(x))
    float uws[kMaxDim];
    for (int j = 0; j < d; ++j) uws[j] = 0.f;
    add_warm_start(st, a.warmA + (size_t)i * d * d, a.warmc + (size_t)i * d, x, ldx, uws);
    if (st.sigma_is_identity) {
      for (int j = 0; j < d; ++j) diff[j] -= uws[j];
    } else {
      matvec_t(st.sigma_inv, d, uws, r);
      for (int j = 0; j < d; ++j) diff[j] -= r[j];
    }
  }
  const float* trow = a.target + (size_t)m * a.ldt + (size_t)i * d;
  for (int j = 0; j < d; ++j) diff[j] -= __ldg(trow + j);
  const float s = a.stop ? __ldg(a.stop + (size_t)i * a.B + m) : 1.f;
  const float coef = s * __ldg(a.w + m) * a.scale;
  float sq = 0.f;
  if (st.sigma_is_identity) {
    for (int j = 0; j < d; ++j) {
      sq = fmaf(diff[j], diff[j], sq);
      dv[j] = 2.f * coef * diff[j];
    }
  } else {
    matvec_t(st.sigma, d, diff, r);  // r = sigma^T diff
    for (int j = 0; j < d; ++j) sq = fmaf(r[j], r[j], sq);
    float t[kMaxDim];
    matvec(st.sigma, d, r, t);       // sigma sigma^T diff
    for (int j = 0; j < d; ++j) dv[j] = 2.f * coef * t[j];
  }
  float* grow = a.G + (size_t)m * a.ldt + (size_t)i * d;
  for (int j = 0; j < d; ++j) grow[j] = -dv[j];
  return coef * sq;
}

}  // namespace socm
