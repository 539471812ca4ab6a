// target.cu -- SOCM matching target (method.py:584-690) in the re-associated form of
// SURVEY.md A.3:   target = R L^T,   R = per-path right-hand side [a_0 c_0 a_1 c_1 ... grad_g],
// L = block upper-triangular table of M_t(s) and d/ds M_t(s).  The reference's
// (K+1, K+1, B, d, d) intermediate never exists.
//
//   target_prep_kernel      builds R and the importance weights w from the rollout outputs
//   target_gemm_kernel      target[B][ldt] = R L^T        (only K-blocks with j >= i are read)
//   target_gemm_bwd_kernel  dL = G^T R                    (contraction over paths, split-K)
//   target_const_m_kernel   SOCM_const_M: suffix sums (method.py:289-369)
//   target_adjoint_kernel   SOCM_adjoint: backward recursion of the adjoint state per path (method.py:722-749)
//   weight_stats_kernel     sum w, sum w^2, sum stop with a warp-shuffle block reduction
// The two GEMMs here are the FP32 SIMT versions (parity path).
#include "kernels.h"

namespace socm {

// ---------------------------------------------------------------- R and w
// Per (step j, path m):  c = sigma^{-T}(sqrt(lmbd eff) eps + eff u),  a = eff grad_f - grad_b . c   
__device__ __forceinline__ void prep_point(const socm_setting& st, float sq_lmbd, const float* __restrict__ states,
                                           const float* __restrict__ noises, const float* __restrict__ controls,
                                           const float* __restrict__ eff_dt, int B, int j, int m, float* a_out,
                                           float* c_out) {
  const int d = st.d;
  float x[kMaxDim], t0[kMaxDim], t1[kMaxDim];
  const float* xs = states + ((size_t)j * B + m) * d;
  for (int i = 0; i < d; ++i) x[i] = __ldg(xs + i);
  const float eff = __ldg(eff_dt + (size_t)j * B + m);
  const float ce = sq_lmbd * sqrtf(eff);
  const float* ep = noises + ((size_t)j * B + m) * d;
  const float* up = controls + ((size_t)j * B + m) * d;
  if (st.sigma_is_identity) {
    for (int i = 0; i < d; ++i) c_out[i] = fmaf(ce, __ldg(ep + i), eff * __ldg(up + i));
  } else {
    for (int i = 0; i < d; ++i) t0[i] = fmaf(ce, __ldg(ep + i), eff * __ldg(up + i));
    matvec_t(st.sigma_inv, d, t0, c_out);  // sigma^{-T} (.)
  }
  grad_run_cost(st, x, 1, t0);
  grad_drift_dot(st, x, 1, c_out, t1);
  for (int i = 0; i < d; ++i) a_out[i] = fmaf(eff, t0[i], -t1[i]);
}

// R[m][2jd .. 2jd+2d) = [a_j c_j] for j < K.  A block owns 32 paths x TJ steps: the inputs are read with the
// paths along the lanes (states / noises / controls are [step][path][d]: coalesced), staged in shared memory
// and written with the row of a path along the lanes (R is [path][ldr]: TJ * 2d contiguous floats per path),
// so both sides move whole 128-byte lines.  HBM-bound: 12d + 4 bytes read, 8d bytes written per (j, m).
constexpr int PREP_TM = 32, PREP_ROW = 320;  // paths per block, floats of a staged row (TJ = PREP_ROW / 2d steps)
__global__ void __launch_bounds__(256) target_prep_kernel(socm_setting st, const float* __restrict__ states,
                                                          const float* __restrict__ noises,
                                                          const float* __restrict__ controls,
                                                          const float* __restrict__ eff_dt, int B, int K,
                                                          float* __restrict__ R, int ldr) {
  __shared__ float buf[PREP_TM][PREP_ROW + 1];  // +1: the lanes of phase 1 hit different banks
  const int d = st.d, TJ = PREP_ROW / (2 * d);
  const float sq_lmbd = sqrtf(st.lmbd);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_mt = (B + PREP_TM - 1) / PREP_TM, n_jt = (K + TJ - 1) / TJ;
  for (int tile = blockIdx.x; tile < n_mt * n_jt; tile += gridDim.x) {
    const int jt = tile / n_mt, mt = tile - jt * n_mt;  // consecutive blocks: consecutive path tiles of one step range
    const int j0 = jt * TJ, m0 = mt * PREP_TM;
    const int nj = min(TJ, K - j0);
    const int m = m0 + lane;
    for (int tj = warp; tj < nj; tj += 8) {
      if (m < B) {
        float a[kMaxDim], c[kMaxDim];
        prep_point(st, sq_lmbd, states, noises, controls, eff_dt, B, j0 + tj, m, a, c);
        for (int i = 0; i < d; ++i) {
          buf[lane][tj * 2 * d + i] = a[i];
          buf[lane][tj * 2 * d + d + i] = c[i];
        }
      }
    }
    __syncthreads();
    const int row_len = nj * 2 * d;
    for (int r = warp; r < PREP_TM && m0 + r < B; r += 8) {
      float* dst = R + (size_t)(m0 + r) * ldr + (size_t)2 * j0 * d;
      for (int i = lane; i < row_len; i += 32) dst[i] = buf[r][i];
    }
    __syncthreads();
  }
}

// last column block of R (grad_g at the terminal state), the pitch padding, and the importance weights
__global__ void __launch_bounds__(256) target_prep_tail_kernel(socm_setting st, const float* __restrict__ states,
                                                               const float* __restrict__ lw_det,
                                                               const float* __restrict__ lw_sto,
                                                               const float* __restrict__ lw_term, int B, int K,
                                                               float* __restrict__ R, int ldr, float* __restrict__ w) {
  const int d = st.d;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < B; m += gridDim.x * blockDim.x) {
    float x[kMaxDim], t0[kMaxDim];
    const float* xs = states + ((size_t)K * B + m) * d;
    for (int i = 0; i < d; ++i) x[i] = __ldg(xs + i);
    float* row = R + (size_t)m * ldr;
    grad_term_cost(st, x, 1, t0);
    for (int i = 0; i < d; ++i) row[2 * K * d + i] = t0[i];
    for (int i = (2 * K + 1) * d; i < ldr; ++i) row[i] = 0.f;  // pitch padding
    if (w) w[m] = expf(__fadd_rn(__fadd_rn(__ldg(lw_det + m), __ldg(lw_sto + m)), __ldg(lw_term + m)));
  }
}

// ---------------------------------------------------------------- target = R L^T  (NT SGEMM, 64x64x16)
constexpr int GT = 64, GK = 16;

__global__ void __launch_bounds__(256) target_gemm_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                                          int B, int nrows /*(K+1)d*/, int kdim /*(2K+1)d*/, int d,
                                                          int ldr, float* __restrict__ T, int ldt) {
  __shared__ __align__(16) float As[GK][GT + 4];  // R tile, k-major
  __shared__ __align__(16) float Bs[GK][GT + 4];  // L tile, k-major
  const int m0 = blockIdx.x * GT, n0 = blockIdx.y * GT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // rows (i,k) of L with i >= i_min are zero left of column 2*i_min*d
  const int k_begin = ((2 * (n0 / d) * d) / GK) * GK;
  float acc[4][4] = {};
  const int lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;  // loader: 64 rows x 4 float4
  for (int k0 = k_begin; k0 < kdim; k0 += GK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    const int kk = k0 + lk;
    if (m0 + lrow < B && kk < ldr) a = __ldg(reinterpret_cast<const float4*>(R + (size_t)(m0 + lrow) * ldr + kk));
    if (n0 + lrow < nrows && kk < ldr) b = __ldg(reinterpret_cast<const float4*>(L + (size_t)(n0 + lrow) * ldr + kk));
    if (kk + 0 >= kdim) a.x = b.x = 0.f;
    if (kk + 1 >= kdim) a.y = b.y = 0.f;
    if (kk + 2 >= kdim) a.z = b.z = 0.f;
    if (kk + 3 >= kdim) a.w = b.w = 0.f;
    As[lk + 0][lrow] = a.x; As[lk + 1][lrow] = a.y; As[lk + 2][lrow] = a.z; As[lk + 3][lrow] = a.w;
    Bs[lk + 0][lrow] = b.x; Bs[lk + 1][lrow] = b.y; Bs[lk + 2][lrow] = b.z; Bs[lk + 3][lrow] = b.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < nrows) T[(size_t)m * ldt + n] = acc[i][j];
    }
  }
}

// ---------------------------------------------------------------- dL = G^T R  (TN SGEMM, split over paths)
__global__ void __launch_bounds__(256) target_gemm_bwd_kernel(const float* __restrict__ G, const float* __restrict__ R,
                                                              int B, int nrows, int kdim, int d, int ldr, int ldt,
                                                              int m_per_split, float* __restrict__ dL, int use_atomic) {
  __shared__ __align__(16) float As[GK][GT + 4];  // G tile [m][n]
  __shared__ __align__(16) float Bs[GK][GT + 4];  // R tile [m][r]
  const int n0 = blockIdx.x * GT, r0 = blockIdx.y * GT;
  // block (i, j) is needed only for j >= i: skip tiles entirely left of the diagonal
  const int i_min = n0 / d;
  if (r0 + GT <= 2 * i_min * d) return;
  const int m_begin = blockIdx.z * m_per_split;
  const int m_end = min(B, m_begin + m_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lm = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;  // loader: 16 rows x 16 float4
  float acc[4][4] = {};
  for (int mm = m_begin; mm < m_end; mm += GK) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    const int m = mm + lm;
    if (m < m_end) {
      const int n = n0 + lc, r = r0 + lc;
      if (n + 3 < ldt) {
        a = __ldg(reinterpret_cast<const float4*>(G + (size_t)m * ldt + n));
      }
      if (n + 0 >= nrows) a.x = 0.f;
      if (n + 1 >= nrows) a.y = 0.f;
      if (n + 2 >= nrows) a.z = 0.f;
      if (n + 3 >= nrows) a.w = 0.f;
      if (r + 3 < ldr) b = __ldg(reinterpret_cast<const float4*>(R + (size_t)m * ldr + r));
      if (r + 0 >= kdim) b.x = 0.f;
      if (r + 1 >= kdim) b.y = 0.f;
      if (r + 2 >= kdim) b.z = 0.f;
      if (r + 3 >= kdim) b.w = 0.f;
    }
    *reinterpret_cast<float4*>(&As[lm][lc]) = a;
    *reinterpret_cast<float4*>(&Bs[lm][lc]) = b;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= nrows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + tx * 4 + j;
      if (r >= kdim || r < 2 * (n / d) * d) continue;  // structurally zero blocks (j < i) stay zero
      float* dst = dL + (size_t)n * ldr + r;
      if (use_atomic) atomicAdd(dst, acc[i][j]);
      else *dst += acc[i][j];
    }
  }
}

// ---------------------------------------------------------------- SOCM_const_M: suffix sums
__global__ void __launch_bounds__(256) target_const_m_kernel(const float* __restrict__ R, int B, int K, int d, int ldr,
                                                             float* __restrict__ T, int ldt) {
  const size_t total = (size_t)B * d;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(idx / d), k = (int)(idx - (size_t)m * d);
    const float* row = R + (size_t)m * ldr;
    float* out = T + (size_t)m * ldt;
    float acc = row[2 * K * d + k];
    out[K * d + k] = acc;
    for (int i = K - 1; i >= 0; --i) {
      acc += row[2 * i * d + k];
      out[i * d + k] = acc;
    }
  }
}

// ---------------------------------------------------------------- sum w, sum w^2, sum stop
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- SOCM_adjoint target (method.py:722-735)
// One thread per path walks the grid backwards:  a_K = grad_g(x_K),
//   a_j = a_{j+1} + dt * ((grad_f_j + grad_f_{j+1}) / 2 + ((grad_b_j + grad_b_{j+1}) / 2) a_{j+1})
// grad_b is contracted on its last index ("mkl,ml->mk"); the OU settings return the constant A^T
// (so the average is A^T itself), double well / molecular dynamics a diagonal.
__global__ void __launch_bounds__(128) target_adjoint_kernel(socm_setting st, const float* __restrict__ states, int B,
                                                             int K, float dt, float* __restrict__ target, int ldt) {
  const int d = st.d;
  const bool ou = st.kind == SOCM_OU_QUADRATIC || st.kind == SOCM_OU_LINEAR;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < B; m += gridDim.x * blockDim.x) {
    float a[kMaxDim], x1[kMaxDim], x0[kMaxDim], f1[kMaxDim], f0[kMaxDim], t[kMaxDim];
    const float* xs = states + ((size_t)K * B + m) * d;
    for (int i = 0; i < d; ++i) x1[i] = __ldg(xs + i);
    grad_term_cost(st, x1, 1, a);
    grad_run_cost(st, x1, 1, f1);
    float* row = target + (size_t)m * ldt;
    for (int i = 0; i < d; ++i) row[K * d + i] = a[i];
    for (int j = K - 1; j >= 0; --j) {
      xs = states + ((size_t)j * B + m) * d;
      for (int i = 0; i < d; ++i) x0[i] = __ldg(xs + i);
      grad_run_cost(st, x0, 1, f0);
      if (ou) {
        grad_drift_dot(st, x0, 1, a, t);
      } else {
        for (int l = 0; l < d; ++l) {
          const float kap = __ldg(st.kappa + l);
          const float g = __fmul_rn(__fadd_rn(dw_drift_diag_grad(kap, x0[l]), dw_drift_diag_grad(kap, x1[l])), 0.5f);
          t[l] = __fmul_rn(g, a[l]);
        }
      }
      for (int i = 0; i < d; ++i) {
        const float inc = __fadd_rn(__fmul_rn(__fadd_rn(f0[i], f1[i]), 0.5f), t[i]);
        a[i] = __fadd_rn(a[i], __fmul_rn(dt, inc));
        row[j * d + i] = a[i];
        x1[i] = x0[i];
        f1[i] = f0[i];
      }
    }
    for (int i = (K + 1) * d; i < ldt; ++i) row[i] = 0.f;  // pitch padding
  }
}

__global__ void __launch_bounds__(256) weight_stats_kernel(const float* __restrict__ w, const float* __restrict__ stop,
                                                           int B, int K, double* __restrict__ sums) {
  __shared__ double red[3][8];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = tid; i < (size_t)B; i += stride) {
    const double v = (double)__ldg(w + i);
    s0 += v;
    s1 += v * v;
  }
  if (stop)
    for (size_t i = tid; i < (size_t)(K + 1) * B; i += stride) s2 += (double)__ldg(stop + i);
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = s0;
    red[1][warp] = s1;
    red[2][warp] = s2;
  }
  __syncthreads();
  if (warp == 0) {
    for (int q = 0; q < 3; ++q) {
      double v = lane < 8 ? red[q][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(sums + q, v);
    }
  }
}

}  // namespace socm

// ================================================================ C ABI
using namespace socm;

static int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)sm_count() * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

extern "C" int socm_target_prep_f32(const socm_setting* st, const float* states, const float* noises,
                                    const float* controls, const float* eff_dt, const float* logw_det,
                                    const float* logw_sto, const float* logw_term, int32_t B, int32_t K, float* R,
                                    int32_t ldr, float* w, void* stream_) {
  if (int rc = validate_setting(st)) return rc;
  SOCM_CHECK_ARG(states && noises && controls && eff_dt && R, "required pointer is NULL");
  SOCM_CHECK_ARG(!w || (logw_det && logw_sto && logw_term), "w requested but log-weights missing");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * st->d && ldr % 4 == 0, "ldr=%d must be >= (2K+1)d and a multiple of 4", ldr);
  if (B == 0) return SOCM_OK;
  SOCM_CHECK_ARG(2 * st->d <= PREP_ROW, "d=%d too large for the staged rows", st->d);
  const int tj = PREP_ROW / (2 * st->d);
  const size_t tiles = (size_t)((B + PREP_TM - 1) / PREP_TM) * ((K + tj - 1) / tj);
  target_prep_kernel<<<grid_for(tiles * 256, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      *st, states, noises, controls, eff_dt, B, K, R, ldr);
  SOCM_LAUNCH_CHECK();
  target_prep_tail_kernel<<<grid_for((size_t)B, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      *st, states, logw_det, logw_sto, logw_term, B, K, R, ldr, w);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_target_gemm_f32(const float* L, const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                                    float* target, int32_t ldt, void* stream_) {
  SOCM_CHECK_ARG(L && R && target, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1, "bad sizes");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * d && ldr % 4 == 0 && ldt >= (K + 1) * d, "bad pitches ldr=%d ldt=%d", ldr, ldt);
  if (B == 0) return SOCM_OK;
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  dim3 grid((B + GT - 1) / GT, (nrows + GT - 1) / GT);
  target_gemm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(L, R, B, nrows, kdim, d, ldr, target, ldt);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_target_gemm_bwd_f32(const float* G, const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                                        int32_t ldt, float* dL, int32_t accumulate, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SOCM_CHECK_ARG(G && R && dL, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1, "bad sizes");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * d && ldr % 4 == 0 && ldt >= (K + 1) * d && ldt % 4 == 0,
                 "bad pitches ldr=%d ldt=%d (both must be multiples of 4)", ldr, ldt);
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  if (!accumulate) SOCM_CUDA(cudaMemsetAsync(dL, 0, (size_t)nrows * ldr * sizeof(float), stream));
  if (B == 0) return SOCM_OK;
  const int tiles = ((nrows + GT - 1) / GT) * ((kdim + GT - 1) / GT);
  int splits = (4 * sm_count() + tiles - 1) / tiles;
  const int max_splits = (B + 4 * GK - 1) / (4 * GK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int m_per_split = (B + splits - 1) / splits;
  m_per_split = ((m_per_split + GK - 1) / GK) * GK;
  splits = (B + m_per_split - 1) / m_per_split;
  dim3 grid((nrows + GT - 1) / GT, (kdim + GT - 1) / GT, splits);
  target_gemm_bwd_kernel<<<grid, 256, 0, stream>>>(G, R, B, nrows, kdim, d, ldr, ldt, m_per_split, dL, splits > 1);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_target_const_m_f32(const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr, float* target,
                                       int32_t ldt, void* stream_) {
  SOCM_CHECK_ARG(R && target && d >= 1 && K >= 1, "bad arguments");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * d && ldt >= (K + 1) * d, "bad pitches");
  if (B == 0) return SOCM_OK;
  target_const_m_kernel<<<grid_for((size_t)B * d, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(R, B, K, d, ldr,
                                                                                                       target, ldt);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_target_adjoint_f32(const socm_setting* st, const float* states, int32_t B, int32_t K, float dt,
                                       float* target, int32_t ldt, void* stream_) {
  if (int rc = validate_setting(st)) return rc;
  SOCM_CHECK_ARG(states && target && K >= 1, "bad arguments");
  SOCM_CHECK_ARG(ldt >= (K + 1) * st->d, "bad pitch ldt=%d", ldt);
  if (B == 0) return SOCM_OK;
  target_adjoint_kernel<<<grid_for((size_t)B, 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(*st, states, B, K, dt,
                                                                                               target, ldt);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_weight_stats_f32(const float* w, const float* stop, int32_t B, int32_t K, double* sums,
                                     void* stream_) {
  SOCM_CHECK_ARG(w && sums && B >= 0, "bad arguments");
  if (B == 0) return SOCM_OK;
  weight_stats_kernel<<<grid_for((size_t)B, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(w, stop, B, K, sums);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
