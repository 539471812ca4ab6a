// target_grouped.cu -- SOCM matching target with stopping times (method.py:484-507, 524-564, 584-690).
//
// With use_stopping_time the reference evaluates M(t, s, tau_m) per SAMPLE (TwoBoundarySigmoidMLP, models.py:311-393)
// and contracts a materialised (K+1, K+1, B, d, d) table against the per-step vectors.  The table depends on the path
// only through its stopping index q_m = #{k : Phi(x_km) > 0} - 1 in {0..K} (tau_m = q_m / K, method.py:524-531), so
// there are at most K+1 distinct tables.  The host builds all of them once per iteration from the M-network
// (B-independent torch ops, differentiable: mtable.build_L_grouped) in the transposed layout
//     LT[q][c][r],  c = column of R (a_j | c_j | grad_g blocks, see target.cu), r = (i, k) row of the target,
// and the two kernels here are the block-triangular contraction of target.cu with a per-path table:
//     target[m][r]   = sum_c R[m][c] LT[q_m][c][r]                       (forward)
//     dLT[q][c][r]  += sum_{m : q_m = q} R[m][c] G[m][r]                 (backward, contraction over the paths of a group)
// Paths are visited in the order `perm` that sorts them by q (torch.sort on the device), so that neighbouring warps
// read the same table (L1 / L2 hits) and the backward reduces runs of equal q in registers before one atomic flush.
// Only columns c >= 2 d (r / d) can be non-zero (block upper-triangular, j >= i): the loops start there.
#include "kernels.h"

namespace socm {

constexpr int TG_ROWS = 128;  // target rows per CTA of the forward kernel (4 per lane)
constexpr int TG_CCH = 256;   // columns of R staged per step

// grid (ceil(B / 8), ceil(nrows / TG_ROWS)); 8 warps = 8 consecutive sorted paths
__global__ void __launch_bounds__(256) target_grouped_kernel(const float* __restrict__ LT, const float* __restrict__ R,
                                                             const int* __restrict__ group, const int* __restrict__ perm,
                                                             int B, int nrows, int kdim, int d, int ldr, int nrp,
                                                             float* __restrict__ T, int ldt) {
  __shared__ float Rs[8][TG_CCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  const int r0 = blockIdx.y * TG_ROWS;
  const bool live = s < B;
  const int m = live ? perm[s] : 0;
  const int q = live ? group[m] : 0;
  const float* tab = LT + (size_t)q * kdim * nrp;
  const float* rrow = R + (size_t)m * ldr;
  const int c_begin = 2 * d * (r0 / d);  // columns left of it are structurally zero for every row of this tile
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int c0 = c_begin; c0 < kdim; c0 += TG_CCH) {
    const int nc = min(TG_CCH, kdim - c0);
    __syncwarp();
    for (int i = lane; i < nc; i += 32) Rs[warp][i] = live ? __ldg(rrow + c0 + i) : 0.f;
    __syncwarp();
    if (live) {
      const float* tp = tab + (size_t)c0 * nrp + r0 + lane;
#pragma unroll 4
      for (int c = 0; c < nc; ++c, tp += nrp) {
        const float rv = Rs[warp][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = r0 + lane + 32 * j;
          if (r < nrows) acc[j] = fmaf(rv, __ldg(tp + 32 * j), acc[j]);
        }
      }
    }
  }
  if (live) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + lane + 32 * j;
      if (r < nrows) T[(size_t)m * ldt + r] = acc[j];
    }
  }
}

// grid (ceil(B / SEG), row tiles x column tiles); a CTA owns a 32 (rows) x 32 (columns) tile of dLT for SEG sorted paths
constexpr int TGB_SEG = 64;
__global__ void __launch_bounds__(256) target_grouped_bwd_kernel(const float* __restrict__ G, const float* __restrict__ R,
                                                                 const int* __restrict__ group,
                                                                 const int* __restrict__ perm, int B, int nrows, int kdim,
                                                                 int d, int ldr, int ldt, int nrp, int n_ctiles,
                                                                 float* __restrict__ dLT) {
  __shared__ float Gs[TGB_SEG][33], Rs[TGB_SEG][33];
  __shared__ int qs[TGB_SEG];
  const int rt = blockIdx.y / n_ctiles, ct = blockIdx.y - rt * n_ctiles;
  const int r0 = rt * 32, c0 = ct * 32;
  // the whole tile is left of the block diagonal: nothing to do
  if (c0 + 32 <= 2 * d * (r0 / d)) return;
  const int s0 = blockIdx.x * TGB_SEG;
  const int np = min(TGB_SEG, B - s0);
  for (int i = threadIdx.x; i < TGB_SEG * 32; i += 256) {
    const int p = i >> 5, j = i & 31;
    float gv = 0.f, rv = 0.f;
    if (p < np) {
      const int m = perm[s0 + p];
      if (r0 + j < nrows) gv = __ldg(G + (size_t)m * ldt + r0 + j);
      if (c0 + j < kdim) rv = __ldg(R + (size_t)m * ldr + c0 + j);
      if (j == 0) qs[p] = group[m];
    }
    Gs[p][j] = gv;
    Rs[p][j] = rv;
  }
  __syncthreads();
  const int ri = threadIdx.x & 31, cg = threadIdx.x >> 5;  // thread: row r0 + ri, columns c0 + cg + 8 j
  const int r = r0 + ri;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  auto flush = [&](int q) {
    if (r >= nrows) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + cg + 8 * j;
      if (c < kdim && c >= 2 * d * (r / d) && acc[j] != 0.f) atomicAdd(dLT + ((size_t)q * kdim + c) * nrp + r, acc[j]);
      acc[j] = 0.f;
    }
  };
  int q_run = qs[0];
  for (int p = 0; p < np; ++p) {
    const int q = qs[p];
    if (q != q_run) {  // uniform over the block: qs is shared
      flush(q_run);
      q_run = q;
    }
    const float gv = Gs[p][ri];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = fmaf(gv, Rs[p][cg + 8 * j], acc[j]);
  }
  flush(q_run);
}

}  // namespace socm

using namespace socm;

extern "C" int socm_target_grouped_f32(const float* LT, const float* R, const int32_t* group, const int32_t* perm,
                                       int32_t n_groups, int32_t B, int32_t K, int32_t d, int32_t ldr, int32_t nrp,
                                       float* target, int32_t ldt, void* stream_) {
  SOCM_CHECK_ARG(LT && R && group && perm && target, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1 && n_groups >= 1, "bad sizes");
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  SOCM_CHECK_ARG(ldr >= kdim && ldt >= nrows && nrp >= nrows, "bad pitches ldr=%d ldt=%d nrp=%d", ldr, ldt, nrp);
  if (B == 0) return SOCM_OK;
  dim3 grid((B + 7) / 8, (nrows + TG_ROWS - 1) / TG_ROWS);
  target_grouped_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(LT, R, group, perm, B, nrows, kdim, d, ldr,
                                                                              nrp, target, ldt);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_target_grouped_bwd_f32(const float* G, const float* R, const int32_t* group, const int32_t* perm,
                                           int32_t n_groups, int32_t B, int32_t K, int32_t d, int32_t ldr, int32_t ldt,
                                           int32_t nrp, float* dLT, void* stream_) {
  SOCM_CHECK_ARG(G && R && group && perm && dLT, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1 && n_groups >= 1, "bad sizes");
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  SOCM_CHECK_ARG(ldr >= kdim && ldt >= nrows && nrp >= nrows, "bad pitches ldr=%d ldt=%d nrp=%d", ldr, ldt, nrp);
  if (B == 0) return SOCM_OK;
  const int n_rt = (nrows + 31) / 32, n_ct = (kdim + 31) / 32;
  SOCM_CHECK_ARG((int64_t)n_rt * n_ct <= 65535, "(K+1)d x (2K+1)d = %d x %d: too many tiles for one launch", nrows, kdim);
  dim3 grid((B + TGB_SEG - 1) / TGB_SEG, n_rt * n_ct);
  target_grouped_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(G, R, group, perm, B, nrows, kdim, d,
                                                                                  ldr, ldt, nrp, n_ct, dLT);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
