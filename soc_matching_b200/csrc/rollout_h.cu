// rollout_h.cu -- K1 on the fp16-split tcgen05 engine (unet_h.cuh): the Euler-Maruyama rollout of
// utils.stochastic_trajectories (utils.py:17-128), control network on the tensor cores, two CTAs per SM.
//
// One CTA owns a tile of 128 paths for all K steps and 256 TMEM columns; a second CTA on the same SM runs its own tile,
// and the tensor pipe executes the MMAs of whichever tile has work: the epilogue / SDE-step phases of one tile are
// hidden behind the MMA phases of the other.  Per CTA (320 threads):
//   warps 0-7  "E": thread <-> (path p = tid & 127, column half h = tid >> 7).  Bias + ReLU + scale + fp16 hi/lo split of
//              every activation; half 0 owns the path state and runs the SDE step of common.cuh (same operation order
//              as utils.py:45-99, stopping indices bit-exact), half 1 keeps the folded Wc r1 term, draws the Philox
//              noise and finishes the last layer.
//   warp 8     "M": issues every tcgen05.mma of the step (one elected lane).
//   warp 9     "P": streams the weight tape from L2 into a 3-stage ring with cp.async.bulk.
#include <cuda_fp16.h>

#include <type_traits>

#include "kernels.h"
#include "rollout_common.cuh"
#include "unet_generic.cuh"
#include "unet_h.cuh"

namespace socm {
namespace hx {

using namespace umma;

// ---------------------------------------------------------------- per-call setup kernels
// Wc = W_u0 W_r1 [16 x 256] (rows >= d zero) and bc = W_u0 b_r1 + b_u0 (unet_tc.cuh, "folding"); fp64 accumulation
__global__ void fold_h_kernel(socm_unet net, float* __restrict__ wc, float* __restrict__ small) {
  const int d = net.d;
  const int j = blockIdx.x, g = threadIdx.x;  // NY blocks x 256 threads
  double acc = 0.0;
  if (j < d)
    for (int f = 0; f < H0; ++f) acc += (double)net.w[8][(size_t)j * H0 + f] * (double)net.w[4][(size_t)f * H0 + g];
  wc[j * H0 + g] = (float)acc;
  if (g == 0) {
    double b = 0.0;
    if (j < d) {
      b = (double)net.b[8][j];
      for (int f = 0; f < H0; ++f) b += (double)net.w[8][(size_t)j * H0 + f] * (double)net.b[4][f];
    }
    small[small_layout().bc + j] = (float)b;
  }
}

// max |.| of a float array into a uint32 slot (values are >= 0 as bit patterns: ordered like the floats)
__device__ __forceinline__ void warp_max_to(uint32_t* slot, float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(slot, __float_as_uint(v));
}

// Calibration: (a) max |w| per layer (blocks 0..9), (b) max |activation| of the six calibrated layer inputs over
// n_samples sample points evaluated in fp32 (one 256-thread block per point).  Sample points:
//   rollout mode (states == nullptr): x0 of the first paths, spread by r sqrt(lmbd T) z with r in {0, 1/2, 1, 2} and
//       Philox normal z, at t in {0, T/3, 2T/3, T} -- the scale only has to be right within ~2^10 (head room) upwards
//       and ~2^15 downwards (precision floor), see unet_h.cuh;
//   loss mode: points of the stored trajectories, strided over all (K+1) B of them.
constexpr int CALIB_BWD_FLOATS = 32 + 2 * H0 + 2 * H1 + H2 + H0;
__global__ void __launch_bounds__(256) calib_h_kernel(socm_unet net, const float* __restrict__ wc, CalibArgs c,
                                                      uint32_t* __restrict__ mx) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* smem = reinterpret_cast<float*>(smem_raw);
  const int d = net.d;
  if (blockIdx.x == 10) {  // max |w_m| over all paths (exact: the global scale of the loss gradient hangs on it)
    float v = 0.f;
    if (c.w != nullptr)
      for (int i = threadIdx.x; i < c.B; i += blockDim.x) v = fmaxf(v, fabsf(__ldg(c.w + i)));
    warp_max_to(mx + MX_W, v);
    return;
  }
  if (blockIdx.x < 10) {  // weight maxima
    const int l = blockIdx.x;
    const int nout[10] = {H0, H1, H2, d, H0, H1, H1, H0, d, d};
    const int nin[10] = {d + 1, H0, H1, d + 1, H0, H1, H2, H1, H0, H0};
    const float* W = l == WC_LAYER ? wc : net.w[l];
    float v = 0.f;
    for (int i = threadIdx.x; i < nout[l] * nin[l]; i += blockDim.x) v = fmaxf(v, fabsf(W[i]));
    warp_max_to(mx + N_ACT + l, v);
    return;
  }
  // ---- one sample per block: thread n owns output n of a layer (forward) / input k (transposed products)
  const int tid = threadIdx.x;
  generic::FwdBuf b = generic::carve_fwd(smem, d, H0, H1, H2);
  const int i = blockIdx.x - 11;
  if (i >= c.n_samples) return;
  size_t pt = 0;
  int ti = 0;
  if (c.states == nullptr) {
    const int nb = c.B < 64 ? c.B : 64;
    const int m = i % nb, v = (i / nb) & 3;
    const float T = __ldg(c.step_tab + 4 * c.K + c.K - 1) + __ldg(c.step_tab + c.K - 1);
    const float r = (v == 0 ? 0.f : (v == 1 ? 0.5f : (v == 2 ? 1.f : 2.f))) * sqrtf(c.lmbd * T);
    if (tid == 0) b.xin[0] = T * (float)v / 3.f;
    if (tid * 4 < d) {
      float z[4];
      // fixed key: the calibration (hence the power-of-two scales, hence every rounding) must not depend on the
      // call's noise seed, or replaying a run with its own noise injected would not be bit-identical
      philox_normal4(0x5bd1e995c0ffee11ull, (uint64_t)i, 0xffffu, (uint32_t)tid, z);
      for (int j = 0; j < 4 && tid * 4 + j < d; ++j)
        b.xin[1 + tid * 4 + j] = __ldg(c.x0 + (size_t)m * d + tid * 4 + j) + r * z[j];
    }
  } else {
    const size_t n_pts = (size_t)(c.K + 1) * c.B;
    pt = (size_t)(((double)i + 0.5) / c.n_samples * (double)n_pts);
    ti = (int)(pt / c.B);
    if (tid == 0) b.xin[0] = __ldg(c.ts + ti);
    if (tid < d) b.xin[1 + tid] = __ldg(c.states + pt * d + tid);
  }
  __syncthreads();
  // out[n] = bias[n] + sum_k W[n][k] in[k]  (four partial sums: the loads of a row are independent, the chain is short)
  auto dense = [&](int l, int nout, int nin, const float* in, float* out) {
    for (int n = tid; n < nout; n += blockDim.x) {
      const float* row = net.w[l] + (size_t)n * nin;
      float a0 = __ldg(net.b[l] + n), a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int k = 0;
      for (; k + 4 <= nin; k += 4) {
        a0 = fmaf(__ldg(row + k), in[k], a0);
        a1 = fmaf(__ldg(row + k + 1), in[k + 1], a1);
        a2 = fmaf(__ldg(row + k + 2), in[k + 2], a2);
        a3 = fmaf(__ldg(row + k + 3), in[k + 3], a3);
      }
      for (; k < nin; ++k) a0 = fmaf(__ldg(row + k), in[k], a0);
      out[n] = (a0 + a1) + (a2 + a3);
    }
  };
  // din[k] (+)= sum_n W[n][k] dout[n]  (coalesced over k)
  auto dense_t = [&](int l, int nout, int nin, const float* dout, float* din, bool accumulate) {
    for (int k = tid; k < nin; k += blockDim.x) {
      const float* col = net.w[l] + k;
      float a0 = accumulate ? din[k] : 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int n = 0;
      for (; n + 4 <= nout; n += 4) {
        a0 = fmaf(__ldg(col + (size_t)n * nin), dout[n], a0);
        a1 = fmaf(__ldg(col + (size_t)(n + 1) * nin), dout[n + 1], a1);
        a2 = fmaf(__ldg(col + (size_t)(n + 2) * nin), dout[n + 2], a2);
        a3 = fmaf(__ldg(col + (size_t)(n + 3) * nin), dout[n + 3], a3);
      }
      for (; n < nout; ++n) a0 = fmaf(__ldg(col + (size_t)n * nin), dout[n], a0);
      din[k] = (a0 + a1) + (a2 + a3);
    }
  };
  auto relu = [&](float* v, int n) {
    if (tid < n) v[tid] = fmaxf(v[tid], 0.f);
  };
  // max over the block's first n values (n <= 256 = blockDim.x)
  auto amax_to = [&](uint32_t* slot, const float* v, int n, bool relu_v) {
    const float a = tid < n ? (relu_v ? v[tid] : fabsf(v[tid])) : 0.f;
    warp_max_to(slot, a);
  };
  // forward (models.py:202-242), as unet_generic.cuh
  dense(0, H0, d + 1, b.xin, b.r1);
  __syncthreads();
  relu(b.r1, H0);
  __syncthreads();
  dense(1, H1, H0, b.r1, b.r2);
  dense(4, H0, H0, b.r1, b.o1);
  __syncthreads();
  relu(b.r2, H1);
  __syncthreads();
  dense(2, H2, H1, b.r2, b.r3);
  dense(5, H1, H1, b.r2, b.o2);
  __syncthreads();
  relu(b.r3, H2);
  __syncthreads();
  dense(6, H1, H2, b.r3, b.y2);
  __syncthreads();
  if (tid < H1) b.o2[tid] += fmaxf(b.y2[tid], 0.f);
  __syncthreads();
  dense(7, H0, H1, b.o2, b.y1);
  __syncthreads();
  if (tid < H0) b.o1[tid] += fmaxf(b.y1[tid], 0.f);
  __syncthreads();
  dense(8, d, H0, b.o1, b.y0);
  dense(3, d, d + 1, b.xin, b.o0);
  __syncthreads();
  if (tid < d) b.o0[tid] += fmaxf(b.y0[tid], 0.f);
  amax_to(mx + A_X, b.xin, d + 1, false);
  amax_to(mx + A_R1, b.r1, H0, false);
  amax_to(mx + A_R2, b.r2, H1, false);
  amax_to(mx + A_R3, b.r3, H2, false);
  amax_to(mx + A_O2, b.o2, H1, false);
  amax_to(mx + A_Y1, b.y1, H0, true);   // y1 holds the pre-ReLU values of up_1
  __syncthreads();
  if (c.target == nullptr || c.states == nullptr) return;
  // backward gains: max |d_layer| for the row-normalised loss gradient dv / max|dv| at this point (loss_h.cu);
  // dv ~ nabla_V - target up to a positive factor (sigma and the warm start only turn it slightly)
  float* bw = smem + generic::fwd_floats(d, H0, H1, H2);
  float* d_y0 = bw;            // [d]; bw[31] = max |dv|
  float* d_o1 = d_y0 + 32;     // [H0]
  float* d_o2 = d_o1 + H0;     // [H1]
  float* d_r2 = d_o2 + H1;     // [H1]
  float* d_r3 = d_r2 + H1;     // [H2]
  float* d_r1 = d_r3 + H2;     // [H0]
  const int m = (int)(pt - (size_t)ti * c.B);
  if (tid < 32) {
    float dvm = 0.f;
    for (int j = tid; j < d; j += 32) {
      const float dv = b.o0[j] - __ldg(c.target + (size_t)m * c.ldt + (size_t)ti * d + j);
      d_y0[j] = dv;
      dvm = fmaxf(dvm, fabsf(dv));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dvm = fmaxf(dvm, __shfl_xor_sync(0xffffffffu, dvm, o));
    if (tid == 0) {
      if (dvm > 0.f) atomicMax(mx + MX_DIFF, __float_as_uint(dvm));
      bw[31] = dvm;
    }
    __syncwarp();
    for (int j = tid; j < d; j += 32) d_y0[j] = (dvm > 0.f && b.y0[j] > 0.f) ? d_y0[j] / dvm : 0.f;
  }
  __syncthreads();
  if (!(bw[31] > 0.f)) return;
  dense_t(8, d, H0, d_y0, d_o1, false);   // d_o1 = W_u0^T d_y0
  __syncthreads();
  dense_t(4, H0, H0, d_o1, d_r1, false);  // W_r1^T d_o1 (kept for d_z1)
  __syncthreads();
  if (tid < H0) d_o1[tid] = b.y1[tid] > 0.f ? d_o1[tid] : 0.f;   // d_y1
  __syncthreads();
  amax_to(mx + MX_B + B_DY1, d_o1, H0, false);
  dense_t(7, H0, H1, d_o1, d_o2, false);
  __syncthreads();
  amax_to(mx + MX_B + B_DO2, d_o2, H1, false);
  dense_t(5, H1, H1, d_o2, d_r2, false);
  __syncthreads();
  if (tid < H1) d_o2[tid] = b.y2[tid] > 0.f ? d_o2[tid] : 0.f;   // d_y2
  __syncthreads();
  dense_t(6, H1, H2, d_o2, d_r3, false);
  __syncthreads();
  if (tid < H2) d_r3[tid] = b.r3[tid] > 0.f ? d_r3[tid] : 0.f;   // d_z3
  __syncthreads();
  amax_to(mx + MX_B + B_DZ3, d_r3, H2, false);
  dense_t(2, H2, H1, d_r3, d_r2, true);
  __syncthreads();
  if (tid < H1) d_r2[tid] = b.r2[tid] > 0.f ? d_r2[tid] : 0.f;   // d_z2
  __syncthreads();
  amax_to(mx + MX_B + B_DZ2, d_r2, H1, false);
  dense_t(1, H1, H0, d_r2, d_r1, true);   // d_z1 = m_r1 . (W_d1^T d_z2 + W_r1^T d_o1)
  __syncthreads();
  if (tid < H0) d_r1[tid] = b.r1[tid] > 0.f ? d_r1[tid] : 0.f;
  __syncthreads();
  amax_to(mx + MX_B + B_DZ1, d_r1, H0, false);
}

// Scales (every thread recomputes them from the max buffer: a few dozen flops).  sw: weight scales (WScale), sa: forward
// activation scales, sb: backward operand scales.  The row-normalised loss gradient has max |d_y0| in [1, 2).
struct Scales {
  float sa[N_ACT], sw[N_WSCALE], sb[N_BACT];
};
__device__ __forceinline__ Scales make_scales(const uint32_t* __restrict__ mx, float loss_scale) {
  Scales z;
  for (int a = 0; a < N_ACT; ++a) z.sa[a] = pow2_scale(__uint_as_float(mx[a]), ACT_TARGET);
  for (int l = 0; l < 10; ++l) z.sw[l] = pow2_scale(__uint_as_float(mx[N_ACT + l]), W_TARGET);
  // Loss gradient dv = 2 s w scale (nabla_V - target): bounded by 2 |scale| max|w| (exact) x max|diff| (sampled, x8 margin).
  // One GLOBAL power-of-two scale maps that bound to 4096 (16x of fp16 head room left): the hi / lo pairs of every
  // gradient tensor then serve K3a's own dgrad MMAs and, unchanged, K3b's contraction over the points.  Rows far below
  // the bound lose relative precision (absolute error 2^-25 / scale), in proportion to how little they add to the sum.
  float wmax = __uint_as_float(mx[MX_W]), dmax = __uint_as_float(mx[MX_DIFF]);
  wmax = wmax > 0.f ? wmax : 1.f;
  dmax = dmax > 0.f ? dmax : 1.f;
  const float bound = 2.f * fabsf(loss_scale) * wmax * dmax * 8.f;
  z.sb[B_DY0] = pow2_scale(bound, 4096.f);
  for (int a = B_DY1; a < N_BACT; ++a) {
    const float g = __uint_as_float(mx[MX_B + a]);
    z.sb[a] = pow2_scale(bound * (g > 0.f ? g : 1.f), 4096.f);
  }
  // products that accumulate onto another product's accumulator arrive in its units; if that pushes the weight block
  // out of the comfortable fp16 range, move the operand scale instead (it has 2^10 of head room and a 2^15 window)
  auto matched = [&](float& s_op, float target_total, float wmax) {
    float sw = target_total / s_op;
    while (sw * wmax > 16384.f) { sw *= 0.5f; s_op *= 2.f; }
    while (sw * wmax < 16.f && wmax > 0.f) { sw *= 2.f; s_op *= 0.5f; }
    return sw;
  };
  z.sw[WS_D2T] = matched(z.sb[B_DZ3], z.sb[B_DO2] * z.sw[5], __uint_as_float(mx[N_ACT + 2]));
  z.sw[WS_WCT] = matched(z.sb[B_DY0], z.sb[B_DZ2] * z.sw[1], __uint_as_float(mx[N_ACT + WC_LAYER]));
  return z;
}

__global__ void pack_h_kernel(socm_unet net, const float* __restrict__ wc, const uint32_t* __restrict__ mx,
                              unsigned char* __restrict__ tape, float* __restrict__ small, int with_bwd, float loss_scale) {
  const int d = net.d;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const Small so = small_layout();
  const Scales z = make_scales(mx, loss_scale);
  const int n_items = FWD_ITEMS + (with_bwd ? BWD_ITEMS : 0);
  for (int it = 0; it < n_items; ++it) {
    const PackItem pi = it < FWD_ITEMS ? fwd_item(d, it) : bwd_item(d, it - FWD_ITEMS);
    const SlotDesc sd = pi.sd;
    const float* W = sd.layer == WC_LAYER ? wc : net.w[sd.layer];
    const float s = z.sw[sd.sid >= 0 ? sd.sid : sd.layer];
    unsigned char* base = tape + (size_t)(pi.slot + (it < FWD_ITEMS ? 0 : FWD_SLOTS)) * SLOT_BYTES + pi.byte_off;
    const int slab = (sd.slab_n ? sd.slab_n : sd.N) * sd.Kc * 2;
    for (int i = tid; i < sd.N * sd.Kc; i += nth) {
      const int n = i / sd.Kc, k = i - n * sd.Kc;
      float w = 0.f;
      if (sd.k0 + k < sd.klim && sd.n0 + n < sd.nlim)
        w = sd.transposed ? W[(size_t)(sd.k0 + k) * sd.ktot + sd.n0 + n] : W[(size_t)(sd.n0 + n) * sd.ktot + sd.k0 + k];
      w *= s;
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      const int off = wslab_off(sd.slab_row + n, k, sd.Kc);
      *reinterpret_cast<__half*>(base + off) = hi;
      *reinterpret_cast<__half*>(base + slab + off) = lo;
    }
  }
  auto copy = [&](int off, const float* src, int n, float scale) {
    for (int i = tid; i < n; i += nth) small[off + i] = src[i] * scale;
  };
  copy(so.b_d0, net.b[0], H0, z.sa[A_R1]);
  copy(so.b_d1, net.b[1], H1, z.sa[A_R2]);
  copy(so.b_d2, net.b[2], H2, z.sa[A_R3]);
  copy(so.b_u2, net.b[6], H1, z.sa[A_O2]);
  copy(so.b_r2, net.b[5], H1, z.sa[A_O2]);
  copy(so.b_u1, net.b[7], H0, z.sa[A_Y1]);
  for (int i = tid; i < KIN; i += nth) small[so.b_r0 + i] = i < d ? net.b[3][i] : 0.f;
  for (int i = tid; i < KIN * KIN; i += nth) {
    const int j = i / KIN, k = i - j * KIN;
    small[so.r0 + i] = (j < d && k <= d) ? net.w[3][(size_t)j * (d + 1) + k] : 0.f;
  }
  if (tid == 0) {
    for (int a = 0; a < N_ACT; ++a) small[so.sa + a] = z.sa[a];
    for (int l = 0; l < 10; ++l) {
      const float inv = 1.f / (z.sa[act_of_layer(l)] * z.sw[l]);
      small[so.inv + l] = inv;
      small[so.invs + l] = act_out_of_layer(l) >= 0 ? inv * z.sa[act_out_of_layer(l)] : inv;
    }
    for (int a = 0; a < N_BACT; ++a) small[so.sb + a] = z.sb[a];
    // accumulator units of the five backward products: s_in * w
    const float unit[N_BPROD] = {z.sb[B_DY0] * z.sw[8], z.sb[B_DY1] * z.sw[7], z.sb[B_DO2] * z.sw[6], z.sb[B_DO2] * z.sw[5],
                                 z.sb[B_DZ2] * z.sw[1]};
    const float s_out[N_BPROD] = {z.sb[B_DY1], z.sb[B_DO2], z.sb[B_DZ3], z.sb[B_DZ2], z.sb[B_DZ1]};
    for (int q = 0; q < N_BPROD; ++q) {
      small[so.bt + q] = 1.f / unit[q];
      small[so.bf + q] = s_out[q] / unit[q];
    }
    const float sk[N_SCRATCH_T] = {z.sa[A_X], z.sa[A_R1], z.sa[A_R2], z.sa[A_R3], z.sa[A_O2], z.sa[A_Y1], z.sb[B_DY0], z.sb[B_DY0],
                                   z.sb[B_DY1], z.sb[B_DO2], z.sb[B_DO2], z.sb[B_DZ3], z.sb[B_DZ2], z.sb[B_DZ1]};
    for (int t = 0; t < N_SCRATCH_T; ++t) small[so.sk + t] = sk[t];
  }
}

int setup_h(const socm_unet* net, unsigned char* ws, const CalibArgs& c, bool with_bwd, cudaStream_t stream) {
  float* small = small_ptr(ws);
  float* wc = wc_ptr(ws);
  uint32_t* mx = max_ptr(ws);
  SOCM_CUDA(cudaMemsetAsync(mx, 0, 64 * sizeof(uint32_t), stream));
  fold_h_kernel<<<NY, H0, 0, stream>>>(*net, wc, small);
  SOCM_LAUNCH_CHECK();
  const size_t smem = (size_t)(generic::fwd_floats(net->d, H0, H1, H2) + CALIB_BWD_FLOATS) * sizeof(float);
  calib_h_kernel<<<11 + c.n_samples, 256, smem, stream>>>(*net, wc, c, mx);
  SOCM_LAUNCH_CHECK();
  pack_h_kernel<<<96, 256, 0, stream>>>(*net, wc, mx, ws, small, with_bwd ? 1 : 0, c.loss_scale);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

// ---------------------------------------------------------------- shared memory / barriers
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * SLOT_BYTES;
constexpr int SM_XIN = SM_CHUNK + 2 * CHUNK_BYTES;
constexpr int SM_SMALL = SM_XIN + 2 * XIN_HALF;
enum Bar {
  W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2, PC_FULL = CH_EMPTY + 2,
  PC_EMPTY = PC_FULL + 2, XIN_FULL = PC_EMPTY + 2, D1_FULL, R2H_FULL, R2_FULL, D2_FULL, R3_FULL, R2A_DONE, D3_FULL,
  O2_FULL, Y0_FULL, N_BARS
};
__host__ __device__ inline int rollout_h_smem_bytes() { return SM_SMALL + small_layout().total * 4 + N_BARS * 8 + 16; }
constexpr int NT = 320, NE = 256;
// TMEM columns
constexpr uint32_t C_PC = 0, C_D2 = 0, C_R2A = 0, C_SA = 64, C_WC = 192, C_R2B = 192, C_Y0 = 192;

#ifdef SOCM_H_PROF
// per-phase cycle accumulators of block 0 (owner thread 0: slots 0-15, helper thread 128: 16-31, M warp: 32-47)
__device__ unsigned long long g_h_prof[64];
#define HP_DECL long long prof_t = clock64(); unsigned long long prof_acc[16] = {0}
#define HP_MARK(i) do { const long long t_ = clock64(); prof_acc[i] += (unsigned long long)(t_ - prof_t); prof_t = t_; } while (0)
#define HP_FLUSH(base, cond) do { if (blockIdx.x == 0 && (cond)) for (int i_ = 0; i_ < 16; ++i_) g_h_prof[(base) + i_] = prof_acc[i_]; } while (0)
#define HP2_DECL long long prof2_t = clock64(); unsigned long long prof2_acc[8] = {0}
#define HP2_START prof2_t = clock64()
#define HP2_MARK(i) do { const long long t_ = clock64(); prof2_acc[i] += (unsigned long long)(t_ - prof2_t); prof2_t = t_; } while (0)
#define HP2_FLUSH(cond) do { if (blockIdx.x == 0 && (cond)) for (int i_ = 0; i_ < 8; ++i_) g_h_prof[48 + i_] = prof2_acc[i_]; } while (0)
#else
#define HP2_DECL
#define HP2_START
#define HP2_MARK(i)
#define HP2_FLUSH(cond)
#define HP_DECL
#define HP_MARK(i)
#define HP_FLUSH(base, cond)
#endif

template <int KINP>
__device__ __forceinline__ void draw_noise_reg(const RolloutArgs& a, int m, int k, float* eps) {
  const int d = a.st.d;
  if (a.noise_in != nullptr) {
    const float* src = a.noise_in + ((size_t)k * a.B + m) * d;
#pragma unroll
    for (int j = 0; j < KINP; ++j)
      if (j < d) eps[j] = __ldg(src + j);
  } else {
#pragma unroll
    for (int blk = 0; blk < KINP / 4; ++blk) {
      if (blk * 4 < d) {
        float z[4];
        philox_normal4(a.seed, a.path_offset + (uint64_t)m, (uint32_t)k, (uint32_t)blk, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) eps[blk * 4 + j] = (blk * 4 + j < d) ? z[j] : 0.f;
      }
    }
  }
}

__global__ void __launch_bounds__(NT, 2) rollout_h_kernel(RolloutArgs a, const unsigned char* __restrict__ tape,
                                                          const float* __restrict__ small_g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  const int d = a.st.d, K = a.K, B = a.B;
  const Small so = small_layout();
  float* sm_small = reinterpret_cast<float*>(smem + SM_SMALL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_SMALL + so.total * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool diag = diag_fast_path(a.st, a.warmA != nullptr);

  for (int i = tid; i < so.total; i += NT) sm_small[i] = __ldg(small_g + i);
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bars[W_FULL + s], 1);
      mbar_init(&bars[W_EMPTY + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[CH_FULL + b], NE / 32);
      mbar_init(&bars[CH_EMPTY + b], 1);
      mbar_init(&bars[PC_FULL + b], 1);
      mbar_init(&bars[PC_EMPTY + b], NE / 32);
    }
    mbar_init(&bars[XIN_FULL], TP / 32);
    const int e2m[] = {R2H_FULL, R2_FULL, R3_FULL, O2_FULL};
    for (int i = 0; i < 4; ++i) mbar_init(&bars[e2m[i]], NE / 32);
    const int m2e[] = {D1_FULL, D2_FULL, R2A_DONE, D3_FULL, Y0_FULL};
    for (int i = 0; i < 5; ++i) mbar_init(&bars[m2e[i]], 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK), xin_s = smem_addr(smem + SM_XIN);

  if (warp < 8) {
    // =================================================================== E: epilogue / path threads
    auto e_program = [&](auto h_const) {
      constexpr int h = decltype(h_const)::value;
      const int p = tid & (TP - 1);
      const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
      uint32_t g = 0;    // steps done (all tiles)
      uint32_t cu = 0;   // shared-memory chunks produced
      uint32_t pu = 0;   // TMEM pieces consumed
      HP_DECL;
      HP2_DECL;
      float* stage_f = reinterpret_cast<float*>(smem + SM_CHUNK);  // exchange area = chunk buffer 0 (idle after up_0)
      const float s_x = sm_small[so.sa + A_X];

      // 256-wide accumulator arriving in 8 TMEM pieces of 32 columns -> 8 shared-memory A chunks; this thread handles
      // features [16 h, 16 h + 16) of every piece.  v = relu(acc * inv + bias) * s_out
      auto pieces_to_chunks = [&](int bias_off, float inv) {
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int b = pu & 1;
          HP2_START;
          mbar_wait_parked(&bars[PC_FULL + b], (pu >> 1) & 1);
          HP2_MARK(0);
          fence_after_sync();
          float v[16];
          tmem_ld16(lane_t + C_PC + 32 * b + 16 * h, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          HP2_MARK(1);
          fence_before_sync();
          warp_arrive(&bars[PC_EMPTY + b]);
          HP2_MARK(2);
          ++pu;
          const float* bias = sm_small + bias_off + 32 * c + 16 * h;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
          uint32_t hi[8], lo[8];
          split16(v, hi, lo);
          HP2_MARK(3);
          const int cb = cu & 1;
          mbar_wait_parked(&bars[CH_EMPTY + cb], ((cu >> 1) & 1) ^ 1);
          HP2_MARK(4);
          store_chunk16(smem + SM_CHUNK + cb * CHUNK_BYTES, p, 2 * h, hi, lo);
          fence_async_smem();
          HP2_MARK(5);
          warp_arrive(&bars[CH_FULL + cb]);
          HP2_MARK(6);
          ++cu;
        }
      };

      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int m = t * TP + p;
        const bool live = m < B;
        float x[KIN];   // owners (h == 0): state, x[j] for j < d
        PathAcc acc{1.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < KIN; ++j) x[j] = (h == 0 && j < d && live) ? __ldg(a.x0 + (size_t)m * d + j) : 0.f;
        if (live && h == 0) {
          if (a.states) {
#pragma unroll
            for (int j = 0; j < KIN; ++j)
              if (j < d) a.states[(size_t)m * d + j] = x[j];
          }
          if (a.stop) a.stop[m] = 1.f;
        }
        for (int k = 0; k < K; ++k, ++g) {
          const uint32_t ph = g & 1;
          // ---- E0 (owners): input operand [t, x, 0..] * s_x as fp16 hi / lo (two core-matrix columns of 8 features)
          if (h == 0) {
            const float tk = __ldg(a.step_tab + 4 * K + k);
            float xb[KIN];
            xb[0] = tk * s_x;
#pragma unroll
            for (int c = 1; c < KIN; ++c) xb[c] = x[c - 1] * s_x;   // x[j >= d] = 0
            uint32_t hi[8], lo[8];
            split16(xb, hi, lo);
            unsigned char* base = smem + SM_XIN + (p % 8) * 16 + (p / 8) * 128;
            *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(base + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
            *reinterpret_cast<uint4*>(base + XIN_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(base + XIN_HALF + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            fence_async_smem();
            warp_arrive(&bars[XIN_FULL]);
          }
          HP_MARK(0);
          // ---- E1: r1 chunks for down_1
          pieces_to_chunks(so.b_d0, sm_small[so.invs + 0]);
          HP_MARK(1);
          // ---- E2: r2 = relu(D1 + b) -> A operand in place; helpers keep the Wc r1 accumulator
          mbar_wait_parked(&bars[D1_FULL], ph);
          HP_MARK(2);
          fence_after_sync();
          float au[NY];
          if (h == 1) {
            tmem_ld16(lane_t + C_WC, reinterpret_cast<uint32_t*>(au));
            tmem_wait_ld();
          }
          {
            const float inv = sm_small[so.invs + 1];
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int gq = 2 * i + h;   // 16-feature group; iterations 0, 1 cover K half 0 (groups 0..3) of down_2
              float v[16];
              tmem_ld16(lane_t + C_SA + 16 * gq, reinterpret_cast<uint32_t*>(v));
              tmem_wait_ld();
              const float* bias = sm_small + so.b_d1 + 16 * gq;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
              uint32_t hi[8], lo[8];
              split16(v, hi, lo);
              store_group16(lane_t + C_SA + 16 * gq, hi, lo);
              if (i == 1) {
                tmem_wait_st();
                fence_before_sync();
                warp_arrive(&bars[R2H_FULL]);
              }
            }
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[R2_FULL]);
          }
          HP_MARK(3);
          // ---- E3: r3 = relu(D2 + b) -> shared-memory A operand: features [32 h, 32 h + 32) = chunk buffer h
          mbar_wait_parked(&bars[D2_FULL], ph);
          HP_MARK(4);
          fence_after_sync();
          {
            const float inv = sm_small[so.invs + 2];
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
              float v[16];
              tmem_ld16(lane_t + C_D2 + 32 * h + 16 * i, reinterpret_cast<uint32_t*>(v));
              tmem_wait_ld();
              const float* bias = sm_small + so.b_d2 + 32 * h + 16 * i;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
              uint32_t hi[8], lo[8];
              split16(v, hi, lo);
              store_chunk16(smem + SM_CHUNK + h * CHUNK_BYTES, p, 2 * i, hi, lo);
            }
          }
          fence_before_sync();
          fence_async_smem();
          warp_arrive(&bars[R3_FULL]);
          HP_MARK(5);
          // ---- E5: o2 = relu(D3 + b_u2) + D3R + b_r2 -> A operand in place over D3
          mbar_wait_parked(&bars[D3_FULL], ph);
          HP_MARK(6);
          fence_after_sync();
          {
            const float inv_u2 = sm_small[so.invs + 6], inv_r2 = sm_small[so.invs + 5];
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int gq = 2 * i + h;
              float y[16], rr[16];
              tmem_ld16(lane_t + C_SA + 16 * gq, reinterpret_cast<uint32_t*>(y));
              tmem_ld16(lane_t + (gq < 4 ? C_R2A + 16 * gq : C_R2B + 16 * (gq - 4)), reinterpret_cast<uint32_t*>(rr));
              tmem_wait_ld();
              const float* bu = sm_small + so.b_u2 + 16 * gq;
              const float* br = sm_small + so.b_r2 + 16 * gq;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                y[j] = fmaxf(fmaf(y[j], inv_u2, bu[j]), 0.f) + fmaf(rr[j], inv_r2, br[j]);
              uint32_t hi[8], lo[8];
              split16(y, hi, lo);
              store_group16(lane_t + C_SA + 16 * gq, hi, lo);
            }
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[O2_FULL]);
          }
          HP_MARK(7);
          // up_1 runs now: the helper half draws this step's noise meanwhile
          float eps[KIN];
#pragma unroll
          for (int j = 0; j < KIN; ++j) eps[j] = 0.f;
          if (h == 1 && live) draw_noise_reg<KIN>(a, m, k, eps);
          HP_MARK(8);
          // ---- E6: y1 = relu(D4 + b_u1) -> A chunks of the folded up_0
          pieces_to_chunks(so.b_u1, sm_small[so.invs + 7]);
          HP_MARK(9);
          // ---- E8: y0 = W_u0 y1 (TMEM) + Wc r1 (helper registers) + bc; helpers hand relu(y0) and the noise to the owners
          mbar_wait_parked(&bars[Y0_FULL], ph);
          HP_MARK(10);
          fence_after_sync();
          if (h == 1) {
            float yp[NY];
            tmem_ld16(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
            tmem_wait_ld();
            const float inv_u0 = sm_small[so.inv + 8], inv_wc = sm_small[so.inv + WC_LAYER];
#pragma unroll
            for (int j = 0; j < KIN; ++j) {
              const float y0 = fmaf(yp[j], inv_u0, fmaf(au[j], inv_wc, sm_small[so.bc + j]));
              stage_f[j * TP + p] = fmaxf(y0, 0.f);
              stage_f[(KIN + j) * TP + p] = eps[j];
            }
          }
          fence_before_sync();  // orders the TMEM reads before the next step's MMAs (via XIN_FULL)
          HP_MARK(11);
          e_sync();
          HP_MARK(12);
          if (h == 0 && live) {
            float gv[KIN], u[KIN];
            const float tk = __ldg(a.step_tab + 4 * K + k);
#pragma unroll
            for (int j = 0; j < KIN; ++j) {  // res_0 [t, x]; lanes >= d come out as exact zeros (zero-padded parameters)
              const float4* wr = reinterpret_cast<const float4*>(sm_small + so.r0 + j * KIN);
              float acc_r = sm_small[so.b_r0 + j];
#pragma unroll
              for (int c4 = 0; c4 < KIN / 4; ++c4) {
                const float4 w = wr[c4];
                acc_r = fmaf(w.x, c4 == 0 ? tk : x[4 * c4 - 1], acc_r);
                acc_r = fmaf(w.y, x[4 * c4], acc_r);
                acc_r = fmaf(w.z, x[4 * c4 + 1], acc_r);
                acc_r = fmaf(w.w, x[4 * c4 + 2], acc_r);
              }
              gv[j] = stage_f[j * TP + p] + acc_r;
              eps[j] = stage_f[(KIN + j) * TP + p];
            }
            const float dt = __ldg(a.step_tab + k), sq_ldt = __ldg(a.step_tab + K + k);
            const float dt_l = __ldg(a.step_tab + 2 * K + k), sq_dtl = __ldg(a.step_tab + 3 * K + k);
            float eff;
            if (diag) {
              float kap[KIN];
#pragma unroll
              for (int j = 0; j < KIN; ++j) kap[j] = j < d ? __ldg(a.st.kappa + j) : 0.f;
              eff = sde_step_diag<KIN>(a.st.kind == SOCM_MOLECULAR_DYNAMICS, a.st.lmbd, kap, x, gv, eps, u, dt, sq_ldt,
                                       dt_l, sq_dtl, acc);
            } else {
              const float* wA = a.warmA ? a.warmA + (size_t)k * d * d : nullptr;
              const float* wc = a.warmc ? a.warmc + (size_t)k * d : nullptr;
              float xl[kMaxDim], gl[kMaxDim], el[kMaxDim], ul[kMaxDim];
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) {
                  xl[j] = x[j];
                  gl[j] = gv[j];
                  el[j] = eps[j];
                }
              eff = sde_step(a.st, wA, wc, xl, 1, gl, 1, el, ul, dt, sq_ldt, dt_l, sq_dtl, acc);
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) {
                  x[j] = xl[j];
                  u[j] = ul[j];
                }
            }
            const size_t row = (size_t)k * B + m;
            if (a.states) {
              float* sp = a.states + (row + B) * d;
              float* cp = a.controls + row * d;
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) {
                  sp[j] = x[j];
                  cp[j] = u[j];
                }
              if (a.noise_in == nullptr) {
                float* np_ = a.noises + row * d;
#pragma unroll
                for (int j = 0; j < KIN; ++j)
                  if (j < d) np_[j] = eps[j];
              }
              a.stop[row + B] = acc.alive;
              a.eff_dt[row] = eff;
            }
          }
          HP_MARK(13);
          e_sync();  // the exchange area is chunk buffer 0 again from here on
          HP_MARK(14);
        }
        if (live && h == 0) {
          a.lw_det[m] = acc.lw_det;
          a.lw_sto[m] = acc.lw_sto;
          float xl[kMaxDim];
#pragma unroll
          for (int j = 0; j < KIN; ++j)
            if (j < d) xl[j] = x[j];
          a.lw_term[m] = __fdiv_rn(-term_cost(a.st, xl, 1), a.st.lmbd);  // utils.py:101
        }
      }
      HP_FLUSH(0, tid == 0);
      HP2_FLUSH(tid == 0);
      HP_FLUSH(16, tid == 128);
    };  // e_program
    if (warp < 4) e_program(std::integral_constant<int, 0>{});
    else e_program(std::integral_constant<int, 1>{});
  } else if (warp == 8) {
    // =================================================================== M: MMA issue
    uint32_t ws = 0;  // weight stages consumed
    uint32_t cm = 0;  // chunks consumed
    uint32_t pm = 0;  // pieces produced
    HP_DECL;
    const uint32_t total_steps = (uint32_t)my_tiles * (uint32_t)K;
    auto wait_w = [&]() -> uint32_t {
      const uint32_t s = ws % NSTAGE;
      mbar_wait_parked(&bars[W_FULL + s], (ws / NSTAGE) & 1);
      fence_after_sync();
      return ring_s + s * SLOT_BYTES;
    };
    auto release_w = [&]() {
      if (elect_one()) commit(&bars[W_EMPTY + ws % NSTAGE]);
      __syncwarp();
      ++ws;
    };
    auto signal = [&](int bar) {
      if (elect_one()) commit(&bars[bar]);
      __syncwarp();
    };
    auto wait_e = [&](int bar, uint32_t ph) {
      mbar_wait_parked(&bars[bar], ph);
      fence_after_sync();
    };
    // piece q of down_0: D[128 x 32] = xin (shared memory, K = 16) x W0 rows [32 q, 32 q + 32)  (block at b_smem)
    auto d0_piece = [&](uint32_t b_smem) {
      const uint32_t b = pm & 1;
      mbar_wait_parked(&bars[PC_EMPTY + b], ((pm >> 1) & 1) ^ 1);
      fence_after_sync();
      if (elect_one()) {
        issue_ss<32, 1>(tm + C_PC + 32 * b, xin_s, XIN_HALF, b_smem, KIN, 32, true);
        commit(&bars[PC_FULL + b]);
      }
      __syncwarp();
      ++pm;
    };
    for (uint32_t g = 0; g < total_steps; ++g) {
      const uint32_t ph = g & 1;
      // ---- down_0 pieces 0, 1
      HP_MARK(0);
      wait_e(XIN_FULL, ph);
      HP_MARK(1);
      {
        const uint32_t wb = wait_w();
        d0_piece(wb);
        d0_piece(wb + PIECE_BYTES);
        release_w();
      }
      // ---- down_1 (+ Wc r1) on the r1 chunks; piece c + 2 of down_0 rides in the same slot
      for (int c = 0; c < 8; ++c) {
        const uint32_t wb = wait_w();
        // piece c + 2 only needs the epilogue to have READ piece c (same buffer), not chunk c to be finished: issued
        // first, it is ready long before the epilogue threads come back for it
        if (c < 6) d0_piece(wb + D1_MAIN);
        const uint32_t b = cm & 1;
        mbar_wait_parked(&bars[CH_FULL + b], (cm >> 1) & 1);
        fence_after_sync();
        if (elect_one()) {
          issue_ss<H1 + NY, 2>(tm + C_SA, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, 32, H1 + NY, c == 0);
          commit(&bars[CH_EMPTY + b]);
        }
        __syncwarp();
        ++cm;
        release_w();
      }
      signal(D1_FULL);
      HP_MARK(2);
      // ---- down_2: A = r2 (TMEM, in place at C_SA), K halves as they become ready
      wait_e(R2H_FULL, ph);
      HP_MARK(3);
      for (int j = 0; j < 2; ++j) {
        if (j == 1) wait_e(R2_FULL, ph);
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<H2, 4>(tm + C_D2, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D2_FULL);
      // ---- res_2, output half b (features 64..127) -> [192,256): runs under the r3 epilogue
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_R2B, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      // ---- res_2, output half a -> [0,64) once the epilogue has read D2 from there
      HP_MARK(4);
      wait_e(R3_FULL, ph);
      HP_MARK(5);
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_R2A, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      // up_2 writes its accumulator over r2, which res_2 is still reading: wait until those MMAs have completed
      HP_MARK(6);
      signal(R2A_DONE);
      wait_e(R2A_DONE, ph);
      HP_MARK(7);
      // ---- up_2: A = r3 in the two shared-memory chunk buffers -> D3 at C_SA
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ss<H1, 2>(tm + C_SA, chunk_s + j * CHUNK_BYTES, CHUNK_HALF, wb, 32, H1, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D3_FULL);
      HP_MARK(8);
      // ---- up_1 in 8 pieces of 32 output features (A = o2 in place at C_SA); the folded up_0 follows one piece behind
      wait_e(O2_FULL, ph);
      HP_MARK(9);
      for (int q = 0; q < 8; ++q) {
        const uint32_t wb = wait_w();
        const uint32_t b = pm & 1;
        mbar_wait_parked(&bars[PC_EMPTY + b], ((pm >> 1) & 1) ^ 1);
        fence_after_sync();
        if (elect_one()) {
          issue_ts<32, 8>(tm + C_PC + 32 * b, tm + C_SA, wb, H1, true);
          commit(&bars[PC_FULL + b]);
        }
        __syncwarp();
        ++pm;
        if (q >= 2) {   // folded up_0 on y1 chunk q - 2 (two pieces behind: that chunk is finished or nearly so)
          const uint32_t cb = cm & 1;
          mbar_wait_parked(&bars[CH_FULL + cb], (cm >> 1) & 1);
          fence_after_sync();
          if (elect_one()) {
            issue_ss<NY, 2>(tm + C_Y0, chunk_s + cb * CHUNK_BYTES, CHUNK_HALF, wb + U1_MAIN, 32, NY, q == 2);
            commit(&bars[CH_EMPTY + cb]);
          }
          __syncwarp();
          ++cm;
        }
        release_w();
      }
      {
        const uint32_t wb = wait_w();
        for (int j = 0; j < 2; ++j) {
          const uint32_t cb = cm & 1;
          mbar_wait_parked(&bars[CH_FULL + cb], (cm >> 1) & 1);
          fence_after_sync();
          if (elect_one()) {
            issue_ss<NY, 2>(tm + C_Y0, chunk_s + cb * CHUNK_BYTES, CHUNK_HALF, wb + j * PIECE_BYTES, 32, NY, false);
            commit(&bars[CH_EMPTY + cb]);
          }
          __syncwarp();
          ++cm;
        }
        release_w();
      }
      signal(Y0_FULL);
      HP_MARK(10);
    }
    HP_FLUSH(32, (tid & 31) == 0);
  } else {
    // =================================================================== P: weight tape producer
    if (elect_one()) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)K * FWD_SLOTS;
      uint32_t in_step = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % NSTAGE;
        mbar_wait_parked(&bars[W_EMPTY + s], ((i / NSTAGE) & 1) ^ 1);
        const uint32_t bytes = fwd_slot_bytes((int)in_step);
        mbar_expect_tx(&bars[W_FULL + s], bytes);
        bulk_g2s(smem + SM_RING + s * SLOT_BYTES, tape + (size_t)in_step * SLOT_BYTES, bytes, &bars[W_FULL + s]);
        if (++in_step == FWD_SLOTS) in_step = 0;
      }
    }
    __syncwarp();
  }

  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 256);
}

#ifdef SOCM_H_PROF
extern "C" int socm_debug_h_prof(unsigned long long* out48) {
  return (int)cudaMemcpyFromSymbol(out48, g_h_prof, sizeof(g_h_prof));
}
#endif

bool rollout_h_supported(const socm_unet* net) { return is_default_arch(net) && net->d <= MAX_D; }
int64_t rollout_h_workspace_bytes() { return workspace_bytes(); }

int launch_rollout_h(const RolloutArgs& a, const socm_unet* net, void* workspace, cudaStream_t stream) {
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  CalibArgs c{};
  c.x0 = a.x0;
  c.B = a.B;
  c.K = a.K;
  c.n_samples = 256;
  c.step_tab = a.step_tab;
  c.lmbd = a.st.lmbd;
  if (int rc = setup_h(net, ws, c, false, stream)) return rc;
  const int smem = rollout_h_smem_bytes();
  const int n_tiles = (a.B + TP - 1) / TP;
  const int max_ctas = 2 * sm_count();
  const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
  SOCM_CUDA(cudaFuncSetAttribute(rollout_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rollout_h_kernel<<<grid, NT, smem, stream>>>(a, ws, small_ptr(ws));
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace hx
}  // namespace socm
