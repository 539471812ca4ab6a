// loss_h.cu -- K3a on the fp16-split tcgen05 engine (unet_h.cuh), two CTAs per SM: UNet forward at 128 trajectory points,
// weighted loss and the whole dgrad chain.  Replaces method.py:272-287 (nabla_V at all points), 692-720 (loss) and the
// activation-gradient half of loss.backward() (main.py:323); the weight gradients are K3b (wgrad_h.cu), which reads the
// operands this kernel leaves in the scratch: tile -> 4 quarters -> 59 feature blocks of 4 KB as in loss_tc.cuh, but a block
// holds fp16 hi and lo planes of the operand-scaled values in the no-swizzle MN-major core-matrix layout (store_fb16 below)
// -- the hi / lo pairs this kernel computes for its own MMAs anyway, so the scratch costs one 16-byte store per 8 values at
// a compile-time offset, and K3b needs no split pass.
//
// A tile is 128 consecutive paths at one grid time t_i.  Forward = the rollout's forward (rollout_h.cu) from the stored
// state, ReLU masks kept in registers.  Backward = its mirror image on the transposed weight tape.  The loss gradient
// d loss / d nabla_V carries ONE global power-of-two scale (from the exact max |w| and the sampled max |nabla_V - target|,
// rollout_h.cu: make_scales), every later gradient tensor a per-layer one from the calibration pass.
//   d_o1 = W_u0^T d_y0 (8 pieces)  ->  d_y1 = m_y1 . d_o1 (chunks)  ->  d_o2 = W_u1^T d_y1
//   d_r3 = W_u2^T (m_y2 . d_o2),  d_z3 = m_r3 . d_r3
//   d_r2 = W_r2^T d_o2 + W_d2^T d_z3,  d_z2 = m_r2 . d_r2
//   d_r1 = W_d1^T d_z2 + Wc^T d_y0 (8 pieces),  d_z1 = m_r1 . d_r1            (res_1 folded into up_0, unet_tc.cuh)
// TMEM columns (backward): [0,64) piece buffers of d_o1 -> d_r3 acc -> d_r2 acc a -> d_z2 operand a;  [64,192) d_o2 acc ->
// d_o2 operand (in place) -> piece buffers of d_r1;  [192,256) d_r2 acc b -> d_z2 operand b.
#include <cuda_fp16.h>

#include <type_traits>

#include <cstdlib>

#include "kernels.h"
#include "loss_common.cuh"
#include "loss_tc.cuh"
#include "unet_h.cuh"

namespace socm {
namespace tc {
__global__ void fold_finish_kernel(socm_unet net, const float* __restrict__ aux, float* __restrict__ grad);
}
namespace hx {

using namespace umma;
using tc::FB_BYTES;
using tc::QUARTER_BYTES;
using tc::TILE_BYTES;

namespace k3 {
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * SLOT_BYTES;
constexpr int SM_XIN = SM_CHUNK + 2 * CHUNK_BYTES;   // [t,x] operand (forward) / d_y0 operand (backward)
constexpr int SM_SMALL = SM_XIN + 2 * XIN_HALF;
enum Bar {
  W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2, PC_FULL = CH_EMPTY + 2,
  PC_EMPTY = PC_FULL + 2,
  // forward, one completion per tile each
  XIN_FULL = PC_EMPTY + 2, D1_FULL, R2H_FULL, R2_FULL, D2_FULL, R3_FULL, R2A_DONE, D3_FULL, O2_FULL, Y0_FULL,
  // backward
  DY0_FULL, BDO2_FULL, BO2_FULL, BD3_FULL, BZ3_FULL, BR2_FULL, BZ2_FULL, N_BARS
};
constexpr int NT = 320, NE = 256;
constexpr uint32_t C_PC = 0, C_D2 = 0, C_R2A = 0, C_SA = 64, C_WC = 192, C_R2B = 192, C_Y0 = 192;
constexpr uint32_t C_DR3 = 0, C_DR2A = 0, C_DR2B = 192, C_BPC = 64;   // backward
}  // namespace k3

__host__ __device__ inline int loss_h_smem_bytes() { return k3::SM_SMALL + small_layout().total * 4 + k3::N_BARS * 8 + 16; }

// Scratch stores: thread <-> point r of the quarter.  A feature block (32 features x 32 points, fp16 hi and lo) is four
// groups of 8 features x 1 KB in the canonical no-swizzle MN-MAJOR layout of tcgen05 (K = points, MN = features): per group
// eight 128-byte core matrices of 8 points x (8 features = 16 bytes), point groups 0..3 = hi plane, 4..7 = lo plane:
//     byte(f, r, plane) = (f / 8) * 1024 + plane * 512 + r * 16 + (f % 8) * 2 .
// A thread owns one point and holds packed pairs of neighbouring features, so eight features of one plane are ONE 16-byte
// store at a compile-time offset from (block + r * 16), and a warp's store covers 512 contiguous bytes.  (A K-major
// scratch needs one 2-byte store per value: ~8x the LSU wavefronts, which bounded this kernel.)
__device__ __forceinline__ int row_off(int r) { return r * 16; }
__device__ __forceinline__ void st_q(unsigned char* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#if defined(SOCM_SCRATCH_ST) && SOCM_SCRATCH_ST == 1
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
#elif defined(SOCM_SCRATCH_ST) && SOCM_SCRATCH_ST == 2
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
#else
  __stcs(reinterpret_cast<uint4*>(p), make_uint4(a, b, c, d));
#endif
}
// features [f0, f0 + 16) of the feature block at `blk` (already offset by row_off) from the packed pairs
// hi[i] / lo[i] = features (f0 + 2 i, f0 + 2 i + 1); f0 a multiple of 16
__device__ __forceinline__ void store_fb16(unsigned char* blk, int f0, const uint32_t* hi, const uint32_t* lo) {
  unsigned char* g = blk + (f0 >> 3) * 1024;
  st_q(g, hi[0], hi[1], hi[2], hi[3]);
  st_q(g + 512, lo[0], lo[1], lo[2], lo[3]);
  st_q(g + 1024, hi[4], hi[5], hi[6], hi[7]);
  st_q(g + 1536, lo[4], lo[5], lo[6], lo[7]);
}
// features [KIN, 32) of a d-sized block: zeros, except the constant-1 feature (hi = 1.0) when `ones`
__device__ __forceinline__ void store_fb_tail(unsigned char* blk, bool ones) {
  static_assert(KIN == 16 && tc::ONES_FEATURE == 31, "tail layout");
  st_q(blk + 2048, 0u, 0u, 0u, 0u);
  st_q(blk + 2048 + 512, 0u, 0u, 0u, 0u);
  st_q(blk + 3072, 0u, 0u, 0u, ones ? 0x3C000000u : 0u);
  st_q(blk + 3072 + 512, 0u, 0u, 0u, 0u);
}
__device__ __forceinline__ uint32_t positive_bits16(const float* v) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
  return m;
}
__device__ __forceinline__ void apply_bits16(float* v, uint32_t m) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
}

__global__ void __launch_bounds__(k3::NT, 2)
    loss_h_kernel(LossArgs a, const unsigned char* __restrict__ tape, const float* __restrict__ small_g,
                  unsigned char* __restrict__ scratch, int tile0, int n_tiles_launch) {
  using namespace k3;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  const int d = a.st.d, B = a.B;
  const Small so = small_layout();
  float* sm_small = reinterpret_cast<float*>(smem + SM_SMALL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_SMALL + so.total * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_mblk = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles_launch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool simple_loss = a.st.sigma_is_identity && a.warmA == nullptr;

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
    mbar_init(&bars[DY0_FULL], TP / 32);
    const int e2m[] = {R2H_FULL, R2_FULL, R3_FULL, O2_FULL, BO2_FULL, BZ3_FULL, BZ2_FULL};
    for (int i = 0; i < 7; ++i) mbar_init(&bars[e2m[i]], NE / 32);
    const int m2e[] = {D1_FULL, D2_FULL, R2A_DONE, D3_FULL, Y0_FULL, BDO2_FULL, BD3_FULL, BR2_FULL};
    for (int i = 0; i < 8; ++i) mbar_init(&bars[m2e[i]], 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 256);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK), xin_s = smem_addr(smem + SM_XIN);

  if (warp < 8) {
    // =================================================================== E: epilogue / point threads
    auto e_program = [&](auto h_const) {
      constexpr int h = decltype(h_const)::value;   // column half; h == 0 threads own the point
      const int p = tid & (TP - 1), r = tid & 31, q = (tid >> 5) & 3;
      const uint32_t lane_t = tm + ((uint32_t)(q * 32) << 16);
      uint32_t g = 0;    // tiles done
      uint32_t cu = 0;   // shared-memory chunks produced (ring)
      uint32_t pu = 0;   // TMEM pieces consumed
      float* stage_f = reinterpret_cast<float*>(smem + SM_CHUNK);   // exchange area = chunk buffer 0 (idle after up_0)
      double loss_acc = 0.0;
      const float s_x = sm_small[so.sa + A_X];

      for (int lt = blockIdx.x; lt < n_tiles_launch; lt += gridDim.x, ++g) {
        const int t = tile0 + lt;
        const int ti = t / n_mblk, m0 = (t - ti * n_mblk) * TP;
        const int m = m0 + p;
        const bool live = m < B;
        const uint32_t ph = g & 1;
        unsigned char* sq = scratch + (size_t)lt * TILE_BYTES + (size_t)q * QUARTER_BYTES + row_off(r);   // this warp's quarter, this thread's point
        uint64_t m_r1a = 0, m_r1b = 0, m_y1a = 0, m_y1b = 0, m_r2 = 0, m_y2 = 0;
        uint32_t m_r3a = 0, m_r3b = 0;

        // 8 TMEM pieces of 32 columns; this thread handles features [16 h, 16 h + 16) of each: fn(c, v) post-processes the
        // 16 accumulator values, the fp16 hi / lo pairs go to feature block fb0 + c of the scratch and, if `to_chunk`,
        // to shared-memory chunk c
        auto pieces = [&](uint32_t pc_col, bool to_chunk, int fb0, auto&& fn) {
#pragma unroll 1
          for (int c = 0; c < 8; ++c) {
            const int b = pu & 1;
            mbar_wait_parked(&bars[PC_FULL + b], (pu >> 1) & 1);
            fence_after_sync();
            float v[16];
            tmem_ld16(lane_t + pc_col + 32 * b + 16 * h, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            fence_before_sync();
            warp_arrive(&bars[PC_EMPTY + b]);
            ++pu;
            fn(c, v);
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_fb16(sq + (fb0 + c) * FB_BYTES, 16 * h, hi, lo);
            if (to_chunk) {
              const int cb = cu & 1;
              mbar_wait_parked(&bars[CH_EMPTY + cb], ((cu >> 1) & 1) ^ 1);
              store_chunk16(smem + SM_CHUNK + cb * CHUNK_BYTES, p, 2 * h, hi, lo);
              fence_async_smem();
              warp_arrive(&bars[CH_FULL + cb]);
              ++cu;
            }
          }
        };

        // ---- F0 (owners): state -> input operand (fp16 hi / lo, scaled) and the XIN block of the scratch
        if (h == 0) {
          const float tk = __ldg(a.ts + ti);
          float xb[KIN];
          xb[0] = tk;
#pragma unroll
          for (int c = 1; c < KIN; ++c)
            xb[c] = (c - 1 < d && live) ? __ldg(a.states + ((size_t)ti * B + m) * d + (c - 1)) : 0.f;
#pragma unroll
          for (int c = 0; c < KIN; ++c) xb[c] *= s_x;
          uint32_t hi[8], lo[8];
          split16(xb, hi, lo);
          // XIN block of the scratch: features [t, x] s_x, zeros, and the constant 1 (exactly 1.0 in fp16) at ONES_FEATURE
          unsigned char* xblk = sq + tc::FB_XIN * FB_BYTES;
          store_fb16(xblk, 0, hi, lo);
          store_fb_tail(xblk, true);
          unsigned char* base = smem + SM_XIN + (p % 8) * 16 + (p / 8) * 128;
          *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(base + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(base + XIN_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(base + XIN_HALF + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          fence_async_smem();
          warp_arrive(&bars[XIN_FULL]);
        }
        // ---- F1: r1 chunks for down_1 (+ mask, + scratch)
        {
          const float inv = sm_small[so.invs + 0];
          pieces(C_PC, true, tc::FB_R1, [&](int c, float* v) {
            const float* bias = sm_small + so.b_d0 + 32 * c + 16 * h;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
            const uint64_t bits = (uint64_t)positive_bits16(v) << (16 * (c & 3));
            if (c < 4) m_r1a |= bits;
            else m_r1b |= bits;
          });
        }
        // ---- F2: r2 in place; helpers keep the Wc r1 accumulator
        mbar_wait_parked(&bars[D1_FULL], ph);
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
            const int gq = 2 * i + h;
            float v[16];
            tmem_ld16(lane_t + C_SA + 16 * gq, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            const float* bias = sm_small + so.b_d1 + 16 * gq;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
            m_r2 |= (uint64_t)positive_bits16(v) << (16 * i);
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_group16(lane_t + C_SA + 16 * gq, hi, lo);
            store_fb16(sq + (tc::FB_R2 + (gq >> 1)) * FB_BYTES, 16 * (gq & 1), hi, lo);
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
        // ---- F3: r3 -> shared-memory A operand (features [32 h, 32 h + 32) = chunk buffer h), mask, scratch
        mbar_wait_parked(&bars[D2_FULL], ph);
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
            if (i == 0) m_r3a = positive_bits16(v);
            else m_r3b = positive_bits16(v);
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_chunk16(smem + SM_CHUNK + h * CHUNK_BYTES, p, 2 * i, hi, lo);
            store_fb16(sq + (tc::FB_R3 + h) * FB_BYTES, 16 * i, hi, lo);
          }
        }
        fence_before_sync();
        fence_async_smem();
        warp_arrive(&bars[R3_FULL]);
        // ---- F5: o2 = relu(D3 + b_u2) + D3R + b_r2 in place over D3; mask of the relu part; scratch
        mbar_wait_parked(&bars[D3_FULL], ph);
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
            for (int j = 0; j < 16; ++j) y[j] = fmaxf(fmaf(y[j], inv_u2, bu[j]), 0.f);
            m_y2 |= (uint64_t)positive_bits16(y) << (16 * i);
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] += fmaf(rr[j], inv_r2, br[j]);
            uint32_t hi[8], lo[8];
            split16(y, hi, lo);
            store_group16(lane_t + C_SA + 16 * gq, hi, lo);
            store_fb16(sq + (tc::FB_O2 + (gq >> 1)) * FB_BYTES, 16 * (gq & 1), hi, lo);
          }
          tmem_wait_st();
          fence_before_sync();
          warp_arrive(&bars[O2_FULL]);
        }
        // ---- F6: y1 = relu(D4 + b_u1) -> A chunks of the folded up_0 (+ mask, + scratch)
        {
          const float inv = sm_small[so.invs + 7];
          pieces(C_PC, true, tc::FB_Y1, [&](int c, float* v) {
            const float* bias = sm_small + so.b_u1 + 32 * c + 16 * h;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], inv, bias[j]), 0.f);
            const uint64_t bits = (uint64_t)positive_bits16(v) << (16 * (c & 3));
            if (c < 4) m_y1a |= bits;
            else m_y1b |= bits;
          });
        }
        // ---- F8: y0 = W_u0 y1 (TMEM) + Wc r1 (helper registers) + bc -> owners, through the exchange area
        mbar_wait_parked(&bars[Y0_FULL], ph);
        fence_after_sync();
        if (h == 1) {
          float yp[NY];
          tmem_ld16(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
          tmem_wait_ld();
          const float inv_u0 = sm_small[so.inv + 8], inv_wc = sm_small[so.inv + WC_LAYER];
#pragma unroll
          for (int j = 0; j < KIN; ++j) stage_f[j * TP + p] = fmaf(yp[j], inv_u0, fmaf(au[j], inv_wc, sm_small[so.bc + j]));
        }
        fence_before_sync();
        e_sync();
        // ---- loss (owners): nabla_V, d loss / d nabla_V, G; row normalisation; d_y0 operand for the backward pass
        if (h == 0) {
          float x[KIN], gv[KIN], y0[KIN], dv[KIN];
          const float tk = __ldg(a.ts + ti);
#pragma unroll
          for (int j = 0; j < KIN; ++j) x[j] = (j < d && live) ? __ldg(a.states + ((size_t)ti * B + m) * d + j) : 0.f;
#pragma unroll
          for (int j = 0; j < KIN; ++j) {   // res_0 [t, x]; lanes >= d come out as exact zeros (zero-padded parameters)
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
            y0[j] = stage_f[j * TP + p];
            gv[j] = fmaxf(y0[j], 0.f) + acc_r;
            dv[j] = 0.f;
          }
          if (live) {
            if (simple_loss) {
              const float* trow = a.target + (size_t)m * a.ldt + (size_t)ti * d;
              const float s = a.stop ? __ldg(a.stop + (size_t)ti * B + m) : 1.f;
              const float coef = s * __ldg(a.w + m) * a.scale;
              float sqs = 0.f;
              float* grow = a.G + (size_t)m * a.ldt + (size_t)ti * d;
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) {
                  const float diff = gv[j] - __ldg(trow + j);
                  sqs = fmaf(diff, diff, sqs);
                  dv[j] = 2.f * coef * diff;
                  grow[j] = -dv[j];
                }
              loss_acc += (double)(coef * sqs);
            } else {
              float xl[kMaxDim], gl[kMaxDim], dl[kMaxDim];
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) {
                  xl[j] = x[j];
                  gl[j] = gv[j];
                }
              loss_acc += (double)point_loss(a, ti, m, xl, 1, gl, 1, dl);
#pragma unroll
              for (int j = 0; j < KIN; ++j)
                if (j < d) dv[j] = dl[j];
            }
          }
          // d_y0 = 1[y0 > 0] dv and d_o0 = dv, both x the global gradient scale: operand of the backward pass + scratch
          const float s_g = sm_small[so.sb + B_DY0];
          float dy[KIN], dz[KIN];
#pragma unroll
          for (int c = 0; c < KIN; ++c) {
            dz[c] = dv[c] * s_g;
            dy[c] = y0[c] > 0.f ? dz[c] : 0.f;
          }
          uint32_t hi[8], lo[8], zh[8], zl[8];
          split16(dy, hi, lo);
          split16(dz, zh, zl);
          unsigned char* yblk = sq + tc::FB_DY0 * FB_BYTES;
          unsigned char* zblk = sq + tc::FB_DO0 * FB_BYTES;
          store_fb16(yblk, 0, hi, lo);
          store_fb16(zblk, 0, zh, zl);
          store_fb_tail(yblk, false);
          store_fb_tail(zblk, false);
          unsigned char* base = smem + SM_XIN + (p % 8) * 16 + (p / 8) * 128;
          *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(base + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          *reinterpret_cast<uint4*>(base + XIN_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          *reinterpret_cast<uint4*>(base + XIN_HALF + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          fence_async_smem();
          warp_arrive(&bars[DY0_FULL]);
        }
        // (no second e_sync: the helpers cannot write chunk buffer 0 -- the exchange area -- before the first d_o1 piece
        //  exists, and that needs DY0_FULL from all four owner warps, who have read the exchange area by then)
        // ---- B1: d_y1 = m_y1 . d_o1 chunks (d_y1 -> scratch)
        {
          const float f = sm_small[so.bf + P_U0T];
          pieces(C_PC, true, tc::FB_DY1, [&](int c, float* v) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= f;
            apply_bits16(v, (uint32_t)((c < 4 ? m_y1a : m_y1b) >> (16 * (c & 3))));
          });
        }
        // ---- B2: d_o2 -> A operand in place; d_y2 = m_y2 . d_o2 -> chunks for up_2^T; both -> scratch
        mbar_wait_parked(&bars[BDO2_FULL], ph);
        fence_after_sync();
        {
          const float f = sm_small[so.bf + P_U1T];
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int gq = 2 * i + h;
            float v[16];
            tmem_ld16(lane_t + C_SA + 16 * gq, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= f;
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_group16(lane_t + C_SA + 16 * gq, hi, lo);
            store_fb16(sq + (tc::FB_DO2 + (gq >> 1)) * FB_BYTES, 16 * (gq & 1), hi, lo);
            apply_bits16(v, (uint32_t)(m_y2 >> (16 * i)));
            split16(v, hi, lo);
            store_fb16(sq + (tc::FB_DY2 + (gq >> 1)) * FB_BYTES, 16 * (gq & 1), hi, lo);
            const int cb = cu & 1;   // chunk i = features [32 i, 32 i + 32): group 2 i from half 0, 2 i + 1 from half 1
            mbar_wait_parked(&bars[CH_EMPTY + cb], ((cu >> 1) & 1) ^ 1);
            store_chunk16(smem + SM_CHUNK + cb * CHUNK_BYTES, p, 2 * h, hi, lo);
            fence_async_smem();
            warp_arrive(&bars[CH_FULL + cb]);
            ++cu;
          }
          tmem_wait_st();
          fence_before_sync();
          warp_arrive(&bars[BO2_FULL]);
        }
        // ---- B4: d_z3 = m_r3 . d_r3 -> shared-memory A operand (chunk buffer h), scratch
        mbar_wait_parked(&bars[BD3_FULL], ph);
        fence_after_sync();
        {
          const float f = sm_small[so.bf + P_U2T];
#pragma unroll 1
          for (int i = 0; i < 2; ++i) {
            float v[16];
            tmem_ld16(lane_t + C_DR3 + 32 * h + 16 * i, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= f;
            apply_bits16(v, i == 0 ? m_r3a : m_r3b);
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_chunk16(smem + SM_CHUNK + h * CHUNK_BYTES, p, 2 * i, hi, lo);
            store_fb16(sq + (tc::FB_DZ3 + h) * FB_BYTES, 16 * i, hi, lo);
          }
        }
        fence_before_sync();
        fence_async_smem();
        warp_arrive(&bars[BZ3_FULL]);
        // ---- B5: d_z2 = m_r2 . d_r2 -> A operand in place (groups 0..3 at [0,64), 4..7 at [192,256)), scratch
        mbar_wait_parked(&bars[BR2_FULL], ph);
        fence_after_sync();
        {
          const float f = sm_small[so.bf + P_R2T];
#pragma unroll 1
          for (int i = 0; i < 4; ++i) {
            const int gq = 2 * i + h;
            const uint32_t col = gq < 4 ? C_DR2A + 16 * gq : C_DR2B + 16 * (gq - 4);
            float v[16];
            tmem_ld16(lane_t + col, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= f;
            apply_bits16(v, (uint32_t)(m_r2 >> (16 * i)));
            uint32_t hi[8], lo[8];
            split16(v, hi, lo);
            store_group16(lane_t + col, hi, lo);
            store_fb16(sq + (tc::FB_DZ2 + (gq >> 1)) * FB_BYTES, 16 * (gq & 1), hi, lo);
          }
          tmem_wait_st();
          fence_before_sync();
          warp_arrive(&bars[BZ2_FULL]);
        }
        // ---- B7: d_z1 = m_r1 . d_r1 -> scratch
        {
          const float f = sm_small[so.bf + P_D1T];
          pieces(C_BPC, false, tc::FB_DZ1, [&](int c, float* v) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= f;
            apply_bits16(v, (uint32_t)((c < 4 ? m_r1a : m_r1b) >> (16 * (c & 3))));
          });
        }
      }
      // ---- loss: warp-shuffle reduction, one fp64 atomic per owner warp
      if (h == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if ((tid & 31) == 0 && loss_acc != 0.0) atomicAdd(a.loss_sums, loss_acc);
      }
    };  // e_program
    if (warp < 4) e_program(std::integral_constant<int, 0>{});
    else e_program(std::integral_constant<int, 1>{});
  } else if (warp == 8) {
    // =================================================================== M: MMA issue
    uint32_t ws = 0, cm = 0, pm = 0;
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
    // one 32-column piece from the shared-memory K = 16 operand (xin / d_y0) and a [32 x 16] weight block
    auto small_piece = [&](uint32_t pc_col, uint32_t b_smem) {
      const uint32_t b = pm & 1;
      mbar_wait_parked(&bars[PC_EMPTY + b], ((pm >> 1) & 1) ^ 1);
      fence_after_sync();
      if (elect_one()) {
        issue_ss<32, 1>(tm + pc_col + 32 * b, xin_s, XIN_HALF, b_smem, KIN, 32, true);
        commit(&bars[PC_FULL + b]);
      }
      __syncwarp();
      ++pm;
    };
    // A = 8 shared-memory chunks of 32 features against [N x 32] blocks; a small piece rides in slots 0..5
    auto chunk_layer = [&](auto n_const, uint32_t d_col, uint32_t slab_rows, uint32_t piece_off) {
      constexpr int N = decltype(n_const)::value;
      for (int c = 0; c < 8; ++c) {
        const uint32_t wb = wait_w();
        if (c < 6) small_piece(C_PC, wb + piece_off);
        const uint32_t b = cm & 1;
        mbar_wait_parked(&bars[CH_FULL + b], (cm >> 1) & 1);
        fence_after_sync();
        if (elect_one()) {
          issue_ss<N, 2>(tm + d_col, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, 32, slab_rows, c == 0);
          commit(&bars[CH_EMPTY + b]);
        }
        __syncwarp();
        ++cm;
        release_w();
      }
    };
    for (uint32_t g = 0; g < (uint32_t)my_tiles; ++g) {
      const uint32_t ph = g & 1;
      // ================= forward (as rollout_h.cu)
      wait_e(XIN_FULL, ph);
      {
        const uint32_t wb = wait_w();
        small_piece(C_PC, wb);
        small_piece(C_PC, wb + PIECE_BYTES);
        release_w();
      }
      chunk_layer(std::integral_constant<int, H1 + NY>{}, C_SA, H1 + NY, D1_MAIN);   // down_1 (+ Wc r1)
      signal(D1_FULL);
      wait_e(R2H_FULL, ph);
      for (int j = 0; j < 2; ++j) {   // down_2
        if (j == 1) wait_e(R2_FULL, ph);
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<H2, 4>(tm + C_D2, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D2_FULL);
      for (int j = 0; j < 2; ++j) {   // res_2 b
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_R2B, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      wait_e(R3_FULL, ph);
      for (int j = 0; j < 2; ++j) {   // res_2 a
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_R2A, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      signal(R2A_DONE);               // up_2 overwrites r2, which res_2 is still reading
      wait_e(R2A_DONE, ph);
      for (int j = 0; j < 2; ++j) {   // up_2
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ss<H1, 2>(tm + C_SA, chunk_s + j * CHUNK_BYTES, CHUNK_HALF, wb, 32, H1, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D3_FULL);
      wait_e(O2_FULL, ph);
      for (int q = 0; q < 8; ++q) {   // up_1 pieces; folded up_0 two pieces behind
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
        if (q >= 2) {
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
      // ================= backward
      wait_e(DY0_FULL, ph);
      {
        const uint32_t wb = wait_w();   // d_o1 pieces 0, 1
        small_piece(C_PC, wb);
        small_piece(C_PC, wb + PIECE_BYTES);
        release_w();
      }
      chunk_layer(std::integral_constant<int, H1>{}, C_SA, H1, 16384);   // d_o2 = d_y1 W_u1 (+ d_o1 pieces 2..7)
      signal(BDO2_FULL);
      // d_r3 = d_y2 W_u2: four K = 32 chunks from the ring, two weight blocks per slot
      for (int j = 0; j < 4; ++j) {
        uint32_t wb = 0;
        if ((j & 1) == 0) wb = wait_w();
        else wb = ring_s + (ws % NSTAGE) * SLOT_BYTES;
        const uint32_t cb = cm & 1;
        mbar_wait_parked(&bars[CH_FULL + cb], (cm >> 1) & 1);
        fence_after_sync();
        if (elect_one()) {
          issue_ss<H2, 2>(tm + C_DR3, chunk_s + cb * CHUNK_BYTES, CHUNK_HALF, wb + (j & 1) * 8192, 32, H2, j == 0);
          commit(&bars[CH_EMPTY + cb]);
        }
        __syncwarp();
        ++cm;
        if (j & 1) release_w();
      }
      signal(BD3_FULL);
      wait_e(BO2_FULL, ph);           // d_r2 b = d_o2 W_r2 (output half b) -> [192,256)
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_DR2B, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      wait_e(BZ3_FULL, ph);           // d_r3 has been read: d_r2 a -> [0,64)
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_ts<64, 4>(tm + C_DR2A, tm + C_SA + 64 * j, wb, 64, j == 0);
        __syncwarp();
        release_w();
      }
      for (int hb = 0; hb < 2; ++hb) {   // d_r2 += d_z3 W_d2: A = d_z3 in the two chunk buffers, output halves a, b
        const uint32_t wb = wait_w();
        if (elect_one()) {
          for (int j = 0; j < 2; ++j)
            issue_ss<64, 2>(tm + (hb == 0 ? C_DR2A : C_DR2B), chunk_s + j * CHUNK_BYTES, CHUNK_HALF, wb + j * 8192, 32, 64,
                            false);
        }
        __syncwarp();
        release_w();
      }
      signal(BR2_FULL);
      wait_e(BZ2_FULL, ph);           // d_r1 = d_z2 W_d1 + d_y0 Wc in 8 pieces of 32 -> [64,128)
      for (int q = 0; q < 8; ++q) {
        const uint32_t wb = wait_w();
        const uint32_t b = pm & 1;
        mbar_wait_parked(&bars[PC_EMPTY + b], ((pm >> 1) & 1) ^ 1);
        fence_after_sync();
        if (elect_one()) {
          issue_ts2<32, 8>(tm + C_BPC + 32 * b, tm + C_DR2A, tm + C_DR2B, wb, H1, true);
          issue_ss<32, 1>(tm + C_BPC + 32 * b, xin_s, XIN_HALF, wb + U1_MAIN, KIN, 32, false);
          commit(&bars[PC_FULL + b]);
        }
        __syncwarp();
        ++pm;
        release_w();
      }
    }
  } else {
    // =================================================================== P: weight tape producer (forward then backward tape)
    if (elect_one()) {
      const uint32_t per_tile = FWD_SLOTS + BWD_SLOTS;
      const uint32_t total = (uint32_t)my_tiles * per_tile;
      uint32_t in_tile = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % NSTAGE;
        mbar_wait_parked(&bars[W_EMPTY + s], ((i / NSTAGE) & 1) ^ 1);
        const bool bwd = in_tile >= (uint32_t)FWD_SLOTS;
        const uint32_t bytes = bwd ? bwd_slot_bytes((int)in_tile - FWD_SLOTS) : fwd_slot_bytes((int)in_tile);
        mbar_expect_tx(&bars[W_FULL + s], bytes);
        bulk_g2s(smem + SM_RING + s * SLOT_BYTES, tape + (size_t)in_tile * SLOT_BYTES, bytes, &bars[W_FULL + s]);
        if (++in_tile == per_tile) in_tile = 0;
      }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 256);
}

// ---------------------------------------------------------------- host side
// Tiles per K3a / K3b sub-launch pair = size of the activation scratch (944 KB per tile: 15.5 GB at the default).  Bigger
// sub-launches amortise the pipeline fill / drain of the persistent CTAs (measured: 8192 -> 16384 tiles -1.5% K3 time,
// 32768 another -0.3%); SOCM_SUB_TILES overrides for experiments.
static int sub_tiles_cap() {
  static const int cap = [] {
    const char* e = getenv("SOCM_SUB_TILES");
    return e != nullptr && atoi(e) >= 1024 ? atoi(e) : 16384;
  }();
  return cap;
}
static int sub_tiles_max() { return (sub_tiles_cap() / (2 * sm_count())) * (2 * sm_count()); }
static int64_t ws_head_bytes() { return ((workspace_bytes() + 1023) / 1024) * 1024; }
constexpr int64_t AUX_BYTES = ((tc::AUX_FLOATS * 4 + 1023) / 1024) * 1024;

bool loss_h_supported(const socm_unet* net) { return is_default_arch(net) && net->d <= MAX_D; }

int64_t loss_h_workspace_bytes(int B, int K) {
  const int64_t n_tiles = (int64_t)(K + 1) * ((B + TP - 1) / TP);
  const int64_t sub = n_tiles < sub_tiles_cap() ? n_tiles : sub_tiles_cap();
  return ws_head_bytes() + AUX_BYTES + sub * TILE_BYTES + 2048;
}

int launch_loss_h(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream) {
  SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_loss_workspace_bytes)");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* aux = reinterpret_cast<float*>(ws + ws_head_bytes());
  unsigned char* scratch = ws + ws_head_bytes() + AUX_BYTES;
  scratch += (1024 - (reinterpret_cast<uintptr_t>(scratch) & 1023)) & 1023;
  CalibArgs c{};
  c.states = a.states;
  c.ts = a.ts;
  c.target = a.target;
  c.B = a.B;
  c.K = a.K;
  c.ldt = a.ldt;
  c.w = a.w;
  c.n_samples = 512;
  c.lmbd = a.st.lmbd;
  c.loss_scale = a.scale;
  if (int rc = setup_h(net, ws, c, true, stream)) return rc;
  SOCM_CUDA(cudaMemsetAsync(aux, 0, tc::AUX_FLOATS * sizeof(float), stream));
  const int smem = loss_h_smem_bytes();
  SOCM_CUDA(cudaFuncSetAttribute(loss_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int n_mblk = (a.B + TP - 1) / TP;
  const int64_t n_tiles = (int64_t)(a.K + 1) * n_mblk;
  const int sub_max = sub_tiles_max();
  for (int64_t t0 = 0; t0 < n_tiles; t0 += sub_max) {
    const int nt = (int)((n_tiles - t0) < sub_max ? (n_tiles - t0) : sub_max);
    const int grid = nt < 2 * sm_count() ? nt : 2 * sm_count();
    loss_h_kernel<<<grid, k3::NT, smem, stream>>>(a, ws, small_ptr(ws), scratch, (int)t0, nt);
    SOCM_LAUNCH_CHECK();
    if (int rc = launch_wgrad_h(scratch, nt, a.st.d, small_ptr(ws) + small_layout().sk, grad, aux, stream)) return rc;
  }
  tc::fold_finish_kernel<<<H0, H0, 0, stream>>>(*net, aux, grad);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace hx
}  // namespace socm
