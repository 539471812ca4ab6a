// loss_tc.cu -- K3a: UNet forward at 128 trajectory points, weighted loss and the whole dgrad chain
// on the 5th-gen tensor cores (tcgen05, 3xTF32; engine of unet_tc.cuh), one persistent CTA per SM.
// Replaces method.py:272-287 (nabla_V at all points), 692-720 (loss) and the activation-gradient
// half of loss.backward() (main.py:323).  The weight gradients are K3b (wgrad_tc.cu): this kernel
// leaves every operand they need -- activations and activation gradients, fp32 -- in the scratch
// layout of loss_tc.cuh.
//
// A tile is 128 consecutive paths at one grid time t_i.  Per tile:
//   forward   exactly the rollout's forward (rollout_tc.cu) from the stored state, keeping the ReLU
//             masks in registers (thread <-> point, fixed column ownership);
//   loss      per point: diff = nabla_V - target, loss += s w |sigma^T diff|^2, G = -d loss / d nabla_V;
//   backward  the mirror image of the forward with the transposed weight tape:
//               d_o1 = W_u0^T d_y0            (chunk source in TMEM, like down_0)
//               d_o2 = W_u1^T (m_y1 . d_o1)   (A = 32-feature chunks in shared memory, like down_1)
//               d_r2 = W_r2^T d_o2 + W_d2^T (m_r3 . W_u2^T (m_y2 . d_o2))
//               d_r1 = W_d1^T (m_r2 . d_r2) + Wc^T d_y0      (res_1 folded into up_0, unet_tc.cuh: K = d)
//               d_z1 = m_r1 . d_r1
// TMEM columns (backward): [0,256) d_o1 accumulator / A operands (hi|lo);  [256,384) d_o2 acc, later
// [256,320) d_r3 acc;  [384,512) d_r2 acc;  [256,512) d_r1 acc.
#include <type_traits>

#include "kernels.h"
#include "loss_common.cuh"
#include "loss_tc.cuh"
#include "unet_tc.cuh"

namespace socm {
namespace tc {

using namespace umma;

#ifdef SOCM_TC_PROF
__device__ unsigned long long g_k3_prof[192];   // [E owner | E helper | M] x 64 phase slots (block 0), debug builds only
#define K3P_DECL long long prof_t = clock64(); unsigned long long prof_acc[64] = {0}; int prof_i = 0
#define K3P_RESET prof_i = 0
#define K3P_MARK do { const long long t_ = clock64(); prof_acc[prof_i++ & 63] += (unsigned long long)(t_ - prof_t); prof_t = t_; } while (0)
#define K3P_FLUSH(base, cond) do { if (blockIdx.x == 0 && (cond)) for (int i_ = 0; i_ < 64; ++i_) g_k3_prof[(base) + i_] = prof_acc[i_]; } while (0)
#else
#define K3P_DECL
#define K3P_RESET
#define K3P_MARK
#define K3P_FLUSH(base, cond)
#endif

namespace k3 {
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * SLOT_BYTES;
constexpr int SM_XIN = SM_CHUNK + 2 * CHUNK_BYTES;  // [t,x] operand (forward) / d_y0 operand (backward)
constexpr int SM_SMALL = SM_XIN + MAX_KIN * 1024;
enum Bar {
  W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2,
  // forward, one completion per tile each
  XIN_FULL = CH_EMPTY + 2, D0_FULL, D1_FULL, R2_FULL, D2_FULL, R3_FULL, D3A_FULL, D3B_FULL, O2_FULL,
  D4A_FULL, Y0_FULL,
  // backward
  DY0_FULL, BD0_FULL, BDO2_FULL, BO2_FULL, BR2A_FULL, BY2_FULL, BD3_FULL, BZ3_FULL, BR2B_FULL, BZ2_FULL,
  BD1B_FULL,
  // first column half of a 128-wide A operand is ready (the MMAs over its K blocks start while the epilogue
  // threads still work on the second half)
  // (only where the next layer's accumulator does not overlap the columns the second half is still read from:
  //  up_1 and down_1^T write [256,512) and must wait for the whole epilogue)
  R2H_FULL, BO2H_FULL, BY2H_FULL, N_BARS
};
constexpr int NT = 320, NE = 256;
constexpr uint32_t C_SA = 0, C_D1 = 256, C_D2 = 256, C_D3 = 256, C_D3R = 384, C_D4 = 256;
constexpr uint32_t C_DO2 = 256, C_DR2 = 384, C_DR3 = 256, C_DR1 = 256;
constexpr uint32_t C_Y0P = 384, C_Y0 = 0;
static_assert(C_Y0P == C_D1 + H1, "Wc r1 must sit right behind the down_1 accumulator (joint MMA)");  // folded last layer: Wc r1 (next to down_1) and W_u0 y1 (after up_1)
}  // namespace k3

__host__ __device__ inline int loss_tc_smem_bytes(int d) {
  return k3::SM_SMALL + small_tc(d).total * 4 + k3::N_BARS * 8 + 16;
}

// Scratch stores (layout of loss_tc.cuh): thread <-> point r of the quarter; `ro` = the eight
// lane-dependent chunk offsets ((r >> 2) ^ k) * 16 + (r & 3) * 4, k = f & 7, precomputed once.
struct RowOff {
  int o[8];
  __device__ __forceinline__ explicit RowOff(int r) {
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ((((r >> 2) ^ k) & 7) << 4) + (r & 3) * 4;
  }
};
// features [f0, f0+16) of feature block `blk` (one coalesced 128-byte line per warp-wide store)
__device__ __forceinline__ void store_fb16(unsigned char* blk, const RowOff& ro, int f0, const float* v) {
#ifdef SOCM_K3_NOSTORE
  return;
#endif
#pragma unroll
  for (int j = 0; j < 16; ++j) __stcs(reinterpret_cast<float*>(blk + (f0 + j) * 128 + ro.o[(f0 + j) & 7]), v[j]);
}
__device__ __forceinline__ void store_fb32(unsigned char* blk, const RowOff& ro, const float* v) {
#ifdef SOCM_K3_NOSTORE
  return;
#endif
#pragma unroll
  for (int j = 0; j < 32; ++j) __stcs(reinterpret_cast<float*>(blk + j * 128 + ro.o[j & 7]), v[j]);
}
template <int NV>
__device__ __forceinline__ uint32_t positive_bits(const float* v) {
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) m |= (v[j] > 0.f ? 1u : 0u) << j;
  return m;
}
template <int NV>
__device__ __forceinline__ void apply_bits(float* v, uint32_t m) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = ((m >> j) & 1u) ? v[j] : 0.f;
}
template <int KIN>
__global__ void __launch_bounds__(k3::NT, 1)
    loss_tc_kernel(LossArgs a, const unsigned char* __restrict__ tape, const float* __restrict__ small_g,
                   unsigned char* __restrict__ scratch, int tile0, int n_tiles_launch) {
  using namespace k3;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int d = a.st.d, K = a.K, B = a.B;
  const SmallTc so = small_tc(d);
  float* sm_small = reinterpret_cast<float*>(smem + SM_SMALL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_SMALL + so.total * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_mblk = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles_launch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool simple_loss = a.st.sigma_is_identity && a.warmA == nullptr;
  constexpr int S0 = KIN > 16 ? 2 : 1;
  constexpr int NY = KIN <= 16 ? 16 : 32;  // N of the folded up_0 MMAs
  constexpr int NU0 = NY == 16 ? 1 : 2;    // up_0 slots
  constexpr int BPS = 8 / NU0;             // up_0 K-chunks per slot
  constexpr int NSF = S0 + 24 + NU0;       // forward weight stages per tile (= forward tape slots, in order)
  constexpr int NSB = 2 * S0 + 24;         // backward weight stages per tile
  (void)K;

  for (int i = tid; i < so.total; i += NT) sm_small[i] = __ldg(small_g + i);
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bars[W_FULL + s], 1);
      mbar_init(&bars[W_EMPTY + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[CH_FULL + b], NE / 32);
      mbar_init(&bars[CH_EMPTY + b], 1);
    }
    mbar_init(&bars[XIN_FULL], TP / 32);
    mbar_init(&bars[DY0_FULL], TP / 32);
    const int e2m[] = {R2_FULL, R3_FULL, O2_FULL, BO2_FULL, BY2_FULL, BZ3_FULL, BZ2_FULL,
                       R2H_FULL, BO2H_FULL, BY2H_FULL};
    for (int i = 0; i < 10; ++i) mbar_init(&bars[e2m[i]], NE / 32);
    const int m2e[] = {D0_FULL, D1_FULL, D2_FULL, D3A_FULL, D3B_FULL, D4A_FULL, Y0_FULL, BD0_FULL,
                       BDO2_FULL, BR2A_FULL, BD3_FULL, BR2B_FULL, BD1B_FULL};
    for (int i = 0; i < 13; ++i) mbar_init(&bars[m2e[i]], 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK), xin_s = smem_addr(smem + SM_XIN);

  if (warp < 8) {
    // =================================================================== E: epilogue / point threads
    auto e_program = [&](const int h) {  // column half; h == 0 threads own the point.  ONE copy of the code for both
      // halves: the two warps of a scheduler (an owner and a helper) then fetch the same instructions
      const int p = tid & (TP - 1), r = tid & 31, q = (tid >> 5) & 3;
      const RowOff ro(r);
      const uint32_t lane_t = tm + ((uint32_t)(q * 32) << 16);
      uint32_t g = 0;   // tiles done
      uint32_t cu = 0;  // chunks produced
      float* xin_hi = reinterpret_cast<float*>(smem + SM_XIN);
      float* xin_lo = reinterpret_cast<float*>(smem + SM_XIN + KIN * 512);
      double loss_acc = 0.0;
      K3P_DECL;

      for (int lt = blockIdx.x; lt < n_tiles_launch; lt += gridDim.x, ++g) {
        const int t = tile0 + lt;
        const int ti = t / n_mblk, m0 = (t - ti * n_mblk) * TP;
        const int m = m0 + p;
        const bool live = m < B;
        const uint32_t ph = g & 1;
        unsigned char* sq = scratch + (size_t)lt * TILE_BYTES + (size_t)q * QUARTER_BYTES;  // this warp's quarter
        // ReLU masks of this thread's columns, packed so that the rolled loops below need no indexed registers
        // (the E program is executed once per tile and must stay small: instruction fetch, not issue, limits it)
        uint64_t m_r1a = 0, m_r1b = 0, m_y1a = 0, m_y1b = 0, m_r2 = 0, m_y2 = 0;
        uint32_t m_r3 = 0;
        float x[KIN];
        K3P_RESET;
        K3P_MARK;

        // chunk producer: v(16 cols) = f(accumulator columns) -> shared-memory A chunk; `fn` post-processes the 16
        // values.  The TMEM load of chunk c+1 is in flight while chunk c is processed (a tcgen05.ld that competes
        // with running MMAs takes ~450 cycles), and the chunk buffer is waited for last: nothing before the
        // shared-memory stores depends on it.
        auto chunks = [&](uint32_t src_col, auto&& fn) {
          auto body = [&](int c, float* v) {
            fn(c, v);
            const int b = cu & 1;
            mbar_wait(&bars[CH_EMPTY + b], ((cu >> 1) & 1) ^ 1);
            store_chunk16(smem + SM_CHUNK + b * CHUNK_BYTES, p, 4 * h, v);
            fence_async_smem();
            warp_arrive(&bars[CH_FULL + b]);
            ++cu;
          };
          float va[16], vb[16];
          tmem_ld16(lane_t + src_col + 16 * h, reinterpret_cast<uint32_t*>(va));
#pragma unroll 1
          for (int c = 0; c < 8; c += 2) {
            tmem_wait_ld();
            tmem_ld16(lane_t + src_col + 32 * (c + 1) + 16 * h, reinterpret_cast<uint32_t*>(vb));
            body(c, va);
            tmem_wait_ld();
            if (c + 2 < 8) tmem_ld16(lane_t + src_col + 32 * (c + 2) + 16 * h, reinterpret_cast<uint32_t*>(va));
            body(c + 1, vb);
          }
        };

        // ---- F0 (owners): state -> input operand (hi / lo) and the XIN block of the scratch
        if (h == 0) {
          const float tk = __ldg(a.ts + ti);
#pragma unroll
          for (int j = 0; j < KIN; ++j)
            x[j] = (j < d && live) ? __ldg(a.states + ((size_t)ti * B + m) * d + j) : 0.f;
          float xb[KIN];
          xb[0] = tk;
#pragma unroll
          for (int c = 1; c < KIN; ++c) xb[c] = x[c - 1];
#pragma unroll
          for (int c4 = 0; c4 < KIN / 4; ++c4) {
            float4 hh, ll;
            hh.x = tf32_rn(xb[4 * c4]); hh.y = tf32_rn(xb[4 * c4 + 1]); hh.z = tf32_rn(xb[4 * c4 + 2]); hh.w = tf32_rn(xb[4 * c4 + 3]);
            ll.x = xb[4 * c4] - hh.x; ll.y = xb[4 * c4 + 1] - hh.y; ll.z = xb[4 * c4 + 2] - hh.z; ll.w = xb[4 * c4 + 3] - hh.w;
            const int off = (p % 8) * 4 + (p / 8) * 32 + c4 * 512;
            *reinterpret_cast<float4*>(xin_hi + off) = hh;
            *reinterpret_cast<float4*>(xin_lo + off) = ll;
          }
          fence_async_smem();
          warp_arrive(&bars[XIN_FULL]);
          // XIN block of the scratch: features [t, x, 0.., 1 at ONES_FEATURE]
          unsigned char* xblk = sq + FB_XIN * FB_BYTES;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float val = c < KIN ? xb[c < KIN ? c : 0] : (c == ONES_FEATURE ? 1.0f : 0.f);
            *reinterpret_cast<float*>(xblk + c * 128 + ro.o[c & 7]) = val;
          }
        }
        // ---- F1: r1 chunks for down_1 (+ mask, + scratch)
        K3P_MARK;
        mbar_wait(&bars[D0_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        chunks(C_SA, [&](int c, float* v) {
          bias_relu16(v, sm_small + so.b_d0 + 32 * c + 16 * h);
          const uint64_t bits = (uint64_t)positive_bits<16>(v) << (16 * (c & 3));
          if (c < 4) m_r1a |= bits;
          else m_r1b |= bits;
          store_fb16(sq + (FB_R1 + c) * FB_BYTES, ro, 16 * h, v);
        });
        // ---- F2: r2
        K3P_MARK;
        mbar_wait(&bars[D1_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        float au[KIN];  // owners: (Wc r1)[j], later the whole pre-activation of up_0
        if (h == 0) {
          float yp[NY];
          if constexpr (NY == 16) tmem_ld16(lane_t + C_Y0P, reinterpret_cast<uint32_t*>(yp));
          else tmem_ld32(lane_t + C_Y0P, reinterpret_cast<uint32_t*>(yp));
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < KIN; ++j) au[j] = yp[j];
        }
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int cb = 2 * i + h;  // iteration 0 covers columns [0,64), iteration 1 [64,128)
          float v[32];
          tmem_ld32(lane_t + C_D1 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d1 + 32 * cb);
          m_r2 |= (uint64_t)positive_bits<32>(v) << (32 * i);
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
          store_fb32(sq + (FB_R2 + cb) * FB_BYTES, ro, v);
          if (i == 0) {
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[R2H_FULL]);
          }
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[R2_FULL]);
        // ---- F3: r3 -> shared-memory A operand (features [32h, 32h+32) = chunk buffer h), mask, scratch
        K3P_MARK;
        mbar_wait(&bars[D2_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        {
          float v[32];
          tmem_ld32(lane_t + C_D2 + 32 * h, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d2 + 32 * h);
          m_r3 = positive_bits<32>(v);
          store_chunk32(smem + SM_CHUNK + h * CHUNK_BYTES, p, v);
          store_fb32(sq + (FB_R3 + h) * FB_BYTES, ro, v);
        }
        fence_before_sync();
        fence_async_smem();
        warp_arrive(&bars[R3_FULL]);
        // ---- F5: o2 = relu(D3 + b_u2) + D3R + b_r2 (up_2 and res_2 have both finished) -> A operand, mask, scratch
        K3P_MARK;
        mbar_wait(&bars[D3A_FULL], ph);
        mbar_wait(&bars[D3B_FULL], ph);
        K3P_MARK;
        fence_after_sync();
#pragma unroll 1
        for (int ii = 0; ii < 4; ++ii) {
          const int i = ii >> 1, hh = ii & 1;
          const int cb = 2 * i + h;  // iteration i covers columns [64 i, 64 i + 64) of both halves
          const int c0 = 32 * cb + 16 * hh;
          float y[16], rr[16];
          tmem_ld16(lane_t + C_D3 + c0, reinterpret_cast<uint32_t*>(y));
          tmem_ld16(lane_t + C_D3R + c0, reinterpret_cast<uint32_t*>(rr));
          tmem_wait_ld();
          bias_relu16(y, sm_small + so.b_u2 + c0);
          m_y2 |= (uint64_t)positive_bits<16>(y) << (32 * i + 16 * hh);
          const float* br = sm_small + so.b_r2 + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] += rr[j] + br[j];
          store_split16(lane_t + C_SA + c0, lane_t + C_SA + 128 + c0, y);
          store_fb16(sq + (FB_O2 + cb) * FB_BYTES, ro, 16 * hh, y);
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[O2_FULL]);
        // up_1 runs for ~6k cycles now: the owner half evaluates res_0 [t, x] meanwhile (it depends on the state only)
        float ar[KIN];
#pragma unroll
        for (int j = 0; j < KIN; ++j) ar[j] = 0.f;
        if (h == 0) {
          const float tk = __ldg(a.ts + ti);
#pragma unroll
          for (int j = 0; j < KIN; ++j) {
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
            ar[j] = acc_r;
          }
        }
        // ---- F6: y1 = relu(D4 + b_u1) -> A chunks of the folded up_0 (+ mask, + scratch)
        K3P_MARK;
        mbar_wait(&bars[D4A_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        chunks(C_D4, [&](int c, float* v) {
          bias_relu16(v, sm_small + so.b_u1 + 32 * c + 16 * h);
          const uint64_t bits = (uint64_t)positive_bits<16>(v) << (16 * (c & 3));
          if (c < 4) m_y1a |= bits;
          else m_y1b |= bits;
          store_fb16(sq + (FB_Y1 + c) * FB_BYTES, ro, 16 * h, v);
        });
        // ---- F8 (owners): y0 = W_u0 y1 (TMEM [0,NY)) + Wc r1 (registers) + bc
        K3P_MARK;
        if (h == 0) {
          mbar_wait(&bars[Y0_FULL], ph);
          fence_after_sync();
          float yp[NY];
          if constexpr (NY == 16) tmem_ld16(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
          else tmem_ld32(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < KIN; ++j) au[j] += yp[j];
          fence_before_sync();  // TMEM reads before the backward MMAs into [0,256) (ordered through DY0_FULL)
        }
        K3P_MARK;
        // ---- loss (owners): nabla_V, d loss / d nabla_V, G; d_y0 operand for the backward pass
        if (h == 0) {
          float gv[KIN], y0[KIN], dv[KIN];
#pragma unroll
          for (int j = 0; j < KIN; ++j) {
            y0[j] = au[j] + sm_small[so.bc + j];
            gv[j] = fmaxf(y0[j], 0.f) + ar[j];
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
          float dy[KIN];
          unsigned char* yblk = sq + FB_DY0 * FB_BYTES;
          unsigned char* zblk = sq + FB_DO0 * FB_BYTES;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float vy = 0.f, vz = 0.f;
            if (c < KIN) {
              vz = dv[c < KIN ? c : 0];                               // d_o0
              vy = y0[c < KIN ? c : 0] > 0.f ? vz : 0.f;              // d_y0
              dy[c < KIN ? c : 0] = vy;
            }
            *reinterpret_cast<float*>(yblk + c * 128 + ro.o[c & 7]) = vy;
            *reinterpret_cast<float*>(zblk + c * 128 + ro.o[c & 7]) = vz;
          }
#pragma unroll
          for (int c4 = 0; c4 < KIN / 4; ++c4) {
            float4 hh, ll;
            hh.x = tf32_rn(dy[4 * c4]); hh.y = tf32_rn(dy[4 * c4 + 1]); hh.z = tf32_rn(dy[4 * c4 + 2]); hh.w = tf32_rn(dy[4 * c4 + 3]);
            ll.x = dy[4 * c4] - hh.x; ll.y = dy[4 * c4 + 1] - hh.y; ll.z = dy[4 * c4 + 2] - hh.z; ll.w = dy[4 * c4 + 3] - hh.w;
            const int off = (p % 8) * 4 + (p / 8) * 32 + c4 * 512;
            *reinterpret_cast<float4*>(xin_hi + off) = hh;
            *reinterpret_cast<float4*>(xin_lo + off) = ll;
          }
          fence_async_smem();
          warp_arrive(&bars[DY0_FULL]);
        }
        // ---- B1: d_y1 = m_y1 . d_o1 chunks (d_o1, d_y1 -> scratch)
        K3P_MARK;
        mbar_wait(&bars[BD0_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        chunks(C_SA, [&](int c, float* v) {
          apply_bits<16>(v, (uint32_t)((c < 4 ? m_y1a : m_y1b) >> (16 * (c & 3))));
          store_fb16(sq + (FB_DY1 + c) * FB_BYTES, ro, 16 * h, v);
        });
        // ---- B2: d_o2 -> A operand; d_o2, d_y2 -> scratch
        K3P_MARK;
        mbar_wait(&bars[BDO2_FULL], ph);
        K3P_MARK;
        fence_after_sync();
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int cb = 2 * i + h;  // iteration 0 covers columns [0,64), iteration 1 [64,128)
          float v[32];
          tmem_ld32(lane_t + C_DO2 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
          store_fb32(sq + (FB_DO2 + cb) * FB_BYTES, ro, v);
          apply_bits<32>(v, (uint32_t)(m_y2 >> (32 * i)));
          store_fb32(sq + (FB_DY2 + cb) * FB_BYTES, ro, v);
          if (i == 0) {
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[BO2H_FULL]);
          }
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[BO2_FULL]);
        // ---- B3: d_y2 -> A operand (once res_2^T has finished reading d_o2)
        K3P_MARK;
        mbar_wait(&bars[BR2A_FULL], ph);
        K3P_MARK;
        fence_after_sync();
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int cb = 2 * i + h;  // iteration 0 covers columns [0,64), iteration 1 [64,128)
          float v[32];
          tmem_ld32(lane_t + C_DO2 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          apply_bits<32>(v, (uint32_t)(m_y2 >> (32 * i)));
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
          if (i == 0) {
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[BY2H_FULL]);
          }
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[BY2_FULL]);
        // ---- B4: d_z3 = m_r3 . d_r3 -> A operand [0,64) hi, [64,128) lo
        K3P_MARK;
        mbar_wait(&bars[BD3_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        {
          float v[32];
          tmem_ld32(lane_t + C_DR3 + 32 * h, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          apply_bits<32>(v, m_r3);
          store_split32(lane_t + C_SA + 32 * h, lane_t + C_SA + 64 + 32 * h, v);
          store_fb32(sq + (FB_DZ3 + h) * FB_BYTES, ro, v);
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[BZ3_FULL]);
        // ---- B5: d_z2 = m_r2 . d_r2 -> A operand
        K3P_MARK;
        mbar_wait(&bars[BR2B_FULL], ph);
        K3P_MARK;
        fence_after_sync();
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const int cb = 2 * i + h;  // iteration 0 covers columns [0,64), iteration 1 [64,128)
          float v[32];
          tmem_ld32(lane_t + C_DR2 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          apply_bits<32>(v, (uint32_t)(m_r2 >> (32 * i)));
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
          store_fb32(sq + (FB_DZ2 + cb) * FB_BYTES, ro, v);
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[BZ2_FULL]);
        // ---- B7: d_z1 = m_r1 . d_r1 -> scratch
        K3P_MARK;
        mbar_wait(&bars[BD1B_FULL], ph);
        K3P_MARK;
        fence_after_sync();
        {
          auto body = [&](int c, float* v) {
            apply_bits<16>(v, (uint32_t)((c < 4 ? m_r1a : m_r1b) >> (16 * (c & 3))));
            store_fb16(sq + (FB_DZ1 + c) * FB_BYTES, ro, 16 * h, v);
          };
          float va[16], vb[16];
          tmem_ld16(lane_t + C_DR1 + 16 * h, reinterpret_cast<uint32_t*>(va));
#pragma unroll 1
          for (int c = 0; c < 8; c += 2) {
            tmem_wait_ld();
            tmem_ld16(lane_t + C_DR1 + 32 * (c + 1) + 16 * h, reinterpret_cast<uint32_t*>(vb));
            body(c, va);
            tmem_wait_ld();
            if (c + 2 < 8) tmem_ld16(lane_t + C_DR1 + 32 * (c + 2) + 16 * h, reinterpret_cast<uint32_t*>(va));
            body(c + 1, vb);
          }
        }
        fence_before_sync();  // TMEM reads done before the next tile's MMAs (ordered through XIN_FULL / e_sync)
        K3P_MARK;
        e_sync();
        K3P_MARK;
      }
      K3P_FLUSH(h * 64, (tid & 127) == 0);
      // ---- loss: warp-shuffle reduction, one fp64 atomic per owner warp
      if (h == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if ((tid & 31) == 0 && loss_acc != 0.0) atomicAdd(a.loss_sums, loss_acc);
      }
    };  // e_program
    e_program(warp >> 2);
  } else if (warp == 8) {
    // =================================================================== M: MMA issue
    uint32_t ws = 0, cm = 0;
    auto wait_w = [&]() -> uint32_t {
      const uint32_t s = ws % NSTAGE;
      mbar_wait(&bars[W_FULL + s], (ws / NSTAGE) & 1);
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
      mbar_wait(&bars[bar], ph);
      fence_after_sync();
    };
    // 256 columns at d_col (+)= (SMEM operand [128 x KIN]) x (S0 tape blocks): down_0 / up_0^T / Wc^T
    auto small_k = [&](uint32_t d_col, bool fresh) {
#pragma unroll
      for (int hh = 0; hh < S0; ++hh) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ss<H0 / S0, KIN>(tm + d_col + hh * (H0 / S0), xin_s, KIN * 512, wb, fresh);
        __syncwarp();
        release_w();
      }
    };
    // A = 8 shared-memory chunks of 32 features; one block of K = 32 (N = 128) or two of K = 16 (N = 256) per chunk
    auto chunk_layer_128 = [&](uint32_t d_col, bool with_wc) {  // with_wc: the slot also carries the Wc block (forward)
      for (int c = 0; c < 8; ++c) {
        const uint32_t b = cm & 1;
        mbar_wait(&bars[CH_FULL + b], (cm >> 1) & 1);
        fence_after_sync();
        const uint32_t wb = wait_w();
        if (elect_one()) {
          // forward: down_1 and Wc r1 in one MMA (N = H1 + NY: columns [C_D1, C_D1 + H1) and [C_Y0P, C_Y0P + NY))
          if (with_wc) issue_block_ss<H1 + NY, 32>(tm + d_col, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, c == 0);
          else issue_block_ss<H1, 32>(tm + d_col, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, c == 0);
          commit(&bars[CH_EMPTY + b]);
        }
        __syncwarp();
        release_w();
        ++cm;
      }
    };
    auto up0_layer = [&]() {  // folded up_0: [0,NY) = y1 chunks x W_u0 blocks (8 K-chunks of 32, BPS per slot)
      for (int u = 0; u < NU0; ++u) {
        const uint32_t wb = wait_w();
        for (int cc = 0; cc < BPS; ++cc) {
          const uint32_t b = cm & 1;
          mbar_wait(&bars[CH_FULL + b], (cm >> 1) & 1);
          fence_after_sync();
          if (elect_one()) {
            issue_block_ss<NY, 32>(tm + C_Y0, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb + cc * (NY * 256), u == 0 && cc == 0);
            commit(&bars[CH_EMPTY + b]);
          }
          __syncwarp();
          ++cm;
        }
        release_w();
      }
    };
    for (uint32_t g = 0; g < (uint32_t)my_tiles; ++g) {
      const uint32_t ph = g & 1;
      // ================= forward
      wait_e(XIN_FULL, ph);
      small_k(C_SA, true);
      signal(D0_FULL);
      chunk_layer_128(C_D1, true);  // down_1 (+ Wc r1)
      signal(D1_FULL);
      wait_e(R2H_FULL, ph);         // down_2: A = r2 (TMEM); K block 0 = columns [0,64) is ready first
      for (int j = 0; j < 2; ++j) {
        if (j == 1) wait_e(R2_FULL, ph);
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H2, 64>(tm + C_D2, tm + C_SA + 64 * j, tm + C_SA + 128 + 64 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D2_FULL);
      for (int j = 0; j < 4; ++j) {  // res_2 right behind down_2: A = r2, its own accumulator D3R
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_D3R, tm + C_SA + 32 * j, tm + C_SA + 128 + 32 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D3B_FULL);
      wait_e(R3_FULL, ph);          // up_2: A = r3 in the two shared-memory chunk buffers
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ss<H1, 32>(tm + C_D3, chunk_s + j * CHUNK_BYTES, CHUNK_HALF, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D3A_FULL);
      wait_e(O2_FULL, ph);          // up_1: A = o2
      for (int j = 0; j < 8; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H0, 16>(tm + C_D4, tm + C_SA + 16 * j, tm + C_SA + 128 + 16 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(D4A_FULL);
      up0_layer();                  // W_u0 y1 into [0,NY): the y1 chunks only exist once E has seen D4A_FULL
      signal(Y0_FULL);
      // ================= backward
      wait_e(DY0_FULL, ph);         // d_o1 = d_y0 W_u0 into [0,256)
      small_k(C_SA, true);
      signal(BD0_FULL);
      chunk_layer_128(C_DO2, false);  // d_o2 = d_y1 W_u1
      signal(BDO2_FULL);
      wait_e(BO2H_FULL, ph);         // d_r2 = d_o2 W_r2
      for (int j = 0; j < 4; ++j) {
        if (j == 2) wait_e(BO2_FULL, ph);
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_DR2, tm + C_SA + 32 * j, tm + C_SA + 128 + 32 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(BR2A_FULL);
      wait_e(BY2H_FULL, ph);         // d_r3 = d_y2 W_u2  (N = 64)
      for (int j = 0; j < 2; ++j) {
        if (j == 1) wait_e(BY2_FULL, ph);
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H2, 64>(tm + C_DR3, tm + C_SA + 64 * j, tm + C_SA + 128 + 64 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      signal(BD3_FULL);
      wait_e(BZ3_FULL, ph);          // d_r2 += d_z3 W_d2  (K = 64: A hi [0,64), lo [64,128))
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_DR2, tm + C_SA + 32 * j, tm + C_SA + 64 + 32 * j, wb, false);
        __syncwarp();
        release_w();
      }
      signal(BR2B_FULL);
      wait_e(BZ2_FULL, ph);          // d_r1 = d_z2 W_d1  (N = 256)
      for (int j = 0; j < 8; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H0, 16>(tm + C_DR1, tm + C_SA + 16 * j, tm + C_SA + 128 + 16 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      small_k(C_DR1, false);         // d_r1 += d_y0 Wc  (A = the d_y0 operand still in shared memory)
      signal(BD1B_FULL);
    }
  } else {
    // =================================================================== P: weight tape producer (forward then backward tape)
    if (elect_one()) {
      const int n_fwd = fwd_slots(d);
      const uint32_t per_tile = NSF + NSB;
      const uint32_t total = (uint32_t)my_tiles * per_tile;
      uint32_t in_tile = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % NSTAGE;
        mbar_wait(&bars[W_EMPTY + s], ((i / NSTAGE) & 1) ^ 1);
        const bool bwd = in_tile >= (uint32_t)NSF;
        const int slot_in = bwd ? (int)in_tile - NSF : (int)in_tile;
        const int slot = slot_in + (bwd ? n_fwd : 0);
        const uint32_t bytes = bwd ? bwd_slot_bytes(d, slot_in) : fwd_slot_bytes(d, slot_in);
        mbar_expect_tx(&bars[W_FULL + s], bytes);
        bulk_g2s(smem + SM_RING + s * SLOT_BYTES, tape + (size_t)slot * SLOT_BYTES, bytes, &bars[W_FULL + s]);
        if (++in_tile == per_tile) in_tile = 0;
      }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------- host side
constexpr int SUB_TILES_CAP = 8192;  // upper bound of the tiles per K3a/K3b launch pair (scratch = 1.1 MB per tile)
// whole waves of persistent CTAs per launch pair: the largest multiple of the SM count below the cap
static int sub_tiles_max() { return (SUB_TILES_CAP / sm_count()) * sm_count(); }

bool loss_tc_supported(const socm_unet* net) { return is_default_arch(net) && kin_of(net->d) <= MAX_KIN; }

static int64_t tape_bytes(int d) { return ((tc_workspace_bytes(d, true) + 1023) / 1024) * 1024; }
constexpr int64_t AUX_BYTES = ((AUX_FLOATS * 4 + 1023) / 1024) * 1024;

// gradients of the folded pair (unet_tc.cuh) from S = d_y0^T r1 and sb = sum_p d_y0:
//   dW_r1 += W_u0^T S,  db_r1 += W_u0^T sb,  dW_u0 += S W_r1^T + sb b_r1^T      (dW_u0 already holds d_y0^T y1)
__global__ void fold_finish_kernel(socm_unet net, const float* __restrict__ aux, float* __restrict__ grad) {
  const int d = net.d;
  const GradOffTc go = grad_offsets_tc(d);
  const float* S = aux + AUX_S;
  const float* sb = aux + AUX_SB;
  const int f = blockIdx.x, g = threadIdx.x;  // 256 x 256
  float acc = 0.f;
  for (int j = 0; j < d; ++j) acc = fmaf(net.w[8][(size_t)j * H0 + f], S[j * H0 + g], acc);
  grad[go.w[4] + (size_t)f * H0 + g] += acc;
  if (g == 0) {
    float b = 0.f;
    for (int j = 0; j < d; ++j) b = fmaf(net.w[8][(size_t)j * H0 + f], sb[j], b);
    grad[go.b[4] + f] += b;
  }
  if (g < d) {  // dW_u0[j = g][f]
    const int j = g;
    float u = sb[j] * net.b[4][f];
    for (int q = 0; q < H0; ++q) u = fmaf(S[j * H0 + q], net.w[4][(size_t)f * H0 + q], u);
    grad[go.w[8] + (size_t)j * H0 + f] += u;
  }
}

#ifdef SOCM_TC_PROF
extern "C" int socm_debug_k3_prof(unsigned long long* out192) {
  return (int)cudaMemcpyFromSymbol(out192, g_k3_prof, sizeof(g_k3_prof));
}
#endif

int64_t loss_tc_workspace_bytes(int d, int B, int K) {
  const int64_t n_tiles = (int64_t)(K + 1) * ((B + TP - 1) / TP);
  const int64_t sub = n_tiles < SUB_TILES_CAP ? n_tiles : SUB_TILES_CAP;
  return tape_bytes(d) + AUX_BYTES + sub * TILE_BYTES + 1024;
}

int launch_loss_tc(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream) {
  const int d = a.st.d;
  SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_loss_workspace_bytes)");
  unsigned char* tape = static_cast<unsigned char*>(workspace);
  float* small = tc_small_ptr(tape, d, true);
  float* aux = reinterpret_cast<float*>(tape + tape_bytes(d));
  unsigned char* scratch = tape + tape_bytes(d) + AUX_BYTES;
  scratch += (1024 - (reinterpret_cast<uintptr_t>(scratch) & 1023)) & 1023;
  if (int rc = pack_tc(net, tape, true, stream)) return rc;
  SOCM_CUDA(cudaMemsetAsync(aux, 0, AUX_FLOATS * sizeof(float), stream));
  const int smem = loss_tc_smem_bytes(d);
  const int n_mblk = (a.B + TP - 1) / TP;
  const int64_t n_tiles = (int64_t)(a.K + 1) * n_mblk;
  const int kin = kin_of(d);
  const int sub_max = sub_tiles_max();
  for (int64_t t0 = 0; t0 < n_tiles; t0 += sub_max) {
    const int nt = (int)((n_tiles - t0) < sub_max ? (n_tiles - t0) : sub_max);
    const int grid = nt < sm_count() ? nt : sm_count();
#define SOCM_LAUNCH_K3(KIN)                                                                               \
  do {                                                                                                    \
    SOCM_CUDA(cudaFuncSetAttribute(loss_tc_kernel<KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    loss_tc_kernel<KIN><<<grid, k3::NT, smem, stream>>>(a, tape, small, scratch, (int)t0, nt);            \
  } while (0)
    if (kin == 8) SOCM_LAUNCH_K3(8);
    else if (kin == 16) SOCM_LAUNCH_K3(16);
    else SOCM_LAUNCH_K3(24);
#undef SOCM_LAUNCH_K3
    SOCM_LAUNCH_CHECK();
    if (int rc = launch_wgrad_tc(scratch, nt, d, grad, aux, stream)) return rc;
  }
  fold_finish_kernel<<<H0, H0, 0, stream>>>(*net, aux, grad);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace tc
}  // namespace socm
