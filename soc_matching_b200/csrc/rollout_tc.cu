// rollout_tc.cu -- K1 on the 5th-gen tensor cores: the Euler-Maruyama rollout of
// utils.stochastic_trajectories (utils.py:17-128) with the control network evaluated by
// tcgen05.mma (3xTF32, see unet_tc.cuh).  One persistent CTA per SM owns tiles of 128 paths for
// all K steps; the path state never leaves the thread that owns it.
//
// Warp roles (192 threads):
//   warps 0-3  "E": thread e <-> path e of the tile <-> TMEM lane e.  Builds the MMA operands
//              (input [t, x], ReLU + bias + hi/lo split of every activation), evaluates the last
//              layer and the SDE step (common.cuh: same code as the FFMA kernels).
//   warp 4     "M": issues every tcgen05.mma of the step (one elected lane), hands accumulators to
//              E with tcgen05.commit -> mbarrier.
//   warp 5     "P": streams the weight tape from L2 into the 3-stage ring (cp.async.bulk).
// Hand-offs are mbarriers; each of the per-step events below completes exactly once per step, so
// its wait parity is the step parity.
#include <type_traits>

#include "kernels.h"
#include "rollout_common.cuh"
#include "unet_tc.cuh"

namespace socm {
namespace tc {

using namespace umma;

// ---------------------------------------------------------------- weight repack (once per call)
// Wc = W_u0 W_r1 [d][256] (zero rows up to 32) and bc = W_u0 b_r1 + b_u0 (unet_tc.cuh, "folding"); fp64 accumulation
__global__ void fold_tc_kernel(socm_unet net, float* __restrict__ wc, float* __restrict__ small) {
  const int d = net.d, kin = kin_of(d);
  const int j = blockIdx.x, g = threadIdx.x;  // 32 blocks x 256 threads
  double acc = 0.0;
  if (j < d)
    for (int f = 0; f < H0; ++f) acc += (double)net.w[8][(size_t)j * H0 + f] * (double)net.w[4][(size_t)f * H0 + g];
  wc[j * H0 + g] = (float)acc;
  if (g == 0 && j < kin) {
    double b = 0.0;
    if (j < d) {
      b = (double)net.b[8][j];
      for (int f = 0; f < H0; ++f) b += (double)net.w[8][(size_t)j * H0 + f] * (double)net.b[4][f];
    }
    small[small_tc(d).bc + j] = (float)b;
  }
}

__global__ void pack_tc_kernel(socm_unet net, const float* __restrict__ wc, unsigned char* __restrict__ tape,
                               float* __restrict__ small, int with_bwd) {
  const int d = net.d, kin = kin_of(d);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int n_fi = fwd_items(d), n_items = n_fi + (with_bwd ? bwd_items(d) : 0);
  for (int it = 0; it < n_items; ++it) {
    const PackItem pi = it < n_fi ? fwd_item(d, it) : bwd_item(d, it - n_fi);
    const SlotDesc sd = pi.sd;
    const float* W = sd.layer == WC_LAYER ? wc : net.w[sd.layer];
    unsigned char* base = tape + (size_t)(pi.slot + (it < n_fi ? 0 : fwd_slots(d))) * SLOT_BYTES + pi.byte_off;
    const int slab = (sd.slab_n ? sd.slab_n : sd.N) * sd.Kc * 4;
    for (int i = tid; i < sd.N * sd.Kc; i += nth) {
      const int n = i / sd.Kc, k = i - n * sd.Kc;
      float w = 0.f;
      if (sd.k0 + k < sd.klim && sd.n0 + n < sd.nlim)
        w = sd.transposed ? W[(size_t)(sd.k0 + k) * sd.ktot + sd.n0 + n] : W[(size_t)(sd.n0 + n) * sd.ktot + sd.k0 + k];
      const float hi = tf32_rn(w);
      const int off = wslab_off(sd.slab_row + n, k, sd.Kc);
      *reinterpret_cast<float*>(base + off) = hi;
      *reinterpret_cast<float*>(base + slab + off) = w - hi;
    }
  }
  const SmallTc so = small_tc(d);
  auto copy = [&](int off, const float* src, int n) {
    for (int i = tid; i < n; i += nth) small[off + i] = src[i];
  };
  copy(so.b_d0, net.b[0], H0);
  copy(so.b_d1, net.b[1], H1);
  copy(so.b_d2, net.b[2], H2);
  copy(so.b_u2, net.b[6], H1);
  copy(so.b_r2, net.b[5], H1);
  copy(so.b_u1, net.b[7], H0);
  for (int i = tid; i < kin; i += nth) small[so.b_r0 + i] = i < d ? net.b[3][i] : 0.f;
  for (int i = tid; i < kin * kin; i += nth) {
    const int j = i / kin, k = i - j * kin;
    small[so.r0 + i] = (j < d && k <= d) ? net.w[3][(size_t)j * (d + 1) + k] : 0.f;
  }
}

// ---------------------------------------------------------------- shared memory / barriers
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * SLOT_BYTES;
constexpr int SM_XIN = SM_CHUNK + 2 * CHUNK_BYTES;
constexpr int SM_SMALL = SM_XIN + MAX_KIN * 1024;  // xin: hi + lo, 128 x kin floats each
enum Bar {
  W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2,
  XIN_FULL = CH_EMPTY + 2, D0_FULL, D1_FULL, R2_FULL, D2_FULL, R3_FULL, D3A_FULL, D3B_FULL, O2_FULL,
  D4A_FULL, Y0_FULL,
  R2H_FULL,  // first column half of r2 is ready: down_2 starts on its K block 0 while the second half is written
  N_BARS
};
__host__ __device__ inline int rollout_tc_smem_bytes(int d) { return SM_SMALL + small_tc(d).total * 4 + N_BARS * 8 + 16; }

constexpr int NT_TC = 320;  // 8 epilogue warps + MMA warp + producer warp
constexpr int NE = 256;     // epilogue threads

#ifdef SOCM_TC_PROF
// per-phase cycle accumulators of block 0 (E thread 0: slots 0-15, M warp: slots 16-31); debug builds only
__device__ unsigned long long g_tc_prof[48];
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[16] = {0}
#define PROF_MARK(i) do { const long long t_ = clock64(); prof_acc[i] += (unsigned long long)(t_ - prof_t); prof_t = t_; } while (0)
#define PROF_FLUSH(base, cond) do { if (blockIdx.x == 0 && (cond)) for (int i_ = 0; i_ < 16; ++i_) g_tc_prof[(base) + i_] = prof_acc[i_]; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_FLUSH(base, cond)
#endif
// TMEM columns
constexpr uint32_t C_SA = 0, C_D1 = 256, C_D2 = 256, C_D3 = 256, C_D3R = 384, C_D4 = 256;
// folded last layer (unet_tc.cuh): C_Y0P = Wc r1, accumulated next to down_1 and read out before r3 is
// written there; C_Y0 = W_u0 y1, accumulated over the y1 chunks once up_1 has consumed o2.
constexpr uint32_t C_Y0P = 384, C_Y0 = 0;
static_assert(C_Y0P == C_D1 + H1, "Wc r1 must sit right behind the down_1 accumulator (joint MMA)");


// eps[0..d) for (path m, step k) into registers: injected or Philox (same draws as draw_noise)
template <int KIN>
__device__ __forceinline__ void draw_noise_reg(const RolloutArgs& a, int m, int k, float* eps) {
  const int d = a.st.d;
  if (a.noise_in != nullptr) {
    const float* src = a.noise_in + ((size_t)k * a.B + m) * d;
#pragma unroll
    for (int j = 0; j < KIN; ++j)
      if (j < d) eps[j] = __ldg(src + j);
  } else {
#pragma unroll
    for (int blk = 0; blk < KIN / 4; ++blk) {
      if (blk * 4 < d) {
        float z[4];
        philox_normal4(a.seed, a.path_offset + (uint64_t)m, (uint32_t)k, (uint32_t)blk, z);
#pragma unroll
        for (int j = 0; j < 4; ++j) eps[blk * 4 + j] = (blk * 4 + j < d) ? z[j] : 0.f;
      }
    }
  }
}

template <int KIN>
__global__ void __launch_bounds__(NT_TC, 1) rollout_tc_kernel(RolloutArgs a, const unsigned char* __restrict__ tape,
                                                              const float* __restrict__ small_g) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int d = a.st.d, K = a.K, B = a.B;
  const SmallTc so = small_tc(d);
  float* sm_small = reinterpret_cast<float*>(smem + SM_SMALL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_SMALL + so.total * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const bool diag = diag_fast_path(a.st, a.warmA != nullptr);
  constexpr int S0 = KIN > 16 ? 2 : 1;        // down_0 slots
  constexpr int NY = KIN <= 16 ? 16 : 32;     // N of the folded up_0 MMAs
  constexpr int NU0 = NY == 16 ? 1 : 2;       // up_0 slots
  constexpr int BPS = 8 / NU0;                // up_0 K-chunks per slot
  constexpr int NS = S0 + 24 + NU0;           // weight stages per step (= forward tape slots, in order)

  // ---- one-time setup
  for (int i = tid; i < so.total; i += NT_TC) sm_small[i] = __ldg(small_g + i);
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
    const int e2m[] = {R2_FULL, R3_FULL, O2_FULL, R2H_FULL};
    for (int i = 0; i < 4; ++i) mbar_init(&bars[e2m[i]], NE / 32);
    const int m2e[] = {D0_FULL, D1_FULL, D2_FULL, D3A_FULL, D3B_FULL, D4A_FULL, Y0_FULL};
    for (int i = 0; i < 7; ++i) mbar_init(&bars[m2e[i]], 1);
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK), xin_s = smem_addr(smem + SM_XIN);

  if (warp < 8) {
    // =================================================================== E: epilogue / path threads
    // thread -> TMEM lane p (= path of the tile) and column half h: the two warps that share a lane
    // quarter split every accumulator's columns.  h == 0 threads own the path state.
    // The two halves run separately compiled copies of the program (h is a compile-time constant)
    // so that owner-only and helper-only values do not add up in the register allocation.
    auto e_program = [&](auto h_const) {
    constexpr int h = decltype(h_const)::value;
    const int p = tid & (TP - 1);
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t g = 0;   // steps done (all tiles)
    uint32_t cu = 0;  // chunks produced
    PROF_DECL;
    float* xin_hi = reinterpret_cast<float*>(smem + SM_XIN);
    float* xin_lo = reinterpret_cast<float*>(smem + SM_XIN + KIN * 512);
    float* stage_f = reinterpret_cast<float*>(smem + SM_CHUNK);  // exchange area (chunk buffer 0, idle after up_0)

    float kap[KIN];  // kappa padded with zeros (diagonal fast path of the SDE step)
#pragma unroll
    for (int j = 0; j < KIN; ++j) kap[j] = (h == 0 && diag && j < d) ? __ldg(a.st.kappa + j) : 0.f;

    // relu(acc + bias) of a 256-wide accumulator -> 8 shared-memory A chunks, 16 features per half.
    // The TMEM load of chunk c+1 is in flight while chunk c is processed (a tcgen05.ld that competes with
    // running MMAs takes ~450 cycles) and the chunk buffer is waited for last.
    auto gen_chunks = [&](uint32_t src_col, int bias_off) {
      auto body = [&](int c, float* v) {
        bias_relu16(v, sm_small + bias_off + 32 * c + 16 * h);
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

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + p;
      const bool live = m < B;
      float x[KIN];  // owner threads: x[j], j < d
      PathAcc acc{1.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < KIN; ++j) x[j] = (j < d && live && h == 0) ? __ldg(a.x0 + (size_t)m * d + j) : 0.f;
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
        // ---- E0 (owners): input operand [t, x, 0..] (hi / lo), act_off layout
        if (h == 0) {
          const float tk = __ldg(a.step_tab + 4 * K + k);
#pragma unroll
          for (int c4 = 0; c4 < KIN / 4; ++c4) {
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int c = 4 * c4 + q;
              v[q] = c == 0 ? tk : x[c - 1 < 0 ? 0 : c - 1];
              if (c > d) v[q] = 0.f;
            }
            float4 hh, ll;
            hh.x = tf32_rn(v[0]); hh.y = tf32_rn(v[1]); hh.z = tf32_rn(v[2]); hh.w = tf32_rn(v[3]);
            ll.x = v[0] - hh.x; ll.y = v[1] - hh.y; ll.z = v[2] - hh.z; ll.w = v[3] - hh.w;
            const int off = (p % 8) * 4 + (p / 8) * 32 + c4 * 512;  // floats
            *reinterpret_cast<float4*>(xin_hi + off) = hh;
            *reinterpret_cast<float4*>(xin_lo + off) = ll;
          }
          fence_async_smem();
          warp_arrive(&bars[XIN_FULL]);
        }
        PROF_MARK(0);
        // ---- E1: r1 chunks for down_1
        mbar_wait(&bars[D0_FULL], ph);
        PROF_MARK(2);
        fence_after_sync();
        gen_chunks(C_SA, so.b_d0);
        PROF_MARK(3);
        // ---- E2: r2 = relu(D1 + b) -> TMEM A operand [0,128) hi, [128,256) lo   (64 columns per half)
        mbar_wait(&bars[D1_FULL], ph);
        PROF_MARK(4);
        fence_after_sync();
        float au[KIN];  // owners: (Wc r1)[j], kept in registers until the end of the step
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
          const int cb = 2 * i + h;  // iteration 0 covers columns [0,64) = K block 0 of down_2
          float v[32];
          tmem_ld32(lane_t + C_D1 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d1 + 32 * cb);
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
          if (i == 0) {
            tmem_wait_st();
            fence_before_sync();
            warp_arrive(&bars[R2H_FULL]);
          }
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[R2_FULL]);
        PROF_MARK(5);
        // ---- E3: r3 = relu(D2 + b) -> shared-memory A operand: features [32h, 32h+32) = chunk buffer h
        mbar_wait(&bars[D2_FULL], ph);
        PROF_MARK(6);
        fence_after_sync();
        {
          float v[32];
          tmem_ld32(lane_t + C_D2 + 32 * h, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d2 + 32 * h);
          store_chunk32(smem + SM_CHUNK + h * CHUNK_BYTES, p, v);
        }
        fence_before_sync();
        fence_async_smem();
        warp_arrive(&bars[R3_FULL]);
        PROF_MARK(5);
        // ---- E5: o2 = relu(D3 + b_u2) + D3R + b_r2 -> TMEM A operand [0,128) hi, [128,256) lo  (up_2 and res_2 done)
        mbar_wait(&bars[D3A_FULL], ph);
        mbar_wait(&bars[D3B_FULL], ph);
        PROF_MARK(7);
        fence_after_sync();
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
          const int c0 = 64 * (i >> 1) + 32 * h + 16 * (i & 1);  // iterations 0, 1 cover columns [0,64) of both halves
          float y[16], rr[16];
          tmem_ld16(lane_t + C_D3 + c0, reinterpret_cast<uint32_t*>(y));
          tmem_ld16(lane_t + C_D3R + c0, reinterpret_cast<uint32_t*>(rr));
          tmem_wait_ld();
          const float* bu = sm_small + so.b_u2 + c0;
          const float* br = sm_small + so.b_r2 + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j] + bu[j], 0.f) + (rr[j] + br[j]);
          store_split16(lane_t + C_SA + c0, lane_t + C_SA + 128 + c0, y);
        }
        tmem_wait_st();
        fence_before_sync();
        warp_arrive(&bars[O2_FULL]);
        PROF_MARK(5);
        // up_1 runs for ~6k cycles now: the helper half draws this step's noise meanwhile
        float eps[KIN];
#pragma unroll
        for (int j = 0; j < KIN; ++j) eps[j] = 0.f;
        if (h == 1 && live) draw_noise_reg<KIN>(a, m, k, eps);
        // ... and the owner half evaluates res_0 [t, x] (it depends on the state only), which takes the d x (d+1)
        // mat-vec out of the serial tail of the step
        float ar[KIN];
#pragma unroll
        for (int j = 0; j < KIN; ++j) ar[j] = 0.f;
        if (h == 0) {
          const float tk = __ldg(a.step_tab + 4 * K + k);
#pragma unroll
          for (int j = 0; j < KIN; ++j) {  // lanes >= d come out as exact zeros (zero-padded parameters)
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
        PROF_MARK(1);
        // ---- E6: y1 = relu(D4 + b_u1) -> 8 shared-memory A chunks for the folded up_0 (N = NY)
        mbar_wait(&bars[D4A_FULL], ph);
        PROF_MARK(8);
        fence_after_sync();
        gen_chunks(C_D4, so.b_u1);
        PROF_MARK(11);
        // ---- E8: y0 = W_u0 y1 (TMEM [0,NY)) + Wc r1 (registers) + bc
        mbar_wait(&bars[Y0_FULL], ph);
        PROF_MARK(12);
        fence_after_sync();
        if (h == 0) {
          float yp[NY];
          if constexpr (NY == 16) tmem_ld16(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
          else tmem_ld32(lane_t + C_Y0, reinterpret_cast<uint32_t*>(yp));
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < KIN; ++j) au[j] += yp[j];
        }
        fence_before_sync();  // orders the TMEM reads before the next step's MMAs (via XIN_FULL)
        PROF_MARK(13);
        // helpers hand the noise to the owners through the idle chunk buffer
        if (h == 1) {
#pragma unroll
          for (int j = 0; j < KIN; ++j) stage_f[j * TP + p] = eps[j];
        }
        e_sync();
        PROF_MARK(15);
        if (h == 0 && live) {
          float gv[KIN], u[KIN];
#pragma unroll
          for (int j = 0; j < KIN; ++j) {
            gv[j] = fmaxf(au[j] + sm_small[so.bc + j], 0.f) + ar[j];
            eps[j] = stage_f[j * TP + p];
          }
          const float dt = __ldg(a.step_tab + k), sq_ldt = __ldg(a.step_tab + K + k);
          const float dt_l = __ldg(a.step_tab + 2 * K + k), sq_dtl = __ldg(a.step_tab + 3 * K + k);
          const float* wA = a.warmA ? a.warmA + (size_t)k * d * d : nullptr;
          const float* wc = a.warmc ? a.warmc + (size_t)k * d : nullptr;
          float eff;
          if (diag) {
            eff = sde_step_diag<KIN>(a.st.kind == SOCM_MOLECULAR_DYNAMICS, a.st.lmbd, kap, x, gv, eps, u, dt, sq_ldt,
                                     dt_l, sq_dtl, acc);
          } else {
            // dense sigma / OU drift / warm start: run-time loops on thread-local arrays
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
        e_sync();  // the exchange area is chunk buffer 0 again from here on
        PROF_MARK(14);
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
    PROF_FLUSH(0, tid == 0);
    PROF_FLUSH(32, tid == 128);
    };  // e_program
    if (warp < 4) e_program(std::integral_constant<int, 0>{});
    else e_program(std::integral_constant<int, 1>{});
  } else if (warp == 8) {
    // =================================================================== M: MMA issue
    PROF_DECL;
    uint32_t ws = 0;  // weight stages consumed
    uint32_t cm = 0;  // chunks consumed
    uint32_t g = 0;
    const uint32_t total_steps = (uint32_t)my_tiles * (uint32_t)K;
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
    auto down0 = [&]() {  // D0 = [t,x] W0^T into columns [0,256)
#pragma unroll
      for (int hh = 0; hh < S0; ++hh) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ss<H0 / S0, KIN>(tm + C_SA + hh * (H0 / S0), xin_s, KIN * 512, wb, true);
        __syncwarp();
        release_w();
      }
    };
    for (; g < total_steps; ++g) {
      const uint32_t ph = g & 1;
      // ---- M0: down_0
      PROF_MARK(0);
      mbar_wait(&bars[XIN_FULL], ph);
      PROF_MARK(1);
      fence_after_sync();
      down0();
      if (elect_one()) commit(&bars[D0_FULL]);
      __syncwarp();
      // ---- M1: down_1, A = r1 chunks (shared memory), 8 blocks of K = 32
      for (int c = 0; c < 8; ++c) {
        const uint32_t b = cm & 1;
        mbar_wait(&bars[CH_FULL + b], (cm >> 1) & 1);
        fence_after_sync();
        const uint32_t wb = wait_w();
        if (elect_one()) {
          // down_1 and Wc r1 in one MMA: N = H1 + NY, columns [C_D1, C_D1 + H1) and [C_Y0P, C_Y0P + NY)
          issue_block_ss<H1 + NY, 32>(tm + C_D1, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, c == 0);
          commit(&bars[CH_EMPTY + b]);
        }
        __syncwarp();
        release_w();
        ++cm;
      }
      if (elect_one()) commit(&bars[D1_FULL]);
      __syncwarp();
      PROF_MARK(2);
      // ---- M2: down_2, A = r2 (TMEM), 2 blocks of K = 64
      mbar_wait(&bars[R2H_FULL], ph);
      PROF_MARK(3);
      fence_after_sync();
      for (int j = 0; j < 2; ++j) {
        if (j == 1) {
          mbar_wait(&bars[R2_FULL], ph);
          fence_after_sync();
        }
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H2, 64>(tm + C_D2, tm + C_SA + 64 * j, tm + C_SA + 128 + 64 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D2_FULL]);
      __syncwarp();
      PROF_MARK(4);
      // ---- M3: res_2 right behind down_2 (A = r2 in TMEM, own accumulator D3R), 4 blocks of K = 32
      for (int j = 0; j < 4; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_D3R, tm + C_SA + 32 * j, tm + C_SA + 128 + 32 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D3B_FULL]);
      __syncwarp();
      PROF_MARK(4);
      // ---- M4: up_2, A = r3 in the two shared-memory chunk buffers, 2 blocks of K = 32
      mbar_wait(&bars[R3_FULL], ph);
      PROF_MARK(5);
      fence_after_sync();
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ss<H1, 32>(tm + C_D3, chunk_s + j * CHUNK_BYTES, CHUNK_HALF, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D3A_FULL]);
      __syncwarp();
      PROF_MARK(4);
      // ---- M5: up_1, A = o2 (TMEM), 8 blocks of K = 16
      mbar_wait(&bars[O2_FULL], ph);
      PROF_MARK(6);
      fence_after_sync();
      for (int j = 0; j < 8; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H0, 16>(tm + C_D4, tm + C_SA + 16 * j, tm + C_SA + 128 + 16 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D4A_FULL]);
      __syncwarp();
      PROF_MARK(7);
      // ---- M6: folded up_0 = W_u0 y1 into [0,NY), A = y1 chunks (E produces them only after D4A_FULL,
      //      i.e. once up_1 has finished reading o2 from there), 8 blocks of K = 32
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
      if (elect_one()) commit(&bars[Y0_FULL]);
      __syncwarp();
      PROF_MARK(11);
    }
    PROF_FLUSH(16, (tid & 31) == 0);
  } else {
    // =================================================================== P: weight tape producer
    if (elect_one()) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)K * NS;
      uint32_t in_step = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % NSTAGE;
        mbar_wait(&bars[W_EMPTY + s], ((i / NSTAGE) & 1) ^ 1);
        const int slot = (int)in_step;
        const uint32_t bytes = fwd_slot_bytes(d, slot);
        mbar_expect_tx(&bars[W_FULL + s], bytes);
        bulk_g2s(smem + SM_RING + s * SLOT_BYTES, tape + (size_t)slot * SLOT_BYTES, bytes, &bars[W_FULL + s]);
        if (++in_step == NS) in_step = 0;
      }
    }
    __syncwarp();
  }

  // ---- teardown: every MMA has completed (E waited for the last Y0_FULL)
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

#ifdef SOCM_TC_PROF
extern "C" int socm_debug_tc_prof(unsigned long long* out32) {
  return (int)cudaMemcpyFromSymbol(out32, g_tc_prof, sizeof(g_tc_prof));
}
#endif

int pack_tc(const socm_unet* net, unsigned char* tape, bool with_bwd, cudaStream_t stream) {
  float* small = tc_small_ptr(tape, net->d, with_bwd);
  float* wc = tc_wc_ptr(tape, net->d, with_bwd);
  fold_tc_kernel<<<32, H0, 0, stream>>>(*net, wc, small);
  SOCM_LAUNCH_CHECK();
  pack_tc_kernel<<<96, 256, 0, stream>>>(*net, wc, tape, small, with_bwd ? 1 : 0);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

int64_t rollout_tc_workspace_bytes(int d) { return tc_workspace_bytes(d); }
bool rollout_tc_supported(const socm_unet* net) { return is_default_arch(net) && kin_of(net->d) <= MAX_KIN; }

int launch_rollout_tc(const RolloutArgs& a, const socm_unet* net, void* workspace, cudaStream_t stream) {
  const int d = a.st.d;
  unsigned char* tape = static_cast<unsigned char*>(workspace);
  float* small = tc_small_ptr(tape, d, false);
  if (int rc = pack_tc(net, tape, false, stream)) return rc;
  const int smem = rollout_tc_smem_bytes(d);
  const int n_tiles = (a.B + TP - 1) / TP;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  const int kin = kin_of(d);
#define SOCM_LAUNCH_TC(KIN)                                                                                   \
  do {                                                                                                        \
    SOCM_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<KIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    rollout_tc_kernel<KIN><<<grid, NT_TC, smem, stream>>>(a, tape, small);                                    \
  } while (0)
  if (kin == 8) SOCM_LAUNCH_TC(8);
  else if (kin == 16) SOCM_LAUNCH_TC(16);
  else SOCM_LAUNCH_TC(24);
#undef SOCM_LAUNCH_TC
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace tc
}  // namespace socm
