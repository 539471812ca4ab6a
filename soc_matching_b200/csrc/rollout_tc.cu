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
#include "kernels.h"
#include "rollout_common.cuh"
#include "unet_tc.cuh"

namespace socm {
namespace tc {

using namespace umma;

// ---------------------------------------------------------------- weight repack (once per call)
__global__ void pack_tc_kernel(socm_unet net, unsigned char* __restrict__ tape, float* __restrict__ small) {
  const int d = net.d, kin = kin_of(d), dp = ((d + 3) / 4) * 4;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int n_slots = fwd_slots(d);
  for (int s = 0; s < n_slots; ++s) {
    const SlotDesc sd = fwd_slot(d, s);
    const float* W = net.w[sd.layer];
    unsigned char* base = tape + (size_t)s * SLOT_BYTES;
    const int slab = sd.N * sd.Kc * 4;
    for (int i = tid; i < sd.N * sd.Kc; i += nth) {
      const int n = i / sd.Kc, k = i - n * sd.Kc;
      const float w = (sd.k0 + k < sd.ktot) ? W[(size_t)(sd.n0 + n) * sd.ktot + sd.k0 + k] : 0.f;
      const float hi = tf32_rn(w);
      const int off = wslab_off(n, k, sd.Kc);
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
  copy(so.b_r1, net.b[4], H0);
  for (int i = tid; i < H0 * dp; i += nth) {
    const int f = i / dp, j = i - f * dp;
    small[so.u0t + i] = j < d ? net.w[8][(size_t)j * H0 + f] : 0.f;
  }
  for (int i = tid; i < dp; i += nth) {
    small[so.b_u0 + i] = i < d ? net.b[8][i] : 0.f;
    small[so.b_r0 + i] = i < d ? net.b[3][i] : 0.f;
  }
  for (int i = tid; i < d * kin; i += nth) {
    const int j = i / kin, k = i - j * kin;
    small[so.r0 + i] = k <= d ? net.w[3][(size_t)j * (d + 1) + k] : 0.f;
  }
}

// ---------------------------------------------------------------- shared memory / barriers
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * SLOT_BYTES;
constexpr int SM_XIN = SM_CHUNK + 2 * CHUNK_BYTES;
constexpr int SM_SMALL = SM_XIN + MAX_KIN * 1024;  // xin: hi + lo, 128 x kin floats each
enum Bar {
  W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2,
  XIN_FULL = CH_EMPTY + 2, D0_FULL, D1_FULL, R2_FULL, D2_FULL, R3_FULL, D3A_FULL, Y2_FULL, D3B_FULL, O2_FULL,
  D4A_FULL, D0B_FULL, Y1_FULL, D4B_FULL, N_BARS
};
__host__ __device__ inline int rollout_tc_smem_bytes(int d) { return SM_SMALL + small_tc(d).total * 4 + N_BARS * 8 + 16; }

constexpr int NT_TC = 192;
// TMEM columns
constexpr uint32_t C_SA = 0, C_D1 = 256, C_D2 = 256, C_D3 = 256, C_R3 = 384, C_D4 = 256;

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
  constexpr int S0 = KIN > 16 ? 2 : 1;        // down_0 slots
  constexpr int NS = 40 + 2 * S0;             // weight stages per step

  // ---- one-time setup
  for (int i = tid; i < so.total; i += NT_TC) sm_small[i] = __ldg(small_g + i);
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bars[W_FULL + s], 1);
      mbar_init(&bars[W_EMPTY + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[CH_FULL + b], TP);
      mbar_init(&bars[CH_EMPTY + b], 1);
    }
    const int e2m[] = {XIN_FULL, R2_FULL, R3_FULL, Y2_FULL, O2_FULL, Y1_FULL};
    for (int i = 0; i < 6; ++i) mbar_init(&bars[e2m[i]], TP);
    const int m2e[] = {D0_FULL, D1_FULL, D2_FULL, D3A_FULL, D3B_FULL, D4A_FULL, D0B_FULL, D4B_FULL};
    for (int i = 0; i < 8; ++i) mbar_init(&bars[m2e[i]], 1);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK), xin_s = smem_addr(smem + SM_XIN);

  if (warp < 4) {
    // =================================================================== E: epilogue / path threads
    const int e = tid;
    const uint32_t lane_t = tm + ((uint32_t)(warp * 32) << 16);  // this warp's TMEM lane quarter
    uint32_t g = 0;   // steps done (all tiles)
    uint32_t cu = 0;  // chunks produced
    float* xin_hi = reinterpret_cast<float*>(smem + SM_XIN);
    float* xin_lo = reinterpret_cast<float*>(smem + SM_XIN + KIN * 512);

    auto gen_chunks = [&]() {  // r1 = relu(D0 + b_d0) -> 8 shared-memory A chunks
      for (int c = 0; c < 8; ++c) {
        const int b = cu & 1;
        mbar_wait(&bars[CH_EMPTY + b], ((cu >> 1) & 1) ^ 1);
        float v[32];
        tmem_ld32(lane_t + C_SA + 32 * c, reinterpret_cast<uint32_t*>(v));
        tmem_wait_ld();
        bias_relu32(v, sm_small + so.b_d0 + 32 * c);
        store_chunk32(smem + SM_CHUNK + b * CHUNK_BYTES, e, v);
        fence_async_smem();
        mbar_arrive(&bars[CH_FULL + b]);
        ++cu;
      }
    };

    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + e;
      const bool live = m < B;
      float x[KIN];  // x[j], j < d
      PathAcc acc{1.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < KIN; ++j) x[j] = (j < d && live) ? __ldg(a.x0 + (size_t)m * d + j) : 0.f;
      if (live) {
        if (a.states)
          for (int j = 0; j < d; ++j) a.states[(size_t)m * d + j] = x[j];
        if (a.stop) a.stop[m] = 1.f;
      }
      for (int k = 0; k < K; ++k, ++g) {
        const uint32_t ph = g & 1;
        // ---- E0: input operand [t, x, 0..] (hi / lo), act_off layout
        {
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
            float4 h, l;
            h.x = tf32_rn(v[0]); h.y = tf32_rn(v[1]); h.z = tf32_rn(v[2]); h.w = tf32_rn(v[3]);
            l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
            const int off = (e % 8) * 4 + (e / 8) * 32 + c4 * 512;  // floats
            *reinterpret_cast<float4*>(xin_hi + off) = h;
            *reinterpret_cast<float4*>(xin_lo + off) = l;
          }
          fence_async_smem();
          mbar_arrive(&bars[XIN_FULL]);
        }
        // noise of this step (overlaps the first MMAs)
        float eps[kMaxDim];
        if (live) draw_noise(a, m, k, eps);
        // ---- E1: r1 chunks for down_1
        mbar_wait(&bars[D0_FULL], ph);
        fence_after_sync();
        gen_chunks();
        // ---- E2: r2 = relu(D1 + b) -> TMEM A operand [0,128) hi, [128,256) lo
        mbar_wait(&bars[D1_FULL], ph);
        fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D1 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d1 + 32 * cb);
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[R2_FULL]);
        // ---- E3: r3 = relu(D2 + b) -> [384,448) hi, [448,512) lo
        mbar_wait(&bars[D2_FULL], ph);
        fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D2 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_d2 + 32 * cb);
          store_split32(lane_t + C_R3 + 32 * cb, lane_t + C_R3 + 64 + 32 * cb, v);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[R3_FULL]);
        // ---- E4: y2 = relu(D3 + b_u2) in place
        mbar_wait(&bars[D3A_FULL], ph);
        fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D3 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_u2 + 32 * cb);
          tmem_st32(lane_t + C_D3 + 32 * cb, reinterpret_cast<const uint32_t*>(v));
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[Y2_FULL]);
        // ---- E5: o2 = D3 + b_r2 -> TMEM A operand [0,128) hi, [128,256) lo
        mbar_wait(&bars[D3B_FULL], ph);
        fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D3 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias32(v, sm_small + so.b_r2 + 32 * cb);
          store_split32(lane_t + C_SA + 32 * cb, lane_t + C_SA + 128 + 32 * cb, v);
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[O2_FULL]);
        // ---- E6: y1 = relu(D4 + b_u1) in place
        mbar_wait(&bars[D4A_FULL], ph);
        fence_after_sync();
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D4 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias_relu32(v, sm_small + so.b_u1 + 32 * cb);
          tmem_st32(lane_t + C_D4 + 32 * cb, reinterpret_cast<const uint32_t*>(v));
        }
        tmem_wait_st();
        fence_before_sync();
        mbar_arrive(&bars[Y1_FULL]);
        // ---- E7: r1 chunks again for res_1
        mbar_wait(&bars[D0B_FULL], ph);
        fence_after_sync();
        gen_chunks();
        // ---- E8: o1 = D4 + b_r1; nabla_V = relu(up_0 o1 + b) + res_0 [t,x] + b; SDE step
        mbar_wait(&bars[D4B_FULL], ph);
        fence_after_sync();
        float au[KIN];
#pragma unroll
        for (int j = 0; j < KIN; ++j) au[j] = 0.f;
        const int dp = ((d + 3) / 4) * 4;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
          float v[32];
          tmem_ld32(lane_t + C_D4 + 32 * cb, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          bias32(v, sm_small + so.b_r1 + 32 * cb);
          const float* wu = sm_small + so.u0t + (32 * cb) * dp;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
#pragma unroll
            for (int q = 0; q < KIN / 4; ++q) {
              if (4 * q < d) {
                const float4 w = *reinterpret_cast<const float4*>(wu + j * dp + 4 * q);
                au[4 * q] = fmaf(w.x, v[j], au[4 * q]);
                au[4 * q + 1] = fmaf(w.y, v[j], au[4 * q + 1]);
                au[4 * q + 2] = fmaf(w.z, v[j], au[4 * q + 2]);
                au[4 * q + 3] = fmaf(w.w, v[j], au[4 * q + 3]);
              }
            }
          }
        }
        fence_before_sync();  // orders the TMEM reads before the next step's MMAs (via XIN_FULL)
        if (live) {
          float gv[kMaxDim], xs[kMaxDim];
          const float tk = __ldg(a.step_tab + 4 * K + k);
#pragma unroll
          for (int j = 0; j < KIN; ++j) {
            if (j < d) {
              const float* wr = sm_small + so.r0 + j * KIN;
              float ar = sm_small[so.b_r0 + j] + wr[0] * tk;
#pragma unroll
              for (int c = 1; c < KIN; ++c)
                if (c <= d) ar = fmaf(wr[c], x[c - 1], ar);
              gv[j] = fmaxf(au[j] + sm_small[so.b_u0 + j], 0.f) + ar;
              xs[j] = x[j];
            }
          }
          path_step_eps(a, m, k, xs, 1, gv, 1, eps, acc);
#pragma unroll
          for (int j = 0; j < KIN; ++j)
            if (j < d) x[j] = xs[j];
        }
      }
      if (live) {
        float xs[kMaxDim];
#pragma unroll
        for (int j = 0; j < KIN; ++j)
          if (j < d) xs[j] = x[j];
        path_finish(a, m, xs, 1, acc);
      }
    }
  } else if (warp == 4) {
    // =================================================================== M: MMA issue
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
      for (int h = 0; h < S0; ++h) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ss<H0 / S0, KIN>(tm + C_SA + h * (H0 / S0), xin_s, KIN * 512, wb, true);
        __syncwarp();
        release_w();
      }
    };
    for (; g < total_steps; ++g) {
      const uint32_t ph = g & 1;
      // ---- M0: down_0
      mbar_wait(&bars[XIN_FULL], ph);
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
          issue_block_ss<H1, 32>(tm + C_D1, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, wb, c == 0);
          commit(&bars[CH_EMPTY + b]);
        }
        __syncwarp();
        release_w();
        ++cm;
      }
      if (elect_one()) commit(&bars[D1_FULL]);
      __syncwarp();
      // ---- M2: down_2, A = r2 (TMEM), 2 blocks of K = 64
      mbar_wait(&bars[R2_FULL], ph);
      fence_after_sync();
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H2, 64>(tm + C_D2, tm + C_SA + 64 * j, tm + C_SA + 128 + 64 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D2_FULL]);
      __syncwarp();
      // ---- M3: up_2, A = r3 (TMEM), 2 blocks of K = 32
      mbar_wait(&bars[R3_FULL], ph);
      fence_after_sync();
      for (int j = 0; j < 2; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_D3, tm + C_R3 + 32 * j, tm + C_R3 + 64 + 32 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D3A_FULL]);
      __syncwarp();
      // ---- M4: res_2 on top of relu(y2), A = r2 (TMEM), 4 blocks of K = 32
      mbar_wait(&bars[Y2_FULL], ph);
      fence_after_sync();
      for (int j = 0; j < 4; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H1, 32>(tm + C_D3, tm + C_SA + 32 * j, tm + C_SA + 128 + 32 * j, wb, false);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D3B_FULL]);
      __syncwarp();
      // ---- M5: up_1, A = o2 (TMEM), 8 blocks of K = 16
      mbar_wait(&bars[O2_FULL], ph);
      fence_after_sync();
      for (int j = 0; j < 8; ++j) {
        const uint32_t wb = wait_w();
        if (elect_one()) issue_block_ts<H0, 16>(tm + C_D4, tm + C_SA + 16 * j, tm + C_SA + 128 + 16 * j, wb, j == 0);
        __syncwarp();
        release_w();
      }
      if (elect_one()) commit(&bars[D4A_FULL]);
      __syncwarp();
      // down_0 again into [0,256) once up_1 has finished reading o2 from there
      mbar_wait(&bars[D4A_FULL], ph);
      fence_after_sync();
      down0();
      if (elect_one()) commit(&bars[D0B_FULL]);
      __syncwarp();
      // ---- M6: res_1 on top of relu(y1), A = r1 chunks, 16 blocks of K = 16
      mbar_wait(&bars[Y1_FULL], ph);
      fence_after_sync();
      for (int c = 0; c < 8; ++c) {
        const uint32_t b = cm & 1;
        mbar_wait(&bars[CH_FULL + b], (cm >> 1) & 1);
        fence_after_sync();
        for (int j = 0; j < 2; ++j) {
          const uint32_t wb = wait_w();
          if (elect_one()) {
            issue_block_ss<H0, 16>(tm + C_D4, chunk_s + b * CHUNK_BYTES + j * 2 * ACT_KSTEP, CHUNK_HALF, wb, false);
            if (j == 1) commit(&bars[CH_EMPTY + b]);
          }
          __syncwarp();
          release_w();
        }
        ++cm;
      }
      if (elect_one()) commit(&bars[D4B_FULL]);
      __syncwarp();
    }
  } else {
    // =================================================================== P: weight tape producer
    if (elect_one()) {
      const uint32_t total = (uint32_t)my_tiles * (uint32_t)K * NS;
      uint32_t in_step = 0;
      for (uint32_t i = 0; i < total; ++i) {
        const uint32_t s = i % NSTAGE;
        mbar_wait(&bars[W_EMPTY + s], ((i / NSTAGE) & 1) ^ 1);
        const int slot = fwd_stage_slot(d, (int)in_step);
        const uint32_t bytes = slot < S0 ? (uint32_t)(2 * (H0 / S0) * KIN * 4) : (uint32_t)SLOT_BYTES;
        mbar_expect_tx(&bars[W_FULL + s], bytes);
        bulk_g2s(smem + SM_RING + s * SLOT_BYTES, tape + (size_t)slot * SLOT_BYTES, bytes, &bars[W_FULL + s]);
        if (++in_step == NS) in_step = 0;
      }
    }
    __syncwarp();
  }

  // ---- teardown: every MMA has completed (E waited for the last D4B_FULL)
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int64_t rollout_tc_workspace_bytes(int d) { return tc_workspace_bytes(d); }
bool rollout_tc_supported(const socm_unet* net) { return is_default_arch(net) && kin_of(net->d) <= MAX_KIN; }

int launch_rollout_tc(const RolloutArgs& a, const socm_unet* net, void* workspace, cudaStream_t stream) {
  const int d = a.st.d;
  unsigned char* tape = static_cast<unsigned char*>(workspace);
  float* small = reinterpret_cast<float*>(tape + (size_t)fwd_slots(d) * SLOT_BYTES);
  pack_tc_kernel<<<96, 256, 0, stream>>>(*net, tape, small);
  SOCM_LAUNCH_CHECK();
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
