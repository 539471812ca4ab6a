// wgrad_h.cu -- K3b of the fp16-split engine: all weight / bias gradients of the default UNet as tcgen05 GEMMs with the
// trajectory points as the contraction index (kind::f16, three products per pair, fp32 accumulation in TMEM).
//
//   dW[out][in] += sum_p dY[p][out] * Act[p][in]
// The scratch written by loss_h.cu keeps the block structure of loss_tc.cuh (tile -> 4 quarters -> 59 feature blocks of
// 32 features x 4 KB), but a block holds the 32 points of the quarter as fp16 hi and lo planes of the OPERAND-SCALED value
// (the very hi / lo pairs K3a feeds its own MMAs) in the canonical no-swizzle MN-major layout: per group of 8 features eight
// 128-byte core matrices (8 points x 8 features), point groups 0..3 = hi, 4..7 = lo.  It is an fp16 operand of K = 64:
// one cp.async.bulk lands a run of feature blocks in shared memory ready for the MMA, and a product is
//     A_hi B_hi + A_lo B_hi + A_hi B_lo   =   k-steps (0,1)x(0,1), (2,3)x(0,1), (0,1)x(2,3)   of 16 points each.
// Compared with wgrad_tc.cu (fp32 scratch) there is no in-kernel hi / lo split -- that pass and its second copy of every
// stage made the old kernel latency-bound with two 114 KB stages -- so three 56 KB stages fit and the kernel streams the
// scratch from HBM.  The stage / product / output tables are shared (wgrad_tables.cuh).  The flush multiplies by the
// inverse operand scales (powers of two from the calibration, passed in `scales`).
//
// Warp roles (192 threads): warps 0-3 flush (TMEM -> red.global.add), warp 4 MMA issue, warp 5 producer.
#include <cstdlib>

#include "kernels.h"
#include "loss_tc.cuh"
#include "umma.cuh"
#include "unet_h.cuh"
#include "wgrad_tables.cuh"

namespace socm {
namespace hx {

using namespace umma;
using namespace tc;   // tables, FB_*, GradOffTc

constexpr int WH_STAGES = 3;
constexpr int WH_SMEM = WH_STAGES * WG_RAW_BYTES + 1024 + 256;
constexpr int WH_NT = 192;
constexpr int WH_PREFETCH = 3;   // stages of L2 prefetch ahead of the bulk copies; 6 and more thrash L2 (measured: 2-4 equal, 6 +9% time, 9 +40%)

// ---------------------------------------------------------------- work description (types of wgrad_tables.cuh)
// Three passes whose accumulators fit the 512 TMEM columns; compared with the tables of the fp32-scratch kernel, res_2
// shares its stage with down_2 (r2 is copied once and serves as the N operand of the one and the M operand of the other),
// and 64-row tensors used as M = 128 operands are copied alone: the other 64 rows are whatever follows them in the stage
// (finite fp16 values of a neighbouring tensor), and land in accumulator rows the flush ignores.  DRAM reads per
// (tile, quarter): 30 + 13 + 18 = 61 feature blocks for the 59 the scratch holds (XIN is read once per pass).
constexpr int H_N_STAGE = 6;
static __constant__ int h_pass_stage[N_PASS + 1] = {0, 3, 4, 6};
static __constant__ int h_pass_out[N_PASS + 1] = {0, 8, 12, 19};
static __constant__ StageDesc h_stage[H_N_STAGE] = {
    // ---- pass 0: down_1 (+bias), S^T = r1^T d_y0, (d_y0^T y1)^T, down_0
    {4, {{WG_A, FB_DZ2, 4 * FBB}, {WG_B, FB_R1, 8 * FBB}, {WG_X, FB_XIN, FBB}, {WG_X2, FB_DY0, FBB}},
     4, {{WG_A, WG_B, 256, 0}, {WG_A, WG_X, 32, 256}, {WG_B, WG_X2, 32, 288}, {WG_B + 16384, WG_X2, 32, 320}}},
    {2, {{WG_B, FB_Y1, 8 * FBB}, {WG_X2, FB_DY0, FBB}, {0, 0, 0}, {0, 0, 0}},
     2, {{WG_B, WG_X2, 32, 352}, {WG_B + 16384, WG_X2, 32, 384}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    {2, {{WG_B, FB_DZ1, 8 * FBB}, {WG_X, FB_XIN, FBB}, {0, 0, 0}, {0, 0, 0}},
     2, {{WG_B, WG_X, 32, 416}, {WG_B + 16384, WG_X, 32, 448}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    // ---- pass 1: up_1 (both row halves; [o2 | XIN] back to back as one N = 160 operand)
    {3, {{20480, FB_DY1, 8 * FBB}, {0, FB_O2, 4 * FBB}, {16384, FB_XIN, FBB}, {0, 0, 0}},
     2, {{20480, 0, 160, 0}, {20480 + 16384, 0, 160, 160}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    // ---- pass 2: up_2; then res_2, down_2 (D[in][out]) + its bias and the d-sized layers on one stage:
    //      d_y0 d_o0 @0 | d_o2 @8 KB | d_z3 @24 KB (d_o2 and d_z3 are neighbours in the scratch: one copy) | r2 @32 KB | XIN @48 KB
    {3, {{WG_A, FB_DY2, 4 * FBB}, {16384, FB_R3, 2 * FBB}, {24576, FB_XIN, FBB}, {0, 0, 0}},
     1, {{WG_A, 16384, 96, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    {4, {{0, FB_DY0, 2 * FBB}, {8192, FB_DO2, 6 * FBB}, {32768, FB_R2, 4 * FBB}, {49152, FB_XIN, FBB}},
     4, {{8192, 32768, 160, 224}, {32768, 24576, 64, 96}, {24576, 49152, 32, 160}, {0, 49152, 32, 192}}},
};
static_assert(FB_DO0 == FB_DY0 + 1 && FB_DZ3 == FB_DO2 + 4, "neighbouring tensors are copied together");
static __constant__ OutDesc h_out[N_OUT] = {
    {0, 256, OUT_DIRECT, 1, 0, 128},    {256, 32, OUT_BIAS, 1, 0, 128},        // down_1
    {288, 32, OUT_AUX_S, 8, 0, 128},    {320, 32, OUT_AUX_S, 8, 128, 128},     // S^T -> aux (res_1 / up_0 via fold_finish_kernel)
    {352, 32, OUT_TRANSPOSED, 8, 0, 128}, {384, 32, OUT_TRANSPOSED, 8, 128, 128},  // up_0, y1 part
    {416, 32, OUT_XIN, 0, 0, 128},      {448, 32, OUT_XIN, 0, 128, 128},       // down_0 (+ bias via the ones feature)
    {0, 128, OUT_DIRECT, 7, 0, 128},    {128, 32, OUT_BIAS, 7, 0, 128},        // up_1 rows 0..127
    {160, 128, OUT_DIRECT, 7, 128, 128}, {288, 32, OUT_BIAS, 7, 128, 128},     // up_1 rows 128..255
    {0, 64, OUT_DIRECT, 6, 0, 128},     {64, 32, OUT_BIAS, 6, 0, 128},         // up_2
    {96, 64, OUT_TRANSPOSED, 2, 0, 128},                                       // down_2
    {160, 32, OUT_BIAS, 2, 0, 64},                                             // bias of down_2 (rows = d_z3 features)
    {192, 32, OUT_SMALL, 3, 0, 64},     // rows 0..31 d_y0 -> b(up_0), aux sb; rows 32..63 d_o0 -> res_0, b(res_0)
    {224, 128, OUT_DIRECT, 5, 0, 128},  {352, 32, OUT_BIAS, 5, 0, 128},        // res_2
};

// MN-major, no swizzle: LBO = 128 (next 8 points), SBO = 1024 (next 8 features)
__device__ __forceinline__ uint64_t mn_desc_h(uint32_t saddr) { return smem_desc(saddr, 128, 1024); }
__device__ __forceinline__ void red_add_h(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// inverse scale of the product that output block `o` of the tables accumulates (index into h_out); `sk` = scratch tensor
// scales (ScratchT order).  Bias columns multiply a dY tensor with the constant-1 feature.
__device__ __forceinline__ void out_scales(int o, const float* __restrict__ sk, float& inv_main, float& inv_bias) {
  // dY tensor / activation tensor of each output block, in the order of c_out
  const int ta[N_OUT] = {T_DZ2, T_DZ2, T_R1, T_R1, T_Y1, T_Y1, T_DZ1, T_DZ1, T_DY1, T_DY1, T_DY1, T_DY1,
                         T_DY2, T_DY2, T_R2, T_DZ3, T_DY0, T_DO2, T_DO2};
  const int tb[N_OUT] = {T_R1, -1, T_DY0, T_DY0, T_DY0, T_DY0, T_XIN, T_XIN, T_O2, -1, T_O2, -1,
                         T_R3, -1, T_DZ3, -1, T_XIN, T_R2, -1};
  inv_bias = 1.f / sk[ta[o]];
  inv_main = tb[o] >= 0 ? 1.f / (sk[ta[o]] * sk[tb[o]]) : inv_bias;
}

__global__ void __launch_bounds__(WH_NT, 1) wgrad_h_kernel(const unsigned char* __restrict__ scratch, int n_tiles, int d,
                                                           const float* __restrict__ sk, float* __restrict__ grad,
                                                           float* __restrict__ aux, int prefetch_dist) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WH_STAGES * WG_RAW_BYTES);
  uint64_t* full = bars;                  // [WH_STAGES] bulk copies landed
  uint64_t* empty = bars + WH_STAGES;     // [WH_STAGES] MMAs done reading
  uint64_t* acc_full = bars + 2 * WH_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < WH_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const GradOffTc go = grad_offsets_tc(d);

  if (warp == 5) {
    // ===================================================== producer: one pipeline stage = (stage desc, tile, quarter)
    if (elect_one()) {
      struct Cursor {
        int pass, ti, q, l;
      };
      auto advance = [&](Cursor& c) {
        if (++c.l == h_pass_stage[c.pass + 1]) {
          c.l = h_pass_stage[c.pass];
          if (++c.q == 4) {
            c.q = 0;
            if (++c.ti == my_tiles) {
              c.ti = 0;
              ++c.pass;
              c.l = c.pass < N_PASS ? h_pass_stage[c.pass] : 0;
            }
          }
        }
      };
      auto prefetch = [&](const Cursor& c) {
        if (c.pass >= N_PASS) return;
        const unsigned char* qb = scratch + (size_t)(blockIdx.x + c.ti * gridDim.x) * TILE_BYTES + (size_t)c.q * QUARTER_BYTES;
        const int nl = h_stage[c.l].n_load;
        for (int j = 0; j < nl; ++j) {
          const Load ld = h_stage[c.l].ld[j];
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qb + (size_t)ld.fb * FB_BYTES), "r"(ld.bytes) : "memory");
        }
      };
      Cursor ahead{0, 0, 0, 0};
      if (my_tiles > 0)
        for (int i = 0; i < prefetch_dist; ++i) {
          prefetch(ahead);
          advance(ahead);
        }
      uint32_t it = 0;
      for (int pass = 0; pass < N_PASS && my_tiles > 0; ++pass) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
          const unsigned char* tile = scratch + (size_t)t * TILE_BYTES;
          for (int q = 0; q < 4; ++q) {
            const unsigned char* qb = tile + (size_t)q * QUARTER_BYTES;
            for (int l = h_pass_stage[pass]; l < h_pass_stage[pass + 1]; ++l, ++it) {
              prefetch(ahead);
              advance(ahead);
              const uint32_t s = it % WH_STAGES;
              mbar_wait_parked(&empty[s], ((it / WH_STAGES) & 1) ^ 1);
              unsigned char* st = smem + s * WG_RAW_BYTES;
              const int nl = h_stage[l].n_load;
              uint32_t total = 0;
              for (int j = 0; j < nl; ++j) total += (uint32_t)h_stage[l].ld[j].bytes;
              mbar_expect_tx(&full[s], total);
              for (int j = 0; j < nl; ++j) {
                const Load ld = h_stage[l].ld[j];
                bulk_g2s(st + ld.dst, qb + (size_t)ld.fb * FB_BYTES, (uint32_t)ld.bytes, &full[s]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================================================== MMA issue
    uint32_t it = 0, nf = 0;
    for (int pass = 0; pass < N_PASS; ++pass)
      for (int t0 = 0; t0 < my_tiles; t0 += WG_SEG, ++nf) {
        const int t1 = t0 + WG_SEG < my_tiles ? t0 + WG_SEG : my_tiles;
        mbar_wait_parked(acc_empty, (nf & 1) ^ 1);   // the flush warps have drained the previous segment
        fence_after_sync();
        for (int i = t0 * 4; i < t1 * 4; ++i) {
          for (int l = h_pass_stage[pass]; l < h_pass_stage[pass + 1]; ++l, ++it) {
            const uint32_t first = i == t0 * 4 ? 1u : 0u;
            const uint32_t s = it % WH_STAGES;
            mbar_wait_parked(&full[s], (it / WH_STAGES) & 1);
            fence_after_sync();
            const uint32_t st = smem_addr(smem + s * WG_RAW_BYTES);
            if (elect_one()) {
              const int nm = h_stage[l].n_mma;
              for (int j = 0; j < nm; ++j) {
                const Mma m = h_stage[l].mma[j];
                const uint32_t idesc = idesc_f16_mn(128, m.N);
                const uint32_t dcol = tm + (uint32_t)m.col;
                const uint32_t a0 = st + m.a_off, b0 = st + m.b_off;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {   // 16 points = two core matrices = 256 bytes; lo plane 512 bytes further
                  const uint64_t ah = mn_desc_h(a0 + ks * 256), al = mn_desc_h(a0 + 512 + ks * 256);
                  const uint64_t bh = mn_desc_h(b0 + ks * 256), bl = mn_desc_h(b0 + 512 + ks * 256);
                  mma_ss_f16(dcol, ah, bh, idesc, (first && ks == 0) ? 0u : 1u);
                  mma_ss_f16(dcol, al, bh, idesc, 1u);
                  mma_ss_f16(dcol, ah, bl, idesc, 1u);
                }
              }
              commit(&empty[s]);
            }
            __syncwarp();
          }
        }
        if (elect_one()) commit(acc_full);
        __syncwarp();
      }
  } else {
    // ===================================================== flush (warps 0-3: TMEM lane r <-> output row)
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int r = tid & 127;
    const float inv_x = 1.f / __ldg(sk + T_XIN);
    uint32_t nf = 0;
    for (int pass = 0; pass < N_PASS; ++pass)
      for (int t0 = 0; t0 < my_tiles; t0 += WG_SEG, ++nf) {
        mbar_wait_parked(acc_full, nf & 1);
        fence_after_sync();
        for (int o = h_pass_out[pass]; o < h_pass_out[pass + 1]; ++o) {
          const OutDesc od = h_out[o];
          const int row = od.row0 + r;
          float inv_main, inv_bias;
          out_scales(o, sk, inv_main, inv_bias);
          if (od.kind == OUT_BIAS) {   // column ONES_FEATURE of (rows x XIN)
            float v[16];
            tmem_ld16(lane_t + od.col + 16, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            if (r < od.rows) red_add_h(grad + go.b[od.layer] + row, v[ONES_FEATURE - 16] * inv_bias);
            continue;
          }
          for (int c0 = 0; c0 < od.N; c0 += 16) {
            float v[16];
            tmem_ld16(lane_t + od.col + c0, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            if (od.kind == OUT_DIRECT) {
              float* dst = grad + go.w[od.layer] + (size_t)row * od.N + c0;
              if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j] * inv_main),
                               "f"(v[j + 1] * inv_main), "f"(v[j + 2] * inv_main), "f"(v[j + 3] * inv_main)
                               : "memory");
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) red_add_h(dst + j, v[j] * inv_main);
              }
            } else if (od.kind == OUT_TRANSPOSED) {
              const int in_total = od.layer == 8 ? 256 : 128;
              const int n_out = od.layer == 8 ? d : 64;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < n_out) red_add_h(grad + go.w[od.layer] + (size_t)(c0 + j) * in_total + row, v[j] * inv_main);
            } else if (od.kind == OUT_AUX_S) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < d) red_add_h(aux + AUX_S + (c0 + j) * 256 + row, v[j] * inv_main);
            } else if (od.kind == OUT_XIN) {
              // down_0: D[out = row][k]: k <= d -> weight (XIN scaled by s_x), k = ONES_FEATURE -> bias (the constant 1)
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int k = c0 + j;
                if (k <= d) red_add_h(grad + go.w[0] + (size_t)row * (d + 1) + k, v[j] * inv_main);
                if (k == ONES_FEATURE) red_add_h(grad + go.b[0] + row, v[j] * inv_bias);
              }
            } else {   // OUT_SMALL: rows 0..31: d_y0[j]; rows 32..63: d_o0[j]   (both carry the d_y0 scale)
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int k = c0 + j;
                if (r < 32) {
                  if (r < d && k == ONES_FEATURE) {
                    red_add_h(grad + go.b[8] + r, v[j] * inv_bias);
                    red_add_h(aux + AUX_SB + r, v[j] * inv_bias);
                  }
                } else if (r < 64 && r - 32 < d) {
                  if (k <= d) red_add_h(grad + go.w[3] + (size_t)(r - 32) * (d + 1) + k, v[j] * inv_bias * inv_x);
                  if (k == ONES_FEATURE) red_add_h(grad + go.b[3] + (r - 32), v[j] * inv_bias);
                }
              }
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(acc_empty);
      }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int launch_wgrad_h(const unsigned char* scratch, int n_tiles, int d, const float* scales, float* grad, float* aux,
                   cudaStream_t stream) {
  if (n_tiles <= 0) return SOCM_OK;
  SOCM_CUDA(cudaFuncSetAttribute(wgrad_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WH_SMEM));
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  static const int prefetch_dist = [] {
    const char* e = getenv("SOCM_WH_PREFETCH");
    return e != nullptr ? atoi(e) : WH_PREFETCH;
  }();
  wgrad_h_kernel<<<grid, WH_NT, WH_SMEM, stream>>>(scratch, n_tiles, d, scales, grad, aux, prefetch_dist);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace hx
}  // namespace socm

// debug / test entry: run the fp16 K3b on a caller-built scratch (tests/test_gpu_wgrad_h.py)
extern "C" int socm_debug_wgrad_h(const void* scratch, int32_t n_tiles, int32_t d, const float* scales, float* grad,
                                  float* aux, void* stream) {
  return socm::hx::launch_wgrad_h(static_cast<const unsigned char*>(scratch), n_tiles, d, scales, grad, aux,
                                  static_cast<cudaStream_t>(stream));
}
