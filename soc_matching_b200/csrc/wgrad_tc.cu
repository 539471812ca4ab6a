// wgrad_tc.cu -- K3b: all weight / bias gradients of the default UNet as tcgen05 GEMMs with the
// trajectory points as the contraction index (3xTF32, fp32 accumulation in TMEM).
//
//   dW[out][in] += sum_p dY[p][out] * Act[p][in]
// Operands come from the scratch written by K3a (loss_tc.cuh): both are K-major with the 128-byte
// swizzle (points contiguous per feature), so a tile quarter of an operand is a run of 4 KB feature
// blocks that one cp.async.bulk lands in shared memory ready for the MMA (one K step = 8 points = 32
// bytes along the row).  A persistent CTA walks the list of
// "layer blocks" (128 output rows x N columns); for each it accumulates over all of its tiles in
// TMEM and flushes once with red.global.add -- the flush traffic is negligible and the kernel is
// bound by streaming the operands from HBM (DESIGN.md, K3b roofline).
//
// The scratch holds plain fp32 values; warps 0-3 split every staged operand in shared memory into
// hi = rn_tf32(x) (in place) and lo = x - hi (exact, second copy), and each K step issues
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo.
//
// Warp roles (192 threads): warps 0-3 split + flush (TMEM -> red.global.add), warp 4 MMA issue, warp 5 producer.
#include "kernels.h"
#include "loss_tc.cuh"
#include "umma.cuh"

namespace socm {
namespace tc {

using namespace umma;

// one unit of work: D[M x N] (+ D2[M x 32] against the XIN block when with_x) over all points
struct LayerBlock {
  int a_fb, M;       // A operand: first feature block, rows (64 or 128)
  int b_fb, N;       // B operand: first feature block, columns (multiple of 32)
  int with_x;        // also multiply A with the FB_XIN block (N = 32): column ONES_FEATURE = bias gradient
  int kind;          // output mapping, see flush()
  int layer, row0;   // socm_unet layer index of the weights, first output row
};
enum { OUT_DIRECT = 0, OUT_TRANSPOSED = 1, OUT_XIN = 2, OUT_SMALL = 3, OUT_BIAS_ONLY = 4, OUT_AUX_S = 5 };
constexpr int N_LB = 14;
__constant__ LayerBlock c_lb[N_LB] = {
    {FB_DZ2, 128, FB_R1, 256, 1, OUT_DIRECT, 1, 0},        // down_1  (+ bias from d_z2)
    {FB_DY1, 128, FB_O2, 128, 1, OUT_DIRECT, 7, 0},        // up_1
    {FB_DY1 + 4, 128, FB_O2, 128, 1, OUT_DIRECT, 7, 128},
    {FB_DO2, 128, FB_R2, 128, 1, OUT_DIRECT, 5, 0},        // res_2
    {FB_DY2, 128, FB_R3, 64, 1, OUT_DIRECT, 6, 0},         // up_2
    {FB_R2, 128, FB_DZ3, 64, 0, OUT_TRANSPOSED, 2, 0},     // down_2: D[in][out]
    {FB_DZ1, 128, FB_XIN, 32, 0, OUT_XIN, 0, 0},           // down_0 rows 0..127 (+ bias via the ones feature)
    {FB_DZ1 + 4, 128, FB_XIN, 32, 0, OUT_XIN, 0, 128},
    {FB_Y1, 128, FB_DY0, 32, 0, OUT_TRANSPOSED, 8, 0},     // up_0, y1 part: D[in][out] = (d_y0^T y1)^T
    {FB_Y1 + 4, 128, FB_DY0, 32, 0, OUT_TRANSPOSED, 8, 128},
    {FB_R1, 128, FB_DY0, 32, 0, OUT_AUX_S, 8, 0},          // S^T = (d_y0^T r1)^T -> aux (res_1 / up_0 via fold_finish_kernel)
    {FB_R1 + 4, 128, FB_DY0, 32, 0, OUT_AUX_S, 8, 128},
    // M is always 128 (an M = 64 accumulator is spread over 16 lanes per TMEM quarter); the extra rows
    // belong to the neighbouring tensors of the scratch and are ignored by the flush
    {FB_DY0, 128, FB_XIN, 32, 0, OUT_SMALL, 3, 0},         // rows 0..31 d_y0 -> b(up_0); rows 32..63 d_o0 -> res_0, b(res_0)
    {FB_DZ3, 128, FB_XIN, 32, 0, OUT_BIAS_ONLY, 2, 0},     // rows 0..63 d_z3 -> bias of down_2
};

constexpr int WG_STAGES = 2;
constexpr int WG_RAW_BYTES = 16384 + 32768 + 4096;  // A | B | XIN as copied from the scratch
constexpr int WG_LO = 53248;                         // offset of the "lo" copy inside a stage
constexpr int WG_STAGE_BYTES = 2 * 53248;
constexpr int WG_A = 0, WG_B = 16384, WG_X = 49152;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers
constexpr int WG_NT = 320;  // warps 0-3 split + flush, 4 MMA, 5 producer, 6-9 split
constexpr int WG_PREFETCH = 6;  // stages of L2 prefetch lookahead
constexpr uint64_t DESC_SW128 = 2ull << 61;  // layout type SWIZZLE_128B

// K-major operand, 128-byte rows (32 points per feature), 8-row swizzle atoms of 1 KB stacked along the features
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) { return smem_desc(saddr, 16, 1024) | DESC_SW128; }

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(WG_NT, 1) wgrad_tc_kernel(const unsigned char* __restrict__ scratch, int n_tiles,
                                                            int d, float* __restrict__ grad, float* __restrict__ aux) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // stage bases must be 512-byte aligned in the shared window: the operand swizzle uses address bits 7-8
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full = bars;                  // [WG_STAGES] bulk copies landed
  uint64_t* empty = bars + WG_STAGES;     // [WG_STAGES] MMAs done reading
  uint64_t* split = bars + 2 * WG_STAGES; // [WG_STAGES] lo copies written
  uint64_t* acc_full = bars + 3 * WG_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split[s], 8);
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
    // ===================================================== producer: one stage = (layer block, tile, quarter)
    // Only two stages fit in shared memory (raw + lo copies), far too little to cover the HBM latency, so
    // the producer runs an L2 prefetch (cp.async.bulk.prefetch.L2) WG_PREFETCH stages ahead of the copies.
    if (elect_one()) {
      struct Cursor {
        int l, ti, q;  // layer block, index into this CTA's tiles, quarter
      };
      auto advance = [&](Cursor& c) {
        if (++c.q == 4) {
          c.q = 0;
          if (++c.ti == my_tiles) {
            c.ti = 0;
            ++c.l;
          }
        }
      };
      auto prefetch = [&](const Cursor& c) {
        if (c.l >= N_LB) return;
        const LayerBlock lb = c_lb[c.l];
        const unsigned char* qb = scratch + (size_t)(blockIdx.x + c.ti * gridDim.x) * TILE_BYTES + (size_t)c.q * QUARTER_BYTES;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qb + (size_t)lb.a_fb * FB_BYTES), "r"(lb.M * 128) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qb + (size_t)lb.b_fb * FB_BYTES), "r"(lb.N * 128) : "memory");
        if (lb.with_x)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qb + (size_t)FB_XIN * FB_BYTES), "r"(FB_BYTES) : "memory");
      };
      Cursor ahead{0, 0, 0};
      if (my_tiles > 0)
        for (int i = 0; i < WG_PREFETCH; ++i) {
          prefetch(ahead);
          advance(ahead);
        }
      uint32_t it = 0;
      for (int l = 0; l < N_LB && my_tiles > 0; ++l) {
        const LayerBlock lb = c_lb[l];
        const uint32_t a_bytes = (uint32_t)lb.M * 128u, b_bytes = (uint32_t)lb.N * 128u;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
          const unsigned char* tile = scratch + (size_t)t * TILE_BYTES;
          for (int q = 0; q < 4; ++q, ++it) {
            prefetch(ahead);
            advance(ahead);
            const uint32_t s = it % WG_STAGES;
            mbar_wait(&empty[s], ((it / WG_STAGES) & 1) ^ 1);
            unsigned char* st = smem + s * WG_STAGE_BYTES;
            const unsigned char* qb = tile + (size_t)q * QUARTER_BYTES;
            mbar_expect_tx(&full[s], a_bytes + b_bytes + (lb.with_x ? FB_BYTES : 0));
            bulk_g2s(st + WG_A, qb + (size_t)lb.a_fb * FB_BYTES, a_bytes, &full[s]);
            bulk_g2s(st + WG_B, qb + (size_t)lb.b_fb * FB_BYTES, b_bytes, &full[s]);
            if (lb.with_x) bulk_g2s(st + WG_X, qb + (size_t)FB_XIN * FB_BYTES, FB_BYTES, &full[s]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================================================== MMA issue
    uint32_t it = 0;
    for (int l = 0; l < N_LB; ++l) {
      const LayerBlock lb = c_lb[l];
      if (my_tiles == 0) break;
      mbar_wait(acc_empty, (l & 1) ^ 1);  // the flush warps have drained the previous block
      fence_after_sync();
      const uint32_t idesc = idesc_tf32(lb.M, lb.N, 0, 0);
      const uint32_t idesc_x = idesc_tf32(lb.M, 32, 0, 0);
      uint32_t first = 1;
      for (int i = 0; i < my_tiles * 4; ++i, ++it) {
        const uint32_t s = it % WG_STAGES;
        mbar_wait(&split[s], (it / WG_STAGES) & 1);
        fence_after_sync();
        const uint32_t st = smem_addr(smem + s * WG_STAGE_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t ad = mn_desc(st + WG_A + ks * 32), al = mn_desc(st + WG_LO + WG_A + ks * 32);
            const uint64_t bd = mn_desc(st + WG_B + ks * 32), bl = mn_desc(st + WG_LO + WG_B + ks * 32);
            mma_ss(tm, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
            mma_ss(tm, al, bd, idesc, 1u);
            mma_ss(tm, ad, bl, idesc, 1u);
            if (lb.with_x) {
              const uint64_t xd = mn_desc(st + WG_X + ks * 32), xl = mn_desc(st + WG_LO + WG_X + ks * 32);
              mma_ss(tm + 256, ad, xd, idesc_x, (first && ks == 0) ? 0u : 1u);
              mma_ss(tm + 256, al, xd, idesc_x, 1u);
              mma_ss(tm + 256, ad, xl, idesc_x, 1u);
            }
          }
          commit(&empty[s]);
        }
        __syncwarp();
        first = 0;
      }
      if (elect_one()) commit(acc_full);
      __syncwarp();
    }
  } else {
    // ===================================================== split (8 warps) + flush (warps 0-3: TMEM lane r <-> output row)
    const bool flusher = warp < 4;
    const int sid = flusher ? tid : tid - 64;  // 0..255 among the split threads
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int r = tid & 127;
    uint32_t it = 0;
    for (int l = 0; l < N_LB; ++l) {
      const LayerBlock lb = c_lb[l];
      if (my_tiles == 0) break;
      // ---- split every stage of this layer block: lo = x - trunc_tf32(x)
      const int a_f4 = lb.M * 8, b_f4 = lb.N * 8;  // float4 counts of the A / B parts (32 points x 4 B per feature)
      for (int i = 0; i < my_tiles * 4; ++i, ++it) {
        const uint32_t s = it % WG_STAGES;
        mbar_wait(&full[s], (it / WG_STAGES) & 1);
        float4* raw = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + WG_LO);
        auto split_range = [&](int f4_begin, int n_f4) {
          for (int j = sid; j < n_f4; j += 256) {
            const float4 x = raw[f4_begin + j];
            float4 hi, y;
            hi.x = tf32_rn(x.x); hi.y = tf32_rn(x.y); hi.z = tf32_rn(x.z); hi.w = tf32_rn(x.w);
            y.x = x.x - hi.x; y.y = x.y - hi.y; y.z = x.z - hi.z; y.w = x.w - hi.w;
            raw[f4_begin + j] = hi;  // round-to-nearest split (unbiased; the MMA would truncate)
            lo[f4_begin + j] = y;
          }
        };
        split_range(WG_A / 16, a_f4);
        split_range(WG_B / 16, b_f4);
        if (lb.with_x) split_range(WG_X / 16, FB_BYTES / 16);
        fence_async_smem();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&split[s]);
      }
      if (!flusher) continue;
      mbar_wait(acc_full, l & 1);
      fence_after_sync();
      const bool row_ok = true;
      const int row = lb.row0 + r;
      {
        for (int c0 = 0; c0 < lb.N; c0 += 16) {
          float v[16];
          tmem_ld16(lane_t + c0, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          if (!row_ok) continue;
          if (lb.kind == OUT_DIRECT) {
            float* dst = grad + go.w[lb.layer] + (size_t)row * lb.N + c0;
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
            } else {  // the flat gradient offsets are not 16-byte aligned for every d
#pragma unroll
              for (int j = 0; j < 16; ++j) red_add(dst + j, v[j]);
            }
          } else if (lb.kind == OUT_TRANSPOSED) {
            // D[in = row][out = c]: weight (layer) is [out][in_total]; up_0: out < d, in_total = 256; down_2: in_total = 128
            const int in_total = lb.layer == 8 ? 256 : 128;
            const int n_out = lb.layer == 8 ? d : 64;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < n_out) red_add(grad + go.w[lb.layer] + (size_t)(c0 + j) * in_total + row, v[j]);
          } else if (lb.kind == OUT_AUX_S) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < d) red_add(aux + AUX_S + (c0 + j) * 256 + row, v[j]);
          } else if (lb.kind == OUT_XIN) {
            // down_0: D[out = row][k]: k <= d -> weight, k = ONES_FEATURE -> bias
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = c0 + j;
              if (k <= d) red_add(grad + go.w[0] + (size_t)row * (d + 1) + k, v[j]);
              if (k == ONES_FEATURE) red_add(grad + go.b[0] + row, v[j]);
            }
          } else if (lb.kind == OUT_SMALL) {
            // rows 0..31: d_y0[j]; rows 32..63: d_o0[j]
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = c0 + j;
              if (r < 32) {
                if (r < d && k == ONES_FEATURE) {
                  red_add(grad + go.b[8] + r, v[j]);
                  red_add(aux + AUX_SB + r, v[j]);
                }
              } else if (r < 64 && r - 32 < d) {
                if (k <= d) red_add(grad + go.w[3] + (size_t)(r - 32) * (d + 1) + k, v[j]);
                if (k == ONES_FEATURE) red_add(grad + go.b[3] + (r - 32), v[j]);
              }
            }
          } else {  // OUT_BIAS_ONLY (down_2: rows = d_z3 features)
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (r < 64 && c0 + j == ONES_FEATURE) red_add(grad + go.b[2] + r, v[j]);
          }
        }
        if (lb.with_x) {  // bias of the layer: column ONES_FEATURE of A x XIN
          float v[16];
          tmem_ld16(lane_t + 256 + 16, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          if (row_ok) red_add(grad + go.b[lb.layer] + row, v[ONES_FEATURE - 16]);
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

int launch_wgrad_tc(const unsigned char* scratch, int n_tiles, int d, float* grad, float* aux, cudaStream_t stream) {
  if (n_tiles <= 0) return SOCM_OK;
  SOCM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  wgrad_tc_kernel<<<grid, WG_NT, WG_SMEM, stream>>>(scratch, n_tiles, d, grad, aux);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace tc
}  // namespace socm

// debug / test entry: run K3b on a caller-built scratch (tests/test_gpu_wgrad_tc.py)
extern "C" int socm_debug_wgrad_tc(const void* scratch, int32_t n_tiles, int32_t d, float* grad, float* aux, void* stream) {
  return socm::tc::launch_wgrad_tc(static_cast<const unsigned char*>(scratch), n_tiles, d, grad, aux,
                                   static_cast<cudaStream_t>(stream));
}
extern "C" int64_t socm_debug_wgrad_tile_bytes(void) { return socm::tc::TILE_BYTES; }
