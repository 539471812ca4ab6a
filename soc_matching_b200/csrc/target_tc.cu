// target_tc.cu -- K2 forward on the 5th-gen tensor cores:  target[B][(K+1)d] = R[B][(2K+1)d] . L^T
// (method.py:584-690 in the re-associated form of SURVEY.md A.3), 3xTF32 with fp32 accumulation.
//
// The block upper-triangular table L (rows (i,k), columns (j,k') with j >= i) plays the role of a
// weight matrix: it is B-independent, so it is split hi/lo and repacked ONCE per call into a tape of
// 32 KB slots ([256 rows x 16 columns] hi + lo, canonical K-major core-matrix order), only the column
// range right of the diagonal for every 256-row block.  A persistent CTA owns 128 paths; for each
// 256-row block of L it
//   * producer warps 0-3 (thread <-> path) read 32 columns of their R row, split them hi/lo and write a
//     shared-memory A chunk (same chunk format and double buffer as the 256-wide UNet layers);
//   * warp 8 issues  D[128 x 256] += A_chunk . L_block^T  (3 MMAs per K step) from the streamed tape;
//   * epilogue warps 4-7 drain the finished accumulator (TMEM, double buffered) to `target`.
// Per 128 paths the whole tape (about half of 2 * (K+1)d * (2K+1)d * 4 bytes) streams from L2 once.
#include "kernels.h"
#include "unet_tc.cuh"

namespace socm {
namespace tc {

using namespace umma;

constexpr int K2_NT = 320;      // warps 0-3 producers, 4-7 epilogue, 8 MMA, 9 tape
constexpr int K2_NB = 256;      // rows of L per block (= accumulator columns)
constexpr int K2_MAX_BLOCKS = 64;
// The tensor core adds into its fp32 accumulator with truncation, so a K = 4010 contraction in one
// accumulator (1500 MMAs) loses ~2e-5; the contraction is therefore cut into segments of K2_SEG chunks
// (192 MMAs, error ~2e-6 like the UNet layers) whose partial sums are added in fp32 (RN) by the
// epilogue threads.
constexpr int K2_SEG = 16;

struct K2Plan {
  int n_blocks;                       // ceil(nrows / 256)
  int chunk_begin[K2_MAX_BLOCKS];     // first 32-column chunk of block nt (left of it L is zero)
  int slot_begin[K2_MAX_BLOCKS + 1];  // prefix sum of 2 * (n_chunks - chunk_begin)
  int n_chunks;                       // ceil(kdim / 32)
};

static K2Plan make_plan(int nrows, int kdim, int d) {
  K2Plan p;
  p.n_blocks = (nrows + K2_NB - 1) / K2_NB;
  p.n_chunks = (kdim + 31) / 32;
  p.slot_begin[0] = 0;
  for (int nt = 0; nt < p.n_blocks; ++nt) {
    const int i_min = (nt * K2_NB) / d;          // first grid time of the block
    p.chunk_begin[nt] = (2 * i_min * d) / 32;    // rows with i >= i_min are zero left of column 2 i_min d
    p.slot_begin[nt + 1] = p.slot_begin[nt] + 2 * (p.n_chunks - p.chunk_begin[nt]);
  }
  return p;
}

// tape slot (nt, j): rows [256 nt, +256) x columns [16 (2 chunk_begin + j), +16) of L, hi slab then lo slab
__global__ void pack_target_tape_kernel(const float* __restrict__ L, int nrows, int kdim, int ldr, K2Plan plan,
                                        unsigned char* __restrict__ tape) {
  const int total_slots = plan.slot_begin[plan.n_blocks];
  for (int s = blockIdx.x; s < total_slots; s += gridDim.x) {
    int nt = 0;
    while (s >= plan.slot_begin[nt + 1]) ++nt;
    const int k0 = 16 * (2 * plan.chunk_begin[nt] + (s - plan.slot_begin[nt]));
    unsigned char* base = tape + (size_t)s * MAIN_BYTES;
    for (int i = threadIdx.x; i < K2_NB * 16; i += blockDim.x) {
      const int n = i >> 4, k = i & 15;
      const int row = nt * K2_NB + n, col = k0 + k;
      const float w = (row < nrows && col < kdim) ? __ldg(L + (size_t)row * ldr + col) : 0.f;
      const float hi = tf32_rn(w);
      const int off = wslab_off(n, k, 16);
      *reinterpret_cast<float*>(base + off) = hi;
      *reinterpret_cast<float*>(base + K2_NB * 16 * 4 + off) = w - hi;
    }
  }
}

namespace k2 {
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + NSTAGE * MAIN_BYTES;
constexpr int SM_BARS = SM_CHUNK + 2 * CHUNK_BYTES;
enum Bar { W_FULL = 0, W_EMPTY = W_FULL + NSTAGE, CH_FULL = W_EMPTY + NSTAGE, CH_EMPTY = CH_FULL + 2,
           ACC_FULL = CH_EMPTY + 2, ACC_EMPTY = ACC_FULL + 2, N_BARS = ACC_EMPTY + 2 };
constexpr int SMEM_BYTES = SM_BARS + N_BARS * 8 + 16;
}  // namespace k2

__global__ void __launch_bounds__(K2_NT, 1)
    target_tc_kernel(const float* __restrict__ R, const unsigned char* __restrict__ tape, K2Plan plan, int B, int nrows,
                     int kdim, int ldr, float* __restrict__ T, int ldt) {
  using namespace k2;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&bars[W_FULL + s], 1);
      mbar_init(&bars[W_EMPTY + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[CH_FULL + b], 4);
      mbar_init(&bars[CH_EMPTY + b], 1);
      mbar_init(&bars[ACC_FULL + b], 1);
      mbar_init(&bars[ACC_EMPTY + b], 4);
    }
    mbar_init_fence();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK);

  if (warp < 4) {
    // ===================================================== producers: R row pieces -> A chunks
    const int p = tid;
    uint32_t cu = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + p;
      const float* row = R + (size_t)(m < B ? m : 0) * ldr;
      for (int nt = 0; nt < plan.n_blocks; ++nt) {
        for (int c = plan.chunk_begin[nt]; c < plan.n_chunks; ++c, ++cu) {
          float v[32];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const int col = 32 * c + 4 * q4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < B && col < ldr) x = __ldg(reinterpret_cast<const float4*>(row + col));  // ldr % 4 == 0; padding is zero
            v[4 * q4] = x.x; v[4 * q4 + 1] = x.y; v[4 * q4 + 2] = x.z; v[4 * q4 + 3] = x.w;
          }
          const int b = cu & 1;
          mbar_wait(&bars[CH_EMPTY + b], ((cu >> 1) & 1) ^ 1);
          store_chunk32(smem + SM_CHUNK + b * CHUNK_BYTES, p, v);
          fence_async_smem();
          warp_arrive(&bars[CH_FULL + b]);
        }
      }
    }
  } else if (warp < 8) {
    // ===================================================== epilogue: accumulator -> target rows
    const int p = tid - 128;
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ia = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + p;
      for (int nt = 0; nt < plan.n_blocks; ++nt) {
        float* out = T + (size_t)(m < B ? m : 0) * ldt + nt * K2_NB;
        for (int c0 = plan.chunk_begin[nt]; c0 < plan.n_chunks; c0 += K2_SEG, ++ia) {  // one accumulator per segment
          const bool first = c0 == plan.chunk_begin[nt];
          const uint32_t a = ia & 1;
          mbar_wait(&bars[ACC_FULL + a], (ia >> 1) & 1);
          fence_after_sync();
#pragma unroll 1
          for (int cb = 0; cb < 8; ++cb) {
            float v[32];
            tmem_ld32(lane_t + a * 256 + 32 * cb, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            if (m < B) {
              const int n_base = nt * K2_NB + 32 * cb;
              if (n_base + 32 <= nrows && ((reinterpret_cast<uintptr_t>(out + 32 * cb) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4* dst = reinterpret_cast<float4*>(out + 32 * cb + j);
                  float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                  if (!first) {
                    const float4 prev = *dst;
                    o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                  }
                  *dst = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n_base + j < nrows) out[32 * cb + j] = first ? v[j] : out[32 * cb + j] + v[j];
              }
            }
          }
          fence_before_sync();
          warp_arrive(&bars[ACC_EMPTY + a]);
        }
      }
    }
  } else if (warp == 8) {
    // ===================================================== MMA issue
    uint32_t ws = 0, cm = 0, ia = 0;
    for (int g = 0; g < my_tiles; ++g) {
      for (int nt = 0; nt < plan.n_blocks; ++nt) {
        for (int c0 = plan.chunk_begin[nt]; c0 < plan.n_chunks; c0 += K2_SEG, ++ia) {
          const uint32_t a = ia & 1;
          mbar_wait(&bars[ACC_EMPTY + a], ((ia >> 1) & 1) ^ 1);
          fence_after_sync();
          const int c1 = c0 + K2_SEG < plan.n_chunks ? c0 + K2_SEG : plan.n_chunks;
          for (int c = c0; c < c1; ++c, ++cm) {
            const uint32_t b = cm & 1;
            mbar_wait(&bars[CH_FULL + b], (cm >> 1) & 1);
            fence_after_sync();
            for (int j = 0; j < 2; ++j, ++ws) {
              const uint32_t s = ws % NSTAGE;
              mbar_wait(&bars[W_FULL + s], (ws / NSTAGE) & 1);
              fence_after_sync();
              if (elect_one()) {
                issue_block_ss<K2_NB, 16>(tm + a * 256, chunk_s + b * CHUNK_BYTES + j * 2 * ACT_KSTEP, CHUNK_HALF,
                                          ring_s + s * MAIN_BYTES, c == c0 && j == 0);
                if (j == 1) commit(&bars[CH_EMPTY + b]);
                commit(&bars[W_EMPTY + s]);
              }
              __syncwarp();
            }
          }
          if (elect_one()) commit(&bars[ACC_FULL + a]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================== tape producer: the whole tape once per path tile
    if (elect_one()) {
      const uint32_t per_tile = (uint32_t)plan.slot_begin[plan.n_blocks];
      const uint64_t total = (uint64_t)my_tiles * per_tile;
      uint32_t slot = 0;
      for (uint64_t i = 0; i < total; ++i) {
        const uint32_t s = (uint32_t)(i % NSTAGE);
        mbar_wait(&bars[W_EMPTY + s], (uint32_t)((i / NSTAGE) & 1) ^ 1);
        mbar_expect_tx(&bars[W_FULL + s], MAIN_BYTES);
        bulk_g2s(smem + SM_RING + s * MAIN_BYTES, tape + (size_t)slot * MAIN_BYTES, MAIN_BYTES, &bars[W_FULL + s]);
        if (++slot == per_tile) slot = 0;
      }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

}  // namespace tc
}  // namespace socm

using namespace socm;

extern "C" int64_t socm_target_gemm_tc_workspace_bytes(int32_t K, int32_t d) {
  if (K < 1 || d < 1) return -1;
  // same bound as socm_target_gemm_tc_f32: make_plan fills fixed-size tables (K2_MAX_BLOCKS row blocks)
  if (((K + 1) * d + tc::K2_NB - 1) / tc::K2_NB > tc::K2_MAX_BLOCKS) return -1;
  const tc::K2Plan p = tc::make_plan((K + 1) * d, (2 * K + 1) * d, d);
  const int64_t tcb = (int64_t)p.slot_begin[p.n_blocks] * tc::MAIN_BYTES + 1024;
  const int64_t hb = hx::target_h_workspace_bytes(K, d);   // the fp16 tape has half as many slots
  return tcb > hb ? tcb : hb;
}

extern "C" int socm_target_gemm_tc_f32(const float* L, const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                                       float* target, int32_t ldt, void* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SOCM_CHECK_ARG(L && R && target && workspace, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1, "bad sizes");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * d && ldr % 4 == 0 && ldt >= (K + 1) * d, "bad pitches ldr=%d ldt=%d", ldr, ldt);
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  SOCM_CHECK_ARG((nrows + tc::K2_NB - 1) / tc::K2_NB <= tc::K2_MAX_BLOCKS, "(K+1)d = %d too large for the tcgen05 target kernel", nrows);
  if (B == 0) return SOCM_OK;
  if (f16_default() != 0) return hx::launch_target_h(L, R, B, K, d, ldr, target, ldt, workspace, stream);
  const tc::K2Plan plan = tc::make_plan(nrows, kdim, d);
  unsigned char* tape = static_cast<unsigned char*>(workspace);
  tape += (1024 - (reinterpret_cast<uintptr_t>(tape) & 1023)) & 1023;
  tc::pack_target_tape_kernel<<<592, 256, 0, stream>>>(L, nrows, kdim, ldr, plan, tape);
  SOCM_LAUNCH_CHECK();
  SOCM_CUDA(cudaFuncSetAttribute(tc::target_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::k2::SMEM_BYTES));
  const int n_tiles = (B + tc::TP - 1) / tc::TP;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  tc::target_tc_kernel<<<grid, tc::K2_NT, tc::k2::SMEM_BYTES, stream>>>(R, tape, plan, B, nrows, kdim, ldr, target, ldt);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
