// target_h.cu -- K2 forward on the fp16-split engine:  target[B][(K+1)d] = R[B][(2K+1)d] . L^T
// (method.py:584-690 in the re-associated form of SURVEY.md A.3), three kind::f16 MMAs per product, fp32 accumulation.
//
// Structure of target_tc.cu (persistent CTA = 128 paths; the B-independent table L is the "weight matrix", repacked once
// per call into a tape that streams from L2; producer warps turn 32-column pieces of their R rows into shared-memory A
// chunks; segmented accumulation in double-buffered TMEM accumulators), with the operand formats of unet_h.cuh:
//   * both operands are scaled by a power of two from their EXACT max (one streaming pass over R, 0.2 ms per 75 776 paths;
//     scaled max in [128, 256): nothing saturates, the lo halves keep 2^-33 of the largest value) and split into fp16
//     hi / lo; the epilogue multiplies by 1 / (s_R s_L);
//   * a tape slot is [256 rows x 32 columns] hi + lo = 32 KB = the whole K range of one A chunk, so a chunk is 6 MMAs of
//     K = 16 (12 tf32 MMAs of K = 8 before) and one slot wait instead of two;
//   * half as many truncating accumulation steps per chunk: segments of 32 chunks (192 MMAs, as before) halve the
//     read-modify-write traffic on `target`.
#include "kernels.h"
#include "unet_h.cuh"

namespace socm {
namespace hx {

using namespace umma;

constexpr int T_NT = 448;       // warps 0-7 producers (two threads per path, 16 columns each), 8-11 epilogue, 12 MMA, 13 tape
constexpr int T_NB = 256;       // rows of L per block (= accumulator columns)
constexpr int T_MAX_BLOCKS = 64;
constexpr int T_SEG = 32;       // chunks per accumulation segment (6 MMAs each)
constexpr int T_SLOT = 2 * T_NB * 32 * 2;   // 32 KB
constexpr int T_STAGES = 4;
constexpr float T_TARGET = 256.f;
#ifndef SOCM_T_AHEAD
#define SOCM_T_AHEAD 6
#endif
constexpr int T_AHEAD = SOCM_T_AHEAD;   // prefetch distance of the R row pieces, in chunks

struct THPlan {
  int n_blocks;                      // ceil(nrows / 256)
  int chunk_begin[T_MAX_BLOCKS];     // first 32-column chunk of block nt (left of it L is zero)
  int slot_begin[T_MAX_BLOCKS + 1];  // prefix sum of (n_chunks - chunk_begin)
  int n_chunks;                      // ceil(kdim / 32)
};

static THPlan make_plan_h(int nrows, int kdim, int d) {
  THPlan p;
  p.n_blocks = (nrows + T_NB - 1) / T_NB;
  p.n_chunks = (kdim + 31) / 32;
  p.slot_begin[0] = 0;
  for (int nt = 0; nt < p.n_blocks; ++nt) {
    const int i_min = (nt * T_NB) / d;           // first grid time of the block
    p.chunk_begin[nt] = (2 * i_min * d) / 32;    // rows with i >= i_min are zero left of column 2 i_min d
    p.slot_begin[nt + 1] = p.slot_begin[nt] + (p.n_chunks - p.chunk_begin[nt]);
  }
  return p;
}

// max |.| over a [rows][ld] matrix (first n columns) into one uint32 slot (float bits; values >= 0 order like integers)
__global__ void __launch_bounds__(256) absmax1_kernel(const float* __restrict__ A, int rows, int ld, int n, uint32_t* __restrict__ mx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = 0.f;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float4* a4 = reinterpret_cast<const float4*>(A + (size_t)r * ld);   // ld % 4 == 0 (checked by the caller)
    for (int j = lane; j < (n + 3) / 4; j += 32) {
      const float4 v = __ldg(a4 + j);
      const int c = 4 * j;
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), c + 1 < n ? fabsf(v.y) : 0.f), fmaxf(c + 2 < n ? fabsf(v.z) : 0.f, c + 3 < n ? fabsf(v.w) : 0.f)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0 && m > 0.f) atomicMax(mx, __float_as_uint(m));
}

// tape slot (nt, j): rows [256 nt, +256) x columns [32 (chunk_begin + j), +32) of s_L L as fp16 hi slab then lo slab
__global__ void pack_target_tape_h_kernel(const float* __restrict__ L, int nrows, int kdim, int ldr, THPlan plan,
                                          unsigned char* __restrict__ tape, const uint32_t* __restrict__ mx) {
  const float s = pow2_scale(__uint_as_float(mx[1]), T_TARGET);
  const int total_slots = plan.slot_begin[plan.n_blocks];
  for (int sl = blockIdx.x; sl < total_slots; sl += gridDim.x) {
    int nt = 0;
    while (sl >= plan.slot_begin[nt + 1]) ++nt;
    const int k0 = 32 * (plan.chunk_begin[nt] + (sl - plan.slot_begin[nt]));
    unsigned char* base = tape + (size_t)sl * T_SLOT;
    for (int i = threadIdx.x; i < T_NB * 16; i += blockDim.x) {   // pairs of neighbouring columns
      const int n = i >> 4, k = 2 * (i & 15);
      const int row = nt * T_NB + n, col = k0 + k;
      const float w0 = (row < nrows && col < kdim) ? __ldg(L + (size_t)row * ldr + col) : 0.f;
      const float w1 = (row < nrows && col + 1 < kdim) ? __ldg(L + (size_t)row * ldr + col + 1) : 0.f;
      uint32_t hi, lo;
      split_h2(w0 * s, w1 * s, hi, lo);
      const int off = wslab_off(n, k, 32);
      *reinterpret_cast<uint32_t*>(base + off) = hi;
      *reinterpret_cast<uint32_t*>(base + T_SLOT / 2 + off) = lo;
    }
  }
}

namespace k2h {
constexpr int SM_RING = 0;
constexpr int SM_CHUNK = SM_RING + T_STAGES * T_SLOT;
constexpr int SM_BARS = SM_CHUNK + 2 * CHUNK_BYTES;
enum Bar { W_FULL = 0, W_EMPTY = W_FULL + T_STAGES, CH_FULL = W_EMPTY + T_STAGES, CH_EMPTY = CH_FULL + 2,
           ACC_FULL = CH_EMPTY + 2, ACC_EMPTY = ACC_FULL + 2, N_BARS = ACC_EMPTY + 2 };
constexpr int SMEM_BYTES = SM_BARS + N_BARS * 8 + 16;
}  // namespace k2h

__global__ void __launch_bounds__(T_NT, 1)
    target_h_kernel(const float* __restrict__ R, const unsigned char* __restrict__ tape, THPlan plan, int B, int nrows,
                    int kdim, int ldr, float* __restrict__ T, int ldt, const uint32_t* __restrict__ mx) {
  using namespace k2h;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + N_BARS);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = (B + TP - 1) / TP;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < T_STAGES; ++s) {
      mbar_init(&bars[W_FULL + s], 1);
      mbar_init(&bars[W_EMPTY + s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bars[CH_FULL + b], 8);
      mbar_init(&bars[CH_EMPTY + b], 1);
      mbar_init(&bars[ACC_FULL + b], 1);
      mbar_init(&bars[ACC_EMPTY + b], 4);
    }
    mbar_init_fence();
  }
  if (warp == 12) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t ring_s = smem_addr(smem + SM_RING), chunk_s = smem_addr(smem + SM_CHUNK);
  const float s_r = pow2_scale(__uint_as_float(mx[0]), T_TARGET), s_l = pow2_scale(__uint_as_float(mx[1]), T_TARGET);

  if (warp < 8) {
    // ===================================================== producers: R row pieces -> A chunks (fp16 hi / lo, scaled)
    const int p = tid & (TP - 1), half = tid >> 7;   // this thread converts columns [16 half, 16 half + 16) of every chunk
    uint32_t cu = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + p;
      const float* row = R + (size_t)(m < B ? m : 0) * ldr + 16 * half;
      auto load = [&](int c, float4* x) {   // 16 columns of this thread's row: half a 128-byte line
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const int col = 32 * c + 16 * half + 4 * q4;
          x[q4] = (m < B && col < ldr) ? __ldg(reinterpret_cast<const float4*>(row + 32 * c + 4 * q4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }  // ldr % 4 == 0; padding is zero
      };
      // the loads of the next chunk are in flight while the current one is converted (the row pieces come from L2 / DRAM:
      // without this the producers, not the MMAs, set the pace)
      int nt = 0, c = plan.chunk_begin[0];
      float4 cur[4];
      load(c, cur);
      while (nt < plan.n_blocks) {
        int nt2 = nt, c2 = c + 1;
        if (c2 == plan.n_chunks) {
          ++nt2;
          c2 = nt2 < plan.n_blocks ? plan.chunk_begin[nt2] : 0;
        }
        float4 nx[4];
        if (nt2 < plan.n_blocks) load(c2, nx);
        if (half == 0) {   // ... and the line of the chunk T_AHEAD further down the sequence is pulled towards the SM
          int ntp = nt, cp = c + T_AHEAD;
          if (cp >= plan.n_chunks && ntp + 1 < plan.n_blocks) {
            cp = plan.chunk_begin[ntp + 1] + (cp - plan.n_chunks);
            ++ntp;
          }
          if (m < B && cp < plan.n_chunks && 32 * cp < ldr)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(row + 32 * cp));
        }
        float v[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          v[4 * q4] = cur[q4].x * s_r; v[4 * q4 + 1] = cur[q4].y * s_r; v[4 * q4 + 2] = cur[q4].z * s_r; v[4 * q4 + 3] = cur[q4].w * s_r;
        }
        uint32_t hi[8], lo[8];
        split16(v, hi, lo);
        const int b = cu & 1;
        mbar_wait_parked(&bars[CH_EMPTY + b], ((cu >> 1) & 1) ^ 1);
        store_chunk16(smem + SM_CHUNK + b * CHUNK_BYTES, p, 2 * half, hi, lo);
        fence_async_smem();
        warp_arrive(&bars[CH_FULL + b]);
        ++cu;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) cur[q4] = nx[q4];
        nt = nt2;
        c = c2;
      }
    }
  } else if (warp < 12) {
    // ===================================================== epilogue: accumulator -> target rows
    const int p = tid - 256;
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const float inv = 1.f / (s_r * s_l);
    uint32_t ia = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m = t * TP + p;
      for (int nt = 0; nt < plan.n_blocks; ++nt) {
        float* out = T + (size_t)(m < B ? m : 0) * ldt + nt * T_NB;
        for (int c0 = plan.chunk_begin[nt]; c0 < plan.n_chunks; c0 += T_SEG, ++ia) {  // one accumulator per segment
          const bool first = c0 == plan.chunk_begin[nt];
          const uint32_t a = ia & 1;
          mbar_wait_parked(&bars[ACC_FULL + a], (ia >> 1) & 1);
          fence_after_sync();
#pragma unroll 1
          for (int cb = 0; cb < 8; ++cb) {
            float v[32];
            tmem_ld32(lane_t + a * 256 + 32 * cb, reinterpret_cast<uint32_t*>(v));
            tmem_wait_ld();
            if (m < B) {
              const int n_base = nt * T_NB + 32 * cb;
              if (n_base + 32 <= nrows && ((reinterpret_cast<uintptr_t>(out + 32 * cb) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4* dst = reinterpret_cast<float4*>(out + 32 * cb + j);
                  float4 o = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
                  if (!first) {
                    const float4 prev = *dst;
                    o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
                  }
                  *dst = o;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n_base + j < nrows) out[32 * cb + j] = first ? v[j] * inv : out[32 * cb + j] + v[j] * inv;
              }
            }
          }
          fence_before_sync();
          warp_arrive(&bars[ACC_EMPTY + a]);
        }
      }
    }
  } else if (warp == 12) {
    // ===================================================== MMA issue
    uint32_t ws = 0, cm = 0, ia = 0;
    for (int g = 0; g < my_tiles; ++g) {
      for (int nt = 0; nt < plan.n_blocks; ++nt) {
        for (int c0 = plan.chunk_begin[nt]; c0 < plan.n_chunks; c0 += T_SEG, ++ia) {
          const uint32_t a = ia & 1;
          mbar_wait_parked(&bars[ACC_EMPTY + a], ((ia >> 1) & 1) ^ 1);
          fence_after_sync();
          const int c1 = c0 + T_SEG < plan.n_chunks ? c0 + T_SEG : plan.n_chunks;
          for (int c = c0; c < c1; ++c, ++cm, ++ws) {
            const uint32_t b = cm & 1;
            mbar_wait_parked(&bars[CH_FULL + b], (cm >> 1) & 1);
            const uint32_t s = ws % T_STAGES;
            mbar_wait_parked(&bars[W_FULL + s], (ws / T_STAGES) & 1);
            fence_after_sync();
            if (elect_one()) {
              issue_ss<T_NB, 2>(tm + a * 256, chunk_s + b * CHUNK_BYTES, CHUNK_HALF, ring_s + s * T_SLOT, 32, T_NB, c == c0);
              commit(&bars[CH_EMPTY + b]);
              commit(&bars[W_EMPTY + s]);
            }
            __syncwarp();
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
        const uint32_t s = (uint32_t)(i % T_STAGES);
        mbar_wait_parked(&bars[W_EMPTY + s], (uint32_t)((i / T_STAGES) & 1) ^ 1);
        mbar_expect_tx(&bars[W_FULL + s], T_SLOT);
        bulk_g2s(smem + SM_RING + s * T_SLOT, tape + (size_t)slot * T_SLOT, T_SLOT, &bars[W_FULL + s]);
        if (++slot == per_tile) slot = 0;
      }
    }
    __syncwarp();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 12) tmem_dealloc(tm, 512);
}

// workspace: [tape][2 x uint32 maxima]; never larger than the 3xTF32 tape (same bytes per L entry)
int64_t target_h_workspace_bytes(int K, int d) {
  if (((K + 1) * d + T_NB - 1) / T_NB > T_MAX_BLOCKS) return -1;
  const THPlan p = make_plan_h((K + 1) * d, (2 * K + 1) * d, d);
  return (int64_t)p.slot_begin[p.n_blocks] * T_SLOT + 2048;
}

int launch_target_h(const float* L, const float* R, int B, int K, int d, int ldr, float* target, int ldt, void* workspace,
                    cudaStream_t stream) {
  const int nrows = (K + 1) * d, kdim = (2 * K + 1) * d;
  const THPlan plan = make_plan_h(nrows, kdim, d);
  unsigned char* tape = static_cast<unsigned char*>(workspace);
  tape += (1024 - (reinterpret_cast<uintptr_t>(tape) & 1023)) & 1023;
  uint32_t* mx = reinterpret_cast<uint32_t*>(tape + (size_t)plan.slot_begin[plan.n_blocks] * T_SLOT);
  SOCM_CUDA(cudaMemsetAsync(mx, 0, 2 * sizeof(uint32_t), stream));
  absmax1_kernel<<<sm_count() * 8, 256, 0, stream>>>(R, B, ldr, kdim, mx);
  SOCM_LAUNCH_CHECK();
  absmax1_kernel<<<sm_count(), 256, 0, stream>>>(L, nrows, ldr, kdim, mx + 1);
  SOCM_LAUNCH_CHECK();
  pack_target_tape_h_kernel<<<592, 256, 0, stream>>>(L, nrows, kdim, ldr, plan, tape, mx);
  SOCM_LAUNCH_CHECK();
  SOCM_CUDA(cudaFuncSetAttribute(target_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2h::SMEM_BYTES));
  const int n_tiles = (B + TP - 1) / TP;
  const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
  target_h_kernel<<<grid, T_NT, k2h::SMEM_BYTES, stream>>>(R, tape, plan, B, nrows, kdim, ldr, target, ldt, mx);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace hx
}  // namespace socm
