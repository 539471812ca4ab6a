// target_bwd_h.cu -- K2 backward on the fp16-split engine:  dL[(K+1)d][(2K+1)d] += G^T R  (the gradient of the SOCM
// target w.r.t. the M-table, contraction over the paths; replaces the autograd backward of method.py:584-690).
//
// Same block decomposition as target_bwd_tc.cu (every CTA owns [128 rows of dL] x [256 columns] blocks right of the block
// diagonal and streams all path quarters), but the operands are prepared once per call as fp16 hi / lo planes of the
// power-of-two-scaled values in the MN-major core-matrix layout of the K3 scratch (loss_h.cu: store_fb16):
//     feature block = 32 features x 32 paths = 4 KB;  byte(f, r, plane) = (f/8) 1024 + plane 512 + r 16 + (f%8) 2 .
//   1. absmax_kernel: exact max |G|, max |R| (the scales map them into [128, 256): nothing saturates, and the lo plane
//      keeps an absolute precision of 2^-25 of that, i.e. 2^-33 of the largest value);
//   2. pack_planes_kernel: [paths][features] fp32 -> planes.  A thread converts 8 neighbouring features of one path
//      (32 contiguous bytes in, two 16-byte stores out, a warp store = 512 contiguous bytes): no transposition is
//      needed for an MN-major operand, so this replaces the two shared-memory transposes of the 3xTF32 version;
//   3. target_bwd_h_kernel: per (block, quarter) one cp.async.bulk stage of 16 KB (A: 128 rows of G^T) + 32 KB
//      (B: 256 rows of R^T) that is MMA-ready as it lands (no in-kernel split: four stages instead of two), and
//      a_hi b_hi + a_lo b_hi + a_hi b_lo as 6 kind::f16 MMAs of K = 16 paths (12 tf32 MMAs of K = 8 before).
//      Accumulation in TMEM in segments, flushed by red.global.add x 1 / (s_G s_R); one owner per block.
#include "kernels.h"
#include "loss_tc.cuh"
#include "umma.cuh"
#include "unet_h.cuh"

namespace socm {
namespace hx {

using namespace umma;
using tc::FB_BYTES;

struct K2hGeom {
  int gfb, rfb;      // feature blocks of G^T (ceil(nrows/128) * 4) and R^T (ceil(kdim/256) * 8)
  int n_tiles;       // ceil(B / 128)
  int nrows, kdim, d;
};
__host__ __device__ inline int64_t k2h_quarter_bytes(const K2hGeom& g) { return (int64_t)(g.gfb + g.rfb) * FB_BYTES; }

constexpr float K2H_TARGET = 256.f;   // scaled max in [128, 256)

// mx[0] = max |G[:, :nG]|, mx[1] = max |R[:, :nR]| as float bit patterns (values >= 0 order like unsigned integers)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ G, int ldg, int nG, const float* __restrict__ R,
                                                     int ldr, int nR, int B, uint32_t* __restrict__ mx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mg = 0.f, mr = 0.f;
  for (int m = blockIdx.x * 8 + warp; m < B; m += gridDim.x * 8) {
    const float4* g4 = reinterpret_cast<const float4*>(G + (size_t)m * ldg);   // ldg, ldr are multiples of 4 (checked)
    const float4* r4 = reinterpret_cast<const float4*>(R + (size_t)m * ldr);
    for (int j = lane; j < (nG + 3) / 4; j += 32) {
      const float4 v = __ldg(g4 + j);
      const int c = 4 * j;
      mg = fmaxf(mg, fmaxf(fmaxf(fabsf(v.x), c + 1 < nG ? fabsf(v.y) : 0.f), fmaxf(c + 2 < nG ? fabsf(v.z) : 0.f, c + 3 < nG ? fabsf(v.w) : 0.f)));
    }
    for (int j = lane; j < (nR + 3) / 4; j += 32) {
      const float4 v = __ldg(r4 + j);
      const int c = 4 * j;
      mr = fmaxf(mr, fmaxf(fmaxf(fabsf(v.x), c + 1 < nR ? fabsf(v.y) : 0.f), fmaxf(c + 2 < nR ? fabsf(v.z) : 0.f, c + 3 < nR ? fabsf(v.w) : 0.f)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
  }
  if (lane == 0) {
    if (mg > 0.f) atomicMax(mx, __float_as_uint(mg));
    if (mr > 0.f) atomicMax(mx + 1, __float_as_uint(mr));
  }
}

// src[B][ld] (row-major fp32) -> planes [tile][quarter][fb0 + f/32]: one warp per (path quarter, feature block), lane = path
__global__ void __launch_bounds__(256) pack_planes_kernel(const float* __restrict__ src, int B, int ld, int nfeat, int n_fb,
                                                          int fb0, int64_t quarter_bytes, unsigned char* __restrict__ scratch,
                                                          const uint32_t* __restrict__ mx) {
  const float s = pow2_scale(__uint_as_float(*mx), K2H_TARGET);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_q = (B + 31) / 32;
  const int64_t total = (int64_t)n_q * n_fb;
  for (int64_t w = (int64_t)blockIdx.x * 8 + warp; w < total; w += (int64_t)gridDim.x * 8) {
    const int qi = (int)(w / n_fb), fb = (int)(w - (int64_t)qi * n_fb);
    const int m = 32 * qi + lane;
    unsigned char* blk = scratch + (size_t)(qi >> 2) * 4 * quarter_bytes + (size_t)(qi & 3) * quarter_bytes +
                         (size_t)(fb0 + fb) * FB_BYTES + lane * 16;
    const float* row = src + (size_t)(m < B ? m : 0) * ld + 32 * fb;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[8];
      const int f0 = 32 * fb + 8 * g;
      if (m < B && f0 + 7 < nfeat) {   // rows are 16-byte aligned (ld % 4 == 0)
        const float4 a = __ldg(reinterpret_cast<const float4*>(row + 8 * g));
        const float4 b = __ldg(reinterpret_cast<const float4*>(row + 8 * g + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (m < B && f0 + j < nfeat) ? __ldg(row + 8 * g + j) : 0.f;
      }
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_h2(v[2 * j] * s, v[2 * j + 1] * s, hi[j], lo[j]);
      __stcs(reinterpret_cast<uint4*>(blk + g * 1024), make_uint4(hi[0], hi[1], hi[2], hi[3]));
      __stcs(reinterpret_cast<uint4*>(blk + g * 1024 + 512), make_uint4(lo[0], lo[1], lo[2], lo[3]));
    }
  }
}

constexpr int KH_STAGES = 4;
constexpr int KH_RAW = 16384 + 32768;        // A (128 features) | B (256 features), 32 paths each, hi and lo planes
constexpr int KH_SMEM = KH_STAGES * KH_RAW + 1024 + 256;
constexpr int KH_NT = 192;                   // warps 0-3 flush, 4 MMA, 5 producer
constexpr int KH_SEG = 32;                   // stages (of 6 MMAs) per TMEM accumulation segment

// MN-major, no swizzle: LBO = 128 (next 8 paths), SBO = 1024 (next 8 features)
__device__ __forceinline__ uint64_t kh_desc(uint32_t saddr) { return smem_desc(saddr, 128, 1024); }

// blocks[i] = rb | (cb << 16): rows [128 rb, +128) of dL, columns [256 cb, +256)
__global__ void __launch_bounds__(KH_NT, 1)
    target_bwd_h_kernel(const unsigned char* __restrict__ scratch, K2hGeom g, const int* __restrict__ blocks, int n_blocks,
                        float* __restrict__ dL, int ldr, const uint32_t* __restrict__ mx) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KH_STAGES * KH_RAW);
  uint64_t* full = bars;
  uint64_t* empty = bars + KH_STAGES;
  uint64_t* acc_full = bars + 2 * KH_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_stage_blk = g.n_tiles * 4;                       // stages (path quarters) per block
  const int n_seg = (n_stage_blk + KH_SEG - 1) / KH_SEG;
  const int64_t qbytes = k2h_quarter_bytes(g);

  if (tid == 0) {
    for (int s = 0; s < KH_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    mbar_init_fence();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 512);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == 5) {
    // ===================================================== producer
    if (elect_one()) {
      uint32_t it = 0;
      for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
        const int rb = blocks[bi] & 0xFFFF, cb = blocks[bi] >> 16;
        for (int q = 0; q < n_stage_blk; ++q, ++it) {
          const uint32_t s = it % KH_STAGES;
          mbar_wait_parked(&empty[s], ((it / KH_STAGES) & 1) ^ 1);
          unsigned char* st = smem + s * KH_RAW;
          const unsigned char* qb = scratch + (size_t)q * qbytes;
          mbar_expect_tx(&full[s], KH_RAW);
          bulk_g2s(st, qb + (size_t)(4 * rb) * FB_BYTES, 16384, &full[s]);
          bulk_g2s(st + 16384, qb + (size_t)(g.gfb + 8 * cb) * FB_BYTES, 32768, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================================================== MMA issue
    constexpr uint32_t idesc = idesc_f16_mn(128, 256);
    uint32_t it = 0, ia = 0;
    for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
      for (int sg = 0; sg < n_seg; ++sg, ++ia) {
        const uint32_t a = ia & 1;
        mbar_wait_parked(&acc_empty[a], ((ia >> 1) & 1) ^ 1);
        fence_after_sync();
        const int q1 = (sg + 1) * KH_SEG < n_stage_blk ? (sg + 1) * KH_SEG : n_stage_blk;
        for (int q = sg * KH_SEG; q < q1; ++q, ++it) {
          const uint32_t s = it % KH_STAGES;
          mbar_wait_parked(&full[s], (it / KH_STAGES) & 1);
          fence_after_sync();
          const uint32_t st = smem_addr(smem + s * KH_RAW);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {   // 16 paths = two core matrices = 256 bytes; lo plane 512 bytes further
              const uint64_t ah = kh_desc(st + ks * 256), al = kh_desc(st + 512 + ks * 256);
              const uint64_t bh = kh_desc(st + 16384 + ks * 256), bl = kh_desc(st + 16384 + 512 + ks * 256);
              mma_ss_f16(tm + a * 256, ah, bh, idesc, (q == sg * KH_SEG && ks == 0) ? 0u : 1u);
              mma_ss_f16(tm + a * 256, al, bh, idesc, 1u);
              mma_ss_f16(tm + a * 256, ah, bl, idesc, 1u);
            }
            commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) commit(&acc_full[a]);
        __syncwarp();
      }
    }
  } else {
    // ===================================================== warps 0-3: flush of every segment
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const float inv = 1.f / (pow2_scale(__uint_as_float(mx[0]), K2H_TARGET) * pow2_scale(__uint_as_float(mx[1]), K2H_TARGET));
    uint32_t ia = 0;
    for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
      const int rb = blocks[bi] & 0xFFFF, cb = blocks[bi] >> 16;
      const int n = 128 * rb + (tid & 127);         // row of dL owned by this thread
      const int k_lo = 2 * (n / g.d) * g.d;         // columns left of it are structurally zero
      for (int sg = 0; sg < n_seg; ++sg, ++ia) {
        const uint32_t a = ia & 1;
        mbar_wait_parked(&acc_full[a], (ia >> 1) & 1);
        fence_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          float v[32];
          tmem_ld32(lane_t + a * 256 + c0, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          if (n < g.nrows) {
            float* dst = dL + (size_t)n * ldr + 256 * cb + c0;   // this thread is the only writer of its row segment
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int k = 256 * cb + c0 + j;
              if (k >= k_lo && k + 3 < g.kdim) {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j] * inv), "f"(v[j + 1] * inv),
                             "f"(v[j + 2] * inv), "f"(v[j + 3] * inv)
                             : "memory");
              } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  if (k + jj < g.kdim && k + jj >= k_lo)
                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + j + jj), "f"(v[j + jj] * inv) : "memory");
              }
            }
          }
        }
        fence_before_sync();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&acc_empty[a]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

// workspace layout: [planes: n_tiles * 4 quarters][block list][2 x uint32 maxima]; same size as the 3xTF32 version's
int launch_target_bwd_h(const float* G, const float* R, int B, int K, int d, int ldr, int ldt, float* dL, void* workspace,
                        cudaStream_t stream) {
  K2hGeom g;
  g.nrows = (K + 1) * d;
  g.kdim = (2 * K + 1) * d;
  g.d = d;
  g.gfb = ((g.nrows + 127) / 128) * 4;
  g.rfb = ((g.kdim + 255) / 256) * 8;
  g.n_tiles = (B + 127) / 128;
  SOCM_CHECK_ARG(ldr % 4 == 0 && ldt % 4 == 0, "pitches must be multiples of 4 floats (ldr=%d ldt=%d)", ldr, ldt);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  ws += (1024 - (reinterpret_cast<uintptr_t>(ws) & 1023)) & 1023;
  const int64_t qbytes = k2h_quarter_bytes(g);
  unsigned char* scratch = ws;
  int* blocks_dev = reinterpret_cast<int*>(ws + (size_t)g.n_tiles * 4 * qbytes);
  int blocks_host[4096];
  int n_blocks = 0;
  for (int rb = 0; rb < g.gfb / 4; ++rb)
    for (int cb = 0; cb < g.rfb / 8; ++cb) {
      const int i_min = (128 * rb) / d;                    // smallest grid time among the rows of the block
      if (256 * cb + 255 >= 2 * i_min * d && 256 * cb < g.kdim && 128 * rb < g.nrows && n_blocks < 4096)
        blocks_host[n_blocks++] = rb | (cb << 16);
    }
  SOCM_CHECK_ARG(n_blocks < 4096, "too many blocks for the tcgen05 target-backward kernel");
  const int64_t n_pairs = (int64_t)(g.gfb / 4) * (g.rfb / 8);
  uint32_t* mx = reinterpret_cast<uint32_t*>(blocks_dev + n_pairs);   // inside the 4096 bytes of slack of the workspace
  SOCM_CUDA(cudaMemcpyAsync(blocks_dev, blocks_host, n_blocks * sizeof(int), cudaMemcpyHostToDevice, stream));
  SOCM_CUDA(cudaMemsetAsync(mx, 0, 2 * sizeof(uint32_t), stream));
  absmax_kernel<<<sm_count() * 8, 256, 0, stream>>>(G, ldt, g.nrows, R, ldr, g.kdim, B, mx);
  SOCM_LAUNCH_CHECK();
  // quarters beyond B inside the last tile are written as zeros by the pack kernels (they cover whole quarters up to
  // ceil(B / 32)); quarters beyond that must hold zeros too: the GEMM streams whole tiles
  if (((B + 31) / 32) % 4)
    SOCM_CUDA(cudaMemsetAsync(scratch + (size_t)(g.n_tiles - 1) * 4 * qbytes, 0, (size_t)4 * qbytes, stream));
  pack_planes_kernel<<<sm_count() * 16, 256, 0, stream>>>(G, B, ldt, g.nrows, g.gfb, 0, qbytes, scratch, mx);
  SOCM_LAUNCH_CHECK();
  pack_planes_kernel<<<sm_count() * 16, 256, 0, stream>>>(R, B, ldr, g.kdim, g.rfb, g.gfb, qbytes, scratch, mx + 1);
  SOCM_LAUNCH_CHECK();
  SOCM_CUDA(cudaFuncSetAttribute(target_bwd_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KH_SMEM));
  const int grid = n_blocks < sm_count() ? n_blocks : sm_count();
  target_bwd_h_kernel<<<grid, KH_NT, KH_SMEM, stream>>>(scratch, g, blocks_dev, n_blocks, dL, ldr, mx);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace hx
}  // namespace socm
