// target_bwd_tc.cu -- K2 backward on the 5th-gen tensor cores:  dL[(K+1)d][(2K+1)d] += G^T R
// (the gradient of the SOCM target w.r.t. the M-table, contraction over the paths; replaces the
// autograd backward of method.py:584-690), 3xTF32 with fp32 accumulation in TMEM.
//
// Both operands are [paths][features] row-major; tcgen05 wants the contraction index contiguous, so
//   1. transpose_pack_kernel rewrites G and R into the K-major / 128-byte-swizzle "feature block"
//      layout of loss_tc.cuh (32 features x 32 paths = 4 KB, one bulk copy lands it MMA-ready);
//   2. target_bwd_tc_kernel gives every CTA a list of [128 rows of dL] x [256 columns] blocks right of
//      the block diagonal; per block it streams ALL path quarters (cp.async.bulk, 2 stages), splits
//      hi/lo in shared memory (warps 0-3), accumulates in TMEM in segments and adds the result into dL
//      (each block has one owner: no atomics, deterministic).
#include "kernels.h"
#include "loss_tc.cuh"
#include "umma.cuh"

namespace socm {
namespace tc {

using namespace umma;

// scratch geometry: per 128-path tile, 4 quarters x (gfb + rfb) feature blocks of 4 KB
struct K2bGeom {
  int gfb, rfb;      // feature blocks of G^T (ceil(nrows/32)) and R^T (ceil(kdim/32))
  int n_tiles;       // ceil(B / 128)
  int nrows, kdim, d;
};
__host__ __device__ inline int64_t k2b_quarter_bytes(const K2bGeom& g) { return (int64_t)(g.gfb + g.rfb) * FB_BYTES; }

// src[B][ld] (row-major) -> feature blocks [tile][quarter][fb0 + f/32][f%32 rows of 32 paths, swizzled]
__global__ void __launch_bounds__(256) transpose_pack_kernel(const float* __restrict__ src, int B, int ld, int nfeat,
                                                             int n_fb, int fb0, int64_t quarter_bytes,
                                                             unsigned char* __restrict__ scratch) {
  __shared__ float tile[32][33];
  const int n_q = (B + 31) / 32;                       // path quarters
  const int64_t total = (int64_t)n_q * n_fb;
  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    const int qi = (int)(w / n_fb), fb = (int)(w - (int64_t)qi * n_fb);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
    for (int r = ty; r < 32; r += 8) {                       // coalesced read: 32 features of path (32 qi + r)
      const int m = 32 * qi + r, f = 32 * fb + tx;
      tile[r][tx] = (m < B && f < nfeat) ? __ldg(src + (size_t)m * ld + f) : 0.f;
    }
    __syncthreads();
    unsigned char* blk = scratch + (size_t)(qi >> 2) * 4 * quarter_bytes + (size_t)(qi & 3) * quarter_bytes +
                         (size_t)(fb0 + fb) * FB_BYTES;
    for (int f = ty; f < 32; f += 8)                         // coalesced write: one 128-byte row per warp
      *reinterpret_cast<float*>(blk + fb_off(tx, f)) = tile[tx][f];
    __syncthreads();
  }
}

constexpr int KB_STAGES = 2;
constexpr int KB_RAW = 16384 + 32768;        // A (128 features) | B (256 features), 32 paths each
constexpr int KB_STAGE_BYTES = 2 * KB_RAW;   // raw (-> hi in place) + lo
constexpr int KB_SMEM = KB_STAGES * KB_STAGE_BYTES + 1024 + 256;
constexpr int KB_NT = 320;                   // warps 0-3 split + flush, 4 MMA, 5 producer, 6-9 split
constexpr int KB_SEG = 32;                   // stages (of 4 K steps x 3 MMAs) per TMEM accumulation segment
constexpr uint64_t KB_SW128 = 2ull << 61;

__device__ __forceinline__ uint64_t kb_desc(uint32_t saddr) { return smem_desc(saddr, 16, 1024) | KB_SW128; }

// blocks[i] = rb | (cb << 16): rows [128 rb, +128) of dL, columns [256 cb, +256)
__global__ void __launch_bounds__(KB_NT, 1)
    target_bwd_tc_kernel(const unsigned char* __restrict__ scratch, K2bGeom g, const int* __restrict__ blocks, int n_blocks,
                         float* __restrict__ dL, int ldr) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + KB_STAGES * KB_STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + KB_STAGES;
  uint64_t* split = bars + 2 * KB_STAGES;
  uint64_t* acc_full = bars + 3 * KB_STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_stage_blk = g.n_tiles * 4;                       // stages (path quarters) per block
  const int n_seg = (n_stage_blk + KB_SEG - 1) / KB_SEG;
  const int64_t qbytes = k2b_quarter_bytes(g);

  if (tid == 0) {
    for (int s = 0; s < KB_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
      mbar_init(&split[s], 8);
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
          const uint32_t s = it % KB_STAGES;
          mbar_wait(&empty[s], ((it / KB_STAGES) & 1) ^ 1);
          unsigned char* st = smem + s * KB_STAGE_BYTES;
          const unsigned char* qb = scratch + (size_t)q * qbytes;
          mbar_expect_tx(&full[s], KB_RAW);
          bulk_g2s(st, qb + (size_t)(4 * rb) * FB_BYTES, 16384, &full[s]);
          bulk_g2s(st + 16384, qb + (size_t)(g.gfb + 8 * cb) * FB_BYTES, 32768, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 4) {
    // ===================================================== MMA issue
    constexpr uint32_t idesc = idesc_tf32(128, 256, 0, 0);
    uint32_t it = 0, ia = 0;
    for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
      for (int sg = 0; sg < n_seg; ++sg, ++ia) {
        const uint32_t a = ia & 1;
        mbar_wait(&acc_empty[a], ((ia >> 1) & 1) ^ 1);
        fence_after_sync();
        const int q1 = (sg + 1) * KB_SEG < n_stage_blk ? (sg + 1) * KB_SEG : n_stage_blk;
        for (int q = sg * KB_SEG; q < q1; ++q, ++it) {
          const uint32_t s = it % KB_STAGES;
          mbar_wait(&split[s], (it / KB_STAGES) & 1);
          fence_after_sync();
          const uint32_t st = smem_addr(smem + s * KB_STAGE_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = kb_desc(st + ks * 32), al = kb_desc(st + KB_RAW + ks * 32);
              const uint64_t bd = kb_desc(st + 16384 + ks * 32), bl = kb_desc(st + KB_RAW + 16384 + ks * 32);
              mma_ss(tm + a * 256, ad, bd, idesc, (q == sg * KB_SEG && ks == 0) ? 0u : 1u);
              mma_ss(tm + a * 256, al, bd, idesc, 1u);
              mma_ss(tm + a * 256, ad, bl, idesc, 1u);
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
    // ===================================================== warps 0-3: hi/lo split of every stage, then the flush
    const bool flusher = warp < 4;
    const int sid = flusher ? tid : tid - 64;       // 0..255 among the split threads
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t it = 0, ia = 0;
    for (int bi = blockIdx.x; bi < n_blocks; bi += gridDim.x) {
      const int rb = blocks[bi] & 0xFFFF, cb = blocks[bi] >> 16;
      const int n = 128 * rb + (tid & 127);         // row of dL owned by this (flusher) thread
      const int k_lo = 2 * (n / g.d) * g.d;         // columns left of it are structurally zero
      for (int sg = 0; sg < n_seg; ++sg, ++ia) {
        const int q1 = (sg + 1) * KB_SEG < n_stage_blk ? (sg + 1) * KB_SEG : n_stage_blk;
        for (int q = sg * KB_SEG; q < q1; ++q, ++it) {
          const uint32_t s = it % KB_STAGES;
          mbar_wait(&full[s], (it / KB_STAGES) & 1);
          float4* raw = reinterpret_cast<float4*>(smem + s * KB_STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(smem + s * KB_STAGE_BYTES + KB_RAW);
          for (int j = sid; j < KB_RAW / 16; j += 256) {
            const float4 x = raw[j];
            float4 hi, y;
            hi.x = tf32_rn(x.x); hi.y = tf32_rn(x.y); hi.z = tf32_rn(x.z); hi.w = tf32_rn(x.w);
            y.x = x.x - hi.x; y.y = x.y - hi.y; y.z = x.z - hi.z; y.w = x.w - hi.w;
            raw[j] = hi;
            lo[j] = y;
          }
          fence_async_smem();
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&split[s]);
        }
        if (!flusher) continue;
        // ---- flush this segment's partial sums: dL[n][256 cb + c] += acc
        const uint32_t a = ia & 1;
        mbar_wait(&acc_full[a], (ia >> 1) & 1);
        fence_after_sync();
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          float v[32];
          tmem_ld32(lane_t + a * 256 + c0, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          if (n < g.nrows) {
            // fire-and-forget adds (this thread is the only writer of its row segment): a read-modify-write
            // here kept the split warps waiting on global loads once per segment
            float* dst = dL + (size_t)n * ldr + 256 * cb + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int k = 256 * cb + c0 + j;
              if (k >= k_lo && k + 3 < g.kdim) {
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
              } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  if (k + jj < g.kdim && k + jj >= k_lo)
                    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + j + jj), "f"(v[j + jj]) : "memory");
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

}  // namespace tc
}  // namespace socm

using namespace socm;

static tc::K2bGeom k2b_geom(int B, int K, int d) {
  tc::K2bGeom g;
  g.nrows = (K + 1) * d;
  g.kdim = (2 * K + 1) * d;
  g.d = d;
  g.gfb = ((g.nrows + 127) / 128) * 4;   // padded to whole A blocks (128 features)
  g.rfb = ((g.kdim + 255) / 256) * 8;    // padded to whole B blocks (256 features)
  g.n_tiles = (B + 127) / 128;
  return g;
}

extern "C" int64_t socm_target_gemm_bwd_tc_workspace_bytes(int32_t B, int32_t K, int32_t d) {
  if (B < 0 || K < 1 || d < 1) return -1;
  const tc::K2bGeom g = k2b_geom(B, K, d);
  const int64_t n_pairs = (int64_t)(g.gfb / 4) * (g.rfb / 8);
  return (int64_t)g.n_tiles * 4 * tc::k2b_quarter_bytes(g) + n_pairs * 4 + 4096;
}

extern "C" int socm_target_gemm_bwd_tc_f32(const float* G, const float* R, int32_t B, int32_t K, int32_t d,
                                           int32_t ldr, int32_t ldt, float* dL, int32_t accumulate, void* workspace,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SOCM_CHECK_ARG(G && R && dL && workspace, "required pointer is NULL");
  SOCM_CHECK_ARG(d >= 1 && d <= SOCM_MAX_DIM && K >= 1, "bad sizes");
  SOCM_CHECK_ARG(ldr >= (2 * K + 1) * d && ldt >= (K + 1) * d, "bad pitches ldr=%d ldt=%d", ldr, ldt);
  const tc::K2bGeom g = k2b_geom(B, K, d);
  if (!(accumulate & 1)) SOCM_CUDA(cudaMemsetAsync(dL, 0, (size_t)g.nrows * ldr * sizeof(float), stream));
  if (B == 0) return SOCM_OK;
  // engine: fp16 hi / lo planes on kind::f16 (target_bwd_h.cu) unless SOCM_TARGET_BWD_TF32 or SOCM_F16=0 ask for 3xTF32
  const bool want_f16 = (accumulate & SOCM_TARGET_BWD_F16) || (f16_default() != 0 && !(accumulate & SOCM_TARGET_BWD_TF32));
  if (want_f16 && ldr % 4 == 0 && ldt % 4 == 0)
    return hx::launch_target_bwd_h(G, R, B, K, d, ldr, ldt, dL, workspace, stream);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  ws += (1024 - (reinterpret_cast<uintptr_t>(ws) & 1023)) & 1023;
  const int64_t qbytes = tc::k2b_quarter_bytes(g);
  unsigned char* scratch = ws;
  int* blocks_dev = reinterpret_cast<int*>(ws + (size_t)g.n_tiles * 4 * qbytes);
  // block list: [128 rows rb] x [256 columns cb] with a non-empty part right of the diagonal
  int blocks_host[4096];
  int n_blocks = 0;
  for (int rb = 0; rb < g.gfb / 4; ++rb)
    for (int cb = 0; cb < g.rfb / 8; ++cb) {
      const int i_min = (128 * rb) / d;                    // smallest grid time among the rows of the block
      if (256 * cb + 255 >= 2 * i_min * d && 256 * cb < g.kdim && 128 * rb < g.nrows && n_blocks < 4096)
        blocks_host[n_blocks++] = rb | (cb << 16);
    }
  SOCM_CHECK_ARG(n_blocks < 4096, "too many blocks for the tcgen05 target-backward kernel");
  SOCM_CUDA(cudaMemcpyAsync(blocks_dev, blocks_host, n_blocks * sizeof(int), cudaMemcpyHostToDevice, stream));
  const int n_q = (B + 31) / 32;
  // quarters beyond B inside the last tile must hold zeros: the kernel streams whole tiles
  if (B % 128) SOCM_CUDA(cudaMemsetAsync(scratch + (size_t)(g.n_tiles - 1) * 4 * qbytes, 0, (size_t)4 * qbytes, stream));
  tc::transpose_pack_kernel<<<2368, 256, 0, stream>>>(G, B, ldt, g.nrows, g.gfb, 0, qbytes, scratch);
  SOCM_LAUNCH_CHECK();
  tc::transpose_pack_kernel<<<2368, 256, 0, stream>>>(R, B, ldr, g.kdim, g.rfb, g.gfb, qbytes, scratch);
  SOCM_LAUNCH_CHECK();
  (void)n_q;
  SOCM_CUDA(cudaFuncSetAttribute(tc::target_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::KB_SMEM));
  const int grid = n_blocks < sm_count() ? n_blocks : sm_count();
  tc::target_bwd_tc_kernel<<<grid, tc::KB_NT, tc::KB_SMEM, stream>>>(scratch, g, blocks_dev, n_blocks, dL, ldr);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
