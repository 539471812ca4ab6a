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
#include "wgrad_tables.cuh"

namespace socm {
namespace tc {

using namespace umma;

// K-major operand, 128-byte rows (32 points per feature), 8-row swizzle atoms of 1 KB stacked along the features
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) { return smem_desc(saddr, 16, 1024) | DESC_SW128; }

__device__ __forceinline__ void red_add(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(WG_NT, 1) wgrad_tc_kernel(const unsigned char* __restrict__ scratch, int n_tiles,
                                                            int d, float* __restrict__ grad, float* __restrict__ aux) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // stage bases must be 1 KB aligned in the shared window: the operand swizzle uses address bits 7-9
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
    // ===================================================== producer: one pipeline stage = (stage desc, tile, quarter)
    // Only two stages fit in shared memory (raw + lo copies), far too little to cover the HBM latency, so
    // the producer runs an L2 prefetch (cp.async.bulk.prefetch.L2) WG_PREFETCH stages ahead of the copies.
    if (elect_one()) {
      struct Cursor {
        int pass, ti, q, l;  // pass, index into this CTA's tiles, quarter, stage desc of the pass
      };
      auto advance = [&](Cursor& c) {
        if (++c.l == c_pass_stage[c.pass + 1]) {
          c.l = c_pass_stage[c.pass];
          if (++c.q == 4) {
            c.q = 0;
            if (++c.ti == my_tiles) {
              c.ti = 0;
              ++c.pass;
              c.l = c.pass < N_PASS ? c_pass_stage[c.pass] : 0;
            }
          }
        }
      };
      auto prefetch = [&](const Cursor& c) {
        if (c.pass >= N_PASS) return;
        const unsigned char* qb = scratch + (size_t)(blockIdx.x + c.ti * gridDim.x) * TILE_BYTES + (size_t)c.q * QUARTER_BYTES;
        const int nl = c_stage[c.l].n_load;
        for (int j = 0; j < nl; ++j) {
          const Load ld = c_stage[c.l].ld[j];
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(qb + (size_t)ld.fb * FB_BYTES), "r"(ld.bytes) : "memory");
        }
      };
      Cursor ahead{0, 0, 0, 0};
      if (my_tiles > 0)
        for (int i = 0; i < WG_PREFETCH; ++i) {
          prefetch(ahead);
          advance(ahead);
        }
      uint32_t it = 0;
      for (int pass = 0; pass < N_PASS && my_tiles > 0; ++pass) {
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
          const unsigned char* tile = scratch + (size_t)t * TILE_BYTES;
          for (int q = 0; q < 4; ++q) {
            const unsigned char* qb = tile + (size_t)q * QUARTER_BYTES;
            for (int l = c_pass_stage[pass]; l < c_pass_stage[pass + 1]; ++l, ++it) {
              prefetch(ahead);
              advance(ahead);
              const uint32_t s = it % WG_STAGES;
              mbar_wait(&empty[s], ((it / WG_STAGES) & 1) ^ 1);
              unsigned char* st = smem + s * WG_STAGE_BYTES;
              const int nl = c_stage[l].n_load;
              uint32_t total = 0;
              for (int j = 0; j < nl; ++j) total += (uint32_t)c_stage[l].ld[j].bytes;
              mbar_expect_tx(&full[s], total);
              for (int j = 0; j < nl; ++j) {
                const Load ld = c_stage[l].ld[j];
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
      mbar_wait(acc_empty, (nf & 1) ^ 1);  // the flush warps have drained the previous segment
      fence_after_sync();
      for (int i = t0 * 4; i < t1 * 4; ++i) {
        for (int l = c_pass_stage[pass]; l < c_pass_stage[pass + 1]; ++l, ++it) {
          const uint32_t first = i == t0 * 4 ? 1u : 0u;
          const uint32_t s = it % WG_STAGES;
          mbar_wait(&split[s], (it / WG_STAGES) & 1);
          fence_after_sync();
          const uint32_t st = smem_addr(smem + s * WG_STAGE_BYTES);
          if (elect_one()) {
            const int nm = c_stage[l].n_mma;
            for (int j = 0; j < nm; ++j) {
              const Mma m = c_stage[l].mma[j];
              const uint32_t idesc = idesc_tf32(128, m.N, 0, 0);
              const uint32_t dcol = tm + (uint32_t)m.col;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ad = mn_desc(st + m.a_off + ks * 32), al = mn_desc(st + WG_LO + m.a_off + ks * 32);
                const uint64_t bd = mn_desc(st + m.b_off + ks * 32), bl = mn_desc(st + WG_LO + m.b_off + ks * 32);
                mma_ss(dcol, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
                mma_ss(dcol, al, bd, idesc, 1u);
                mma_ss(dcol, ad, bl, idesc, 1u);
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
    // ===================================================== split (8 warps) + flush (warps 0-3: TMEM lane r <-> output row)
    const bool flusher = warp < 4;
    const int sid = flusher ? tid : tid - 64;  // 0..255 among the split threads
    const uint32_t lane_t = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int r = tid & 127;
    uint32_t it = 0, nf = 0;
    for (int pass = 0; pass < N_PASS; ++pass)
    for (int t0 = 0; t0 < my_tiles; t0 += WG_SEG, ++nf) {
      const int t1 = t0 + WG_SEG < my_tiles ? t0 + WG_SEG : my_tiles;
      // ---- split every stage of this segment: the MMA reads hi = trunc(x) from the raw copy, lo = x - hi
      for (int i = t0 * 4; i < t1 * 4; ++i)
        for (int l = c_pass_stage[pass]; l < c_pass_stage[pass + 1]; ++l, ++it) {
          const uint32_t s = it % WG_STAGES;
          mbar_wait(&full[s], (it / WG_STAGES) & 1);
          float4* raw = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES);
          float4* lo = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + WG_LO);
          // The MMA reads the raw fp32 value truncated to tf32 (hi = trunc(x)), so only lo = x - trunc(x) has to
          // be written: x = hi + lo exactly, lo < 2^-10 |x| and the hardware keeps 11 bits of it, i.e. the pair
          // represents x to 2^-20 (one-sided).  A round-to-nearest hi would halve that but costs a second
          // shared-memory write of every operand, and this kernel is bound by shared-memory bandwidth.
          // Four independent 16-byte loads are in flight per thread (one at a time left the warps waiting on
          // the shared-memory latency for a fifth of the kernel).
          const int nl = c_stage[l].n_load;
          for (int jl = 0; jl < nl; ++jl) {
            const int f4_begin = c_stage[l].ld[jl].dst / 16, n_f4 = c_stage[l].ld[jl].bytes / 16;
            for (int j0 = sid; j0 < n_f4; j0 += 1024) {
              float4 x[4];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (j0 + 256 * u < n_f4) x[u] = raw[f4_begin + j0 + 256 * u];
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (j0 + 256 * u < n_f4) {
                  float4 y;
                  y.x = x[u].x - tf32_hi(x[u].x); y.y = x[u].y - tf32_hi(x[u].y);
                  y.z = x[u].z - tf32_hi(x[u].z); y.w = x[u].w - tf32_hi(x[u].w);
                  lo[f4_begin + j0 + 256 * u] = y;
                }
            }
          }
          fence_async_smem();
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&split[s]);
        }
      if (!flusher) continue;
      mbar_wait(acc_full, nf & 1);
      fence_after_sync();
      for (int o = c_pass_out[pass]; o < c_pass_out[pass + 1]; ++o) {
        const OutDesc od = c_out[o];
        const int row = od.row0 + r;
        if (od.kind == OUT_BIAS) {  // column ONES_FEATURE of (rows x XIN)
          float v[16];
          tmem_ld16(lane_t + od.col + 16, reinterpret_cast<uint32_t*>(v));
          tmem_wait_ld();
          if (r < od.rows) red_add(grad + go.b[od.layer] + row, v[ONES_FEATURE - 16]);
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
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
            } else {  // the flat gradient offsets are not 16-byte aligned for every d
#pragma unroll
              for (int j = 0; j < 16; ++j) red_add(dst + j, v[j]);
            }
          } else if (od.kind == OUT_TRANSPOSED) {
            // D[in = row][out = c]: weight (layer) is [out][in_total]; up_0: out < d, in_total = 256; down_2: in_total = 128
            const int in_total = od.layer == 8 ? 256 : 128;
            const int n_out = od.layer == 8 ? d : 64;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < n_out) red_add(grad + go.w[od.layer] + (size_t)(c0 + j) * in_total + row, v[j]);
          } else if (od.kind == OUT_AUX_S) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < d) red_add(aux + AUX_S + (c0 + j) * 256 + row, v[j]);
          } else if (od.kind == OUT_XIN) {
            // down_0: D[out = row][k]: k <= d -> weight, k = ONES_FEATURE -> bias
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int k = c0 + j;
              if (k <= d) red_add(grad + go.w[0] + (size_t)row * (d + 1) + k, v[j]);
              if (k == ONES_FEATURE) red_add(grad + go.b[0] + row, v[j]);
            }
          } else {  // OUT_SMALL: rows 0..31: d_y0[j]; rows 32..63: d_o0[j]
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
