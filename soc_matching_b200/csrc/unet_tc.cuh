// unet_tc.cuh -- tcgen05 (5th-gen tensor core) engine for the default FullyConnectedUNet
// (hdims [256,128,64], models.py:202-242) at fp32-class accuracy ("3xTF32").
//
// Geometry.  One CTA owns a tile of TP = 128 trajectory points = the 128 lanes of tensor memory:
// every GEMM is  D[128 points x N out-features] (+)= A[128 x K] * W^T[K x N]  with M = 128.
//   * the accumulators D live in TMEM (fp32, one column per output feature);
//   * the activations (A operand) are written back to TMEM by the epilogue threads
//     (thread e <-> point e <-> TMEM lane e: no cross-thread traffic) and read by the next layer
//     straight from TMEM (tcgen05.mma with A in tensor memory);
//   * the two 256-wide activations (r1 = relu(down_0 x) feeding down_1 and the folded Wc, y1 = relu(up_1 o2)
//     feeding the folded up_0) do not fit next to the accumulators, so they are produced in 32-feature
//     chunks into a double-buffered shared-memory A operand and consumed chunk by chunk;
//   * the weights (B operand, K-major = nn.Linear's own [out][in] layout) are streamed from L2
//     every step as a fixed tape of 32 KB slots by cp.async.bulk into a 3-stage ring
//     (0.8 MB per 128 points and evaluation; measured 132 GB/s per SM, see scripts/umma_probe.cu).
//
// Precision.  kind::tf32 truncates its fp32 inputs to 10 mantissa bits, so every product is
// issued three times on split operands  x = hi + lo,  hi = rn_tf32(x), lo = x - hi (exact):
//     a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo          (dropped a_lo*w_lo ~ 2^-24 |a w|)
// with fp32 accumulation in TMEM.  Weights are split once per call by pack_tc_kernel, the
// activations by the epilogue threads.
//
// TMEM column map (512 columns), forward pass:
//   [0,256)    D0 = down_0 pre-activation (chunk source)  ->  r2 hi|lo  ->  o2 hi|lo  ->  [0,NY) W_u0 y1 acc
//   [256,384)  D1 = down_1 acc  -> D2 = down_2 acc [256,320) -> D3 = up_2 acc
//   [384,400)  Wc r1 acc (joint MMA with down_1; read out before res_2 writes there)
//   [384,512)  D3R = res_2 acc: res_2 runs right behind down_2, under the r3 epilogue and up_2, instead of after a
//              separate relu(y2) pass; r3 (64 features, hi|lo = 64 KB) is an A operand in the two shared-memory
//              chunk buffers, which are idle between the r1 and the y1 chunks
//   [256,512)  D4 = up_1 acc (after D3 and D3R are dead)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace socm {
namespace tc {

constexpr int TP = 128;  // points per tile
constexpr int H0 = 256, H1 = 128, H2 = 64;
constexpr int MAIN_BYTES = 32768;  // main weight block of a tape slot: hi slab + lo slab
constexpr int SLOT_BYTES = 40960;  // tape slot stride / ring stage: main block + (down_1 slots) the folded Wc block
constexpr int NSTAGE = 3;
constexpr int CHUNK_F = 32;                   // features per shared-memory activation chunk
constexpr int CHUNK_HALF = TP * CHUNK_F * 4;  // bytes of the hi (or lo) part of a chunk
constexpr int CHUNK_BYTES = 2 * CHUNK_HALF;
constexpr int MAX_KIN = 24;                   // d + 1 rounded up to 8; the engine covers d <= 23
constexpr int WC_LAYER = 9;                   // pseudo layer index of the folded matrix Wc = W_up0 W_res1  [d][256]

__host__ __device__ inline int kin_of(int d) { return ((d + 1 + 7) / 8) * 8; }
__host__ __device__ inline int w0_slots(int d) { return kin_of(d) > 16 ? 2 : 1; }
// N of the folded last-layer MMAs (output features of up_0, padded to a legal N for M = 128)
__host__ __device__ inline int ny_of(int d) { return kin_of(d) <= 16 ? 16 : 32; }
__host__ __device__ inline int u0_slots(int d) { return ny_of(d) == 16 ? 1 : 2; }

// ---------------------------------------------------------------- folding res_1 into up_0
// models.py:239-241:  o1 = relu(up_1 o2) + res_1 r1,  out0 = relu(up_0 o1) + res_0 x.  There is no
// non-linearity between res_1 and up_0, so
//     up_0 o1 = W_u0 y1 + Wc r1 + bc,     Wc = W_u0 W_r1  [d x 256],  bc = W_u0 b_r1 + b_u0,  y1 = relu(up_1 o2)
// and the K = N = 256 res_1 GEMM (39 % of the network's FLOPs) is never evaluated per point: Wc r1
// rides on the r1 chunks that feed down_1 (an extra N = 16 MMA per chunk), and in the backward pass
//     d_r1 += Wc^T d_y0 (K = d),   dW_r1 = W_u0^T S,   dW_u0 = d_y0^T y1 + S W_r1^T + (sum d_y0) b_r1^T,   S = d_y0^T r1.
//
// ---------------------------------------------------------------- weight tapes
// forward slots: [down_0 x S] [down_1 x 8, each followed by its Wc block] [down_2 x 2] [res_2 x 4] [up_2 x 2]
//                [up_1 x 8] [up_0 x U: 8 K-chunks of 32, several per slot]
// backward slots (W^T blocks): [up_0^T x S] [up_1^T x 8] [res_2^T x 4] [up_2^T x 2] [down_2^T x 2] [down_1^T x 8] [Wc^T x S]
struct SlotDesc {
  int layer;  // index into socm_unet::w, or WC_LAYER
  int n0, N;  // rows of the block (N side of the MMA)
  int k0, Kc; // contraction range of the block (Kc multiple of 8)
  int ktot;   // row length of the weight matrix W[out][ktot]
  int transposed;  // 0: B[n][k] = W[n0+n][k0+k] (forward);  1: B[n][k] = W[k0+k][n0+n] (dgrad: W^T)
  int klim;   // contraction indices >= klim are zero padding
  int nlim;   // rows >= nlim are zero padding
  int slab_n; // rows of the slab this block is part of (0: N) -- the lo slab follows slab_n rows after the hi slab
  int slab_row;  // first row of the block inside the slab
};
struct PackItem {
  int slot, byte_off;
  SlotDesc sd;
};
constexpr int NOLIM = 1 << 30;
__host__ __device__ inline int fwd_slots(int d) { return w0_slots(d) + 24 + u0_slots(d); }
__host__ __device__ inline int bwd_slots(int d) { return 2 * w0_slots(d) + 24; }
__host__ __device__ inline int fwd_items(int d) { return w0_slots(d) + 40; }
__host__ __device__ inline int bwd_items(int d) { return bwd_slots(d); }
__host__ __device__ inline PackItem fwd_item(int d, int i) {
  const int S = w0_slots(d), kin = kin_of(d), NY = ny_of(d);
  if (i < S) return PackItem{i, 0, SlotDesc{0, i * (H0 / S), H0 / S, 0, kin, d + 1, 0, d + 1, NOLIM, 0, 0}};
  i -= S;
  // down_1 block and its Wc block form ONE B operand of H1 + NY rows: D1 and Wc r1 come out of a single MMA
  // (N = 144 / 160) into adjacent TMEM columns; a separate N = 16 MMA would re-read the whole A chunk from
  // shared memory and cost half as much as the N = 128 one
  if (i < 8) return PackItem{S + i, 0, SlotDesc{1, 0, H1, 32 * i, 32, H0, 0, H0, NOLIM, H1 + NY, 0}};
  i -= 8;
  if (i < 8) return PackItem{S + i, 0, SlotDesc{WC_LAYER, 0, NY, 32 * i, 32, H0, 0, H0, d, H1 + NY, H1}};
  i -= 8;
  if (i < 2) return PackItem{S + 8 + i, 0, SlotDesc{2, 0, H2, 64 * i, 64, H1, 0, H1, NOLIM, 0, 0}};
  i -= 2;
  // res_2 is issued right behind down_2 (both read r2; it has its own accumulator), up_2 after the r3 epilogue
  if (i < 4) return PackItem{S + 10 + i, 0, SlotDesc{5, 0, H1, 32 * i, 32, H1, 0, H1, NOLIM, 0, 0}};
  i -= 4;
  if (i < 2) return PackItem{S + 14 + i, 0, SlotDesc{6, 0, H1, 32 * i, 32, H2, 0, H2, NOLIM, 0, 0}};
  i -= 2;
  if (i < 8) return PackItem{S + 16 + i, 0, SlotDesc{7, 0, H0, 16 * i, 16, H1, 0, H1, NOLIM, 0, 0}};
  i -= 8;
  const int bps = MAIN_BYTES / (NY * 256);  // up_0 K-chunks per slot (8 or 4)
  return PackItem{S + 24 + i / bps, (i % bps) * (NY * 256), SlotDesc{8, 0, NY, 32 * i, 32, H0, 0, H0, d, 0, 0}};
}
__host__ __device__ inline PackItem bwd_item(int d, int s) {
  const int S = w0_slots(d), kin = kin_of(d);
  const int slot = s;
  if (s < S) return PackItem{slot, 0, SlotDesc{8, s * (H0 / S), H0 / S, 0, kin, H0, 1, d, NOLIM, 0, 0}};  // d_o1[f] = sum_j d_y0[j] W_u0[j][f]
  s -= S;
  if (s < 8) return PackItem{slot, 0, SlotDesc{7, 0, H1, 32 * s, 32, H1, 1, H0, NOLIM, 0, 0}};   // d_o2[c] = sum_n d_y1[n] W_u1[n][c]
  s -= 8;
  if (s < 4) return PackItem{slot, 0, SlotDesc{5, 0, H1, 32 * s, 32, H1, 1, H1, NOLIM, 0, 0}};   // d_r2[c] = sum_n d_o2[n] W_r2[n][c]
  s -= 4;
  if (s < 2) return PackItem{slot, 0, SlotDesc{6, 0, H2, 64 * s, 64, H2, 1, H1, NOLIM, 0, 0}};   // d_r3[c] = sum_n d_y2[n] W_u2[n][c]
  s -= 2;
  if (s < 2) return PackItem{slot, 0, SlotDesc{2, 0, H1, 32 * s, 32, H1, 1, H2, NOLIM, 0, 0}};   // d_r2[c] += sum_n d_z3[n] W_d2[n][c]
  s -= 2;
  if (s < 8) return PackItem{slot, 0, SlotDesc{1, 0, H0, 16 * s, 16, H0, 1, H1, NOLIM, 0, 0}};   // d_r1[c] = sum_n d_z2[n] W_d1[n][c]
  s -= 8;
  return PackItem{slot, 0, SlotDesc{WC_LAYER, s * (H0 / S), H0 / S, 0, kin, H0, 1, d, NOLIM, 0, 0}};  // d_r1[g] += sum_j d_y0[j] Wc[j][g]
}
// bytes the producer copies for a slot
__host__ __device__ inline uint32_t fwd_slot_bytes(int d, int s) {
  const int S = w0_slots(d);
  if (s < S) return (uint32_t)(2 * (H0 / S) * kin_of(d) * 4);
  if (s < S + 8) return (uint32_t)(MAIN_BYTES + 2 * ny_of(d) * 32 * 4);
  return (uint32_t)MAIN_BYTES;
}
__host__ __device__ inline uint32_t bwd_slot_bytes(int d, int s) {
  const int S = w0_slots(d);
  if (s < S || s >= S + 24) return (uint32_t)(2 * (H0 / S) * kin_of(d) * 4);
  return (uint32_t)MAIN_BYTES;
}

// ---------------------------------------------------------------- small block (floats, fp32, read from shared memory)
struct SmallTc {
  int b_d0, b_d1, b_d2, b_u2, b_r2, b_u1;
  int bc;    // [kin]  folded bias W_u0 b_r1 + b_u0
  int r0;    // res_0 [kin][kin]   (zero padded rows and columns; column 0 multiplies t)
  int b_r0;  // [kin]
  int total;
};
__host__ __device__ inline SmallTc small_tc(int d) {
  const int kin = kin_of(d);
  SmallTc o;
  int p = 0;
  o.b_d0 = p; p += H0;
  o.b_d1 = p; p += H1;
  o.b_d2 = p; p += H2;
  o.b_u2 = p; p += H1;
  o.b_r2 = p; p += H1;
  o.b_u1 = p; p += H0;
  o.bc = p; p += kin;
  o.r0 = p; p += kin * kin;
  o.b_r0 = p; p += kin;
  o.total = ((p + 3) / 4) * 4;
  return o;
}
// workspace: [tape: (fwd_slots (+ bwd_slots)) x SLOT_BYTES][small block][Wc: 32 x 256 floats]
constexpr int WC_FLOATS = 32 * H0;
__host__ __device__ inline int64_t tc_tape_bytes(int d, bool with_bwd) {
  return (int64_t)(fwd_slots(d) + (with_bwd ? bwd_slots(d) : 0)) * SLOT_BYTES;
}
__host__ __device__ inline int64_t tc_workspace_bytes(int d, bool with_bwd = false) {
  return tc_tape_bytes(d, with_bwd) + (int64_t)small_tc(d).total * 4 + (int64_t)WC_FLOATS * 4;
}

__host__ __device__ inline float* tc_small_ptr(unsigned char* tape, int d, bool with_bwd) {
  return reinterpret_cast<float*>(tape + tc_tape_bytes(d, with_bwd));
}
__host__ __device__ inline float* tc_wc_ptr(unsigned char* tape, int d, bool with_bwd) {
  return tc_small_ptr(tape, d, with_bwd) + small_tc(d).total;
}

// canonical no-swizzle K-major offsets (bytes)
// weight slab [N rows][Kc cols]: core matrices contiguous along k, then along n
__host__ __device__ inline int wslab_off(int n, int k, int Kc) {
  return (n % 8) * 16 + (k % 4) * 4 + (n / 8) * (Kc * 32) + (k / 4) * 128;
}
// activation operand [128 points][F cols]: core matrices contiguous along the points, then along f
__host__ __device__ inline int act_off(int p, int f) { return (p % 8) * 16 + (f % 4) * 4 + (p / 8) * 128 + (f / 4) * 2048; }
constexpr uint32_t ACT_LBO = 2048, ACT_SBO = 128, ACT_KSTEP = 4096;  // descriptor strides of act_off
constexpr uint32_t W_LBO = 128, W_KSTEP = 256;                       // weight slab: SBO = Kc * 32

// ---------------------------------------------------------------- MMA issue helpers (elected thread)
// one weight block against an A operand in TMEM (hi columns a_hi.., lo columns a_lo..)
template <int N, int Kc>
__device__ __forceinline__ void issue_block_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_smem,
                                               bool fresh) {
  constexpr uint32_t id = umma::idesc_tf32(TP, N, 0, 0);
  constexpr uint32_t slab = (uint32_t)N * Kc * 4;
#pragma unroll
  for (int ks = 0; ks < Kc / 8; ++ks) {
    const uint64_t bh = umma::smem_desc(b_smem + ks * W_KSTEP, W_LBO, Kc * 32);
    const uint64_t bl = umma::smem_desc(b_smem + slab + ks * W_KSTEP, W_LBO, Kc * 32);
    umma::mma_ts(d_tmem, a_hi + ks * 8, bh, id, (fresh && ks == 0) ? 0u : 1u);
    umma::mma_ts(d_tmem, a_lo + ks * 8, bh, id, 1u);
    umma::mma_ts(d_tmem, a_hi + ks * 8, bl, id, 1u);
  }
}
// one weight block against an A operand in shared memory (act_off layout; lo part at a_smem + a_lo_off)
template <int N, int Kc>
__device__ __forceinline__ void issue_block_ss(uint32_t d_tmem, uint32_t a_smem, uint32_t a_lo_off, uint32_t b_smem,
                                               bool fresh) {
  constexpr uint32_t id = umma::idesc_tf32(TP, N, 0, 0);
  constexpr uint32_t slab = (uint32_t)N * Kc * 4;
#pragma unroll
  for (int ks = 0; ks < Kc / 8; ++ks) {
    const uint64_t bh = umma::smem_desc(b_smem + ks * W_KSTEP, W_LBO, Kc * 32);
    const uint64_t bl = umma::smem_desc(b_smem + slab + ks * W_KSTEP, W_LBO, Kc * 32);
    const uint64_t ah = umma::smem_desc(a_smem + ks * ACT_KSTEP, ACT_LBO, ACT_SBO);
    const uint64_t al = umma::smem_desc(a_smem + a_lo_off + ks * ACT_KSTEP, ACT_LBO, ACT_SBO);
    umma::mma_ss(d_tmem, ah, bh, id, (fresh && ks == 0) ? 0u : 1u);
    umma::mma_ss(d_tmem, al, bh, id, 1u);
    umma::mma_ss(d_tmem, ah, bl, id, 1u);
  }
}

// ---------------------------------------------------------------- epilogue helpers (thread <-> TMEM lane)
// v[j] = max(v[j] + bias[j], 0) on 32 columns; bias in shared memory (broadcast reads)
__device__ __forceinline__ void bias_relu32(float* v, const float* __restrict__ bias) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    v[j] = fmaxf(v[j] + b.x, 0.f);
    v[j + 1] = fmaxf(v[j + 1] + b.y, 0.f);
    v[j + 2] = fmaxf(v[j + 2] + b.z, 0.f);
    v[j + 3] = fmaxf(v[j + 3] + b.w, 0.f);
  }
}
__device__ __forceinline__ void bias32(float* v, const float* __restrict__ bias) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    v[j] += b.x;
    v[j + 1] += b.y;
    v[j + 2] += b.z;
    v[j + 3] += b.w;
  }
}
// TMEM A operand: hi -> columns a_hi.., lo -> columns a_lo..   (32 columns)
__device__ __forceinline__ void store_split32(uint32_t a_hi, uint32_t a_lo, const float* v) {
  uint32_t h[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = __float_as_uint(umma::tf32_rn(v[j]));
  umma::tmem_st32(a_hi, h);
#pragma unroll
  for (int j = 0; j < 32; ++j) h[j] = __float_as_uint(v[j] - __uint_as_float(h[j]));
  umma::tmem_st32(a_lo, h);
}
__device__ __forceinline__ void store_split16(uint32_t a_hi, uint32_t a_lo, const float* v) {
  uint32_t h[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) h[j] = __float_as_uint(umma::tf32_rn(v[j]));
  umma::tmem_st16(a_hi, h);
#pragma unroll
  for (int j = 0; j < 16; ++j) h[j] = __float_as_uint(v[j] - __uint_as_float(h[j]));
  umma::tmem_st16(a_lo, h);
}
// shared-memory A operand chunk (32 features of point p): hi at chunk, lo at chunk + CHUNK_HALF
__device__ __forceinline__ void store_chunk32(unsigned char* chunk, int p, const float* v) {
  unsigned char* base = chunk + (p % 8) * 16 + (p / 8) * 128;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    float4 h, l;
    h.x = umma::tf32_rn(v[4 * g]);
    h.y = umma::tf32_rn(v[4 * g + 1]);
    h.z = umma::tf32_rn(v[4 * g + 2]);
    h.w = umma::tf32_rn(v[4 * g + 3]);
    l.x = v[4 * g] - h.x;
    l.y = v[4 * g + 1] - h.y;
    l.z = v[4 * g + 2] - h.z;
    l.w = v[4 * g + 3] - h.w;
    *reinterpret_cast<float4*>(base + g * 2048) = h;
    *reinterpret_cast<float4*>(base + CHUNK_HALF + g * 2048) = l;
  }
}


// one arrival per warp (the barrier counts warps): every lane's writes are ordered before it by __syncwarp
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}
__device__ __forceinline__ void e_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 epilogue warps

// 16-column variants (used when two warps share a TMEM lane quarter and split the columns)
__device__ __forceinline__ void bias_relu16(float* v, const float* __restrict__ bias) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias + j);
    v[j] = fmaxf(v[j] + b.x, 0.f);
    v[j + 1] = fmaxf(v[j + 1] + b.y, 0.f);
    v[j + 2] = fmaxf(v[j + 2] + b.z, 0.f);
    v[j + 3] = fmaxf(v[j + 3] + b.w, 0.f);
  }
}
// 16 features (4 column groups starting at group g0) of point p into a shared-memory A chunk
__device__ __forceinline__ void store_chunk16(unsigned char* chunk, int p, int g0, const float* v) {
  unsigned char* base = chunk + (p % 8) * 16 + (p / 8) * 128 + g0 * 2048;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float4 h, l;
    h.x = umma::tf32_rn(v[4 * g]);
    h.y = umma::tf32_rn(v[4 * g + 1]);
    h.z = umma::tf32_rn(v[4 * g + 2]);
    h.w = umma::tf32_rn(v[4 * g + 3]);
    l.x = v[4 * g] - h.x;
    l.y = v[4 * g + 1] - h.y;
    l.z = v[4 * g + 2] - h.z;
    l.w = v[4 * g + 3] - h.w;
    *reinterpret_cast<float4*>(base + g * 2048) = h;
    *reinterpret_cast<float4*>(base + CHUNK_HALF + g * 2048) = l;
  }
}

}  // namespace tc
}  // namespace socm
