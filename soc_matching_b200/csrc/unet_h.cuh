// unet_h.cuh -- second-generation tcgen05 engine for the default FullyConnectedUNet (hdims [256,128,64], d <= 15):
// fp16 operands split hi / lo ("2xFP16", 22 mantissa bits), fp32 accumulation in tensor memory, TWO CTAs PER SM.
//
// Why.  The 3xTF32 engine of unet_tc.cuh needs all 512 TMEM columns for one 128-point tile (a 128-wide activation is
// 256 columns as a tf32 hi|lo A operand), so the tensor pipe idles whenever that tile's epilogue threads work: 41-51 %
// tensor-pipe activity (profiles/r1f_kernels_ncu.md).  With kind::f16
//   * an MMA covers K = 16 instead of 8 at the same N/2 cycles: half the tensor time for the same 3-product scheme
//       a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo ,   hi = rn_f16(x), lo = rn_f16(x - hi)      (|x - hi - lo| <= 2^-23 |x|)
//     and half the accumulation steps (the tensor core adds into its fp32 accumulator with truncation: fewer steps,
//     less bias: profiles/r2_f16_probe.log);
//   * an A operand holds two halfs per TMEM column, weights and activation chunks are half as large in shared memory:
//     a tile fits 256 TMEM columns and ~105 KB of shared memory, so two CTAs are resident per SM and the hardware
//     interleaves their MMA streams -- one tile's epilogue runs under the other tile's MMAs.
// fp16 has a narrow exponent range, so every operand is scaled by a power of two (exact): weights per layer from their
// max |w| (scaled max in [512, 1024)), activations per layer from a calibration pass (fp32 forward on sample points,
// scaled max ~ 64, i.e. 2^10 of head room; conversions saturate instead of overflowing).  The hi part keeps 11 bits
// down to 6e-5, the lo part adds 11 more for |x| >= 0.125 and ~2^-25 absolute below: relative to the layer's largest
// values that is 2^-30, far below the 2^-23 of the scheme itself.  The epilogue un-scales with one FMA per element.
//
// TMEM column map of a tile (256 columns), forward pass:
//   [0,64)     two 32-column "piece" buffers: down_0 and up_1 produce their 256 outputs in 8 pieces of 32 features that
//              the epilogue threads turn into shared-memory A chunks;  in between: D2 = down_2 acc, later res_2 acc a
//   [64,192)   D1 = down_1 acc -> r2 (A operand, converted IN PLACE: 16 fp32 columns become 8 hi + 8 lo columns of the
//              same 16 features) -> D3 = up_2 acc -> o2 (in place)
//   [192,208)  Wc r1 acc (folded res_1 / up_0, see unet_tc.cuh) ... later W_u0 y1 acc
//   [192,256)  res_2 acc b (output features 64..127)
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace socm {
namespace hx {

constexpr int TP = 128;
constexpr int H0 = 256, H1 = 128, H2 = 64;
constexpr int KIN = 16;               // d + 1 padded to one K = 16 step: the engine covers d <= 15
constexpr int NY = 16;                // N of the folded up_0 MMAs
constexpr int MAX_D = 15;
constexpr int SLOT_BYTES = 20480;     // tape slot stride / ring stage
constexpr int NSTAGE = 3;
constexpr int CHUNK_F = 32;
constexpr int CHUNK_HALF = TP * CHUNK_F * 2;   // bytes of the hi (or lo) part of a shared-memory A chunk
constexpr int CHUNK_BYTES = 2 * CHUNK_HALF;    // 16 KB
constexpr int XIN_HALF = TP * KIN * 2;         // 4 KB
constexpr int WC_LAYER = 9;
constexpr int PIECE_BYTES = 2048;     // W0 piece (N = 32, K = 16) or up_0 chunk block (N = 16, K = 32): hi + lo
constexpr int D1_MAIN = 2 * (H1 + NY) * 32 * 2;  // 18 432: down_1 (+ Wc) block of one r1 chunk, hi + lo slabs
constexpr int U1_MAIN = 2 * 32 * H1 * 2;         // 16 384: up_1 piece block, hi + lo slabs

// forward tape: slot 0: W0 pieces 0,1 | 1..8: down_1 chunk c (+ Wc) + W0 piece c+2 | 9,10: down_2 | 11,12: res_2 b |
// 13,14: res_2 a | 15,16: up_2 | 17..24: up_1 piece p + up_0 block of chunk p-2 | 25: up_0 blocks of chunks 6, 7
constexpr int FWD_SLOTS = 26;
enum FwdSlot { FS_W0 = 0, FS_D1 = 1, FS_D2 = 9, FS_R2B = 11, FS_R2A = 13, FS_U2 = 15, FS_U1 = 17, FS_U0 = 25 };

// canonical no-swizzle K-major layouts (bytes, halfs)
__host__ __device__ inline int wslab_off(int n, int k, int Kc) {
  return (n % 8) * 16 + (k % 8) * 2 + (k / 8) * 128 + (n / 8) * (Kc * 16);
}
constexpr uint32_t W_LBO = 128, W_KSTEP = 256;                        // weight slab: SBO = Kc * 16
constexpr uint32_t ACT_LBO = 2048, ACT_SBO = 128, ACT_KSTEP = 4096;   // activation chunk / xin: [128 points][features]

struct SlotDesc {
  int layer;       // index into socm_unet::w, or WC_LAYER
  int n0, N;       // rows of the block
  int k0, Kc;      // contraction range
  int ktot;        // row length of W[out][ktot]
  int transposed;  // 1: B[n][k] = W[k0 + k][n0 + n]  (dgrad tapes)
  int klim, nlim;  // zero padding beyond
  int slab_n, slab_row;  // block is rows [slab_row, slab_row + N) of a slab of slab_n rows (0: N)
  int sid;               // weight-scale index (WScale); -1: the forward scale of `layer`
};
struct PackItem {
  int slot, byte_off;
  SlotDesc sd;
};
constexpr int NOLIM = 1 << 30;
constexpr int FWD_ITEMS = 48;
__host__ __device__ inline PackItem fwd_item(int d, int i) {
  if (i < 2) return PackItem{FS_W0, i * PIECE_BYTES, SlotDesc{0, 32 * i, 32, 0, KIN, d + 1, 0, d + 1, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 8) return PackItem{FS_D1 + i, 0, SlotDesc{1, 0, H1, 32 * i, 32, H0, 0, NOLIM, NOLIM, H1 + NY, 0, -1}};
  i -= 8;
  if (i < 8) return PackItem{FS_D1 + i, 0, SlotDesc{WC_LAYER, 0, NY, 32 * i, 32, H0, 0, NOLIM, d, H1 + NY, H1, -1}};
  i -= 8;
  if (i < 6) return PackItem{FS_D1 + i, D1_MAIN, SlotDesc{0, 32 * (i + 2), 32, 0, KIN, d + 1, 0, d + 1, NOLIM, 0, 0, -1}};
  i -= 6;
  if (i < 2) return PackItem{FS_D2 + i, 0, SlotDesc{2, 0, H2, 64 * i, 64, H1, 0, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 2) return PackItem{FS_R2B + i, 0, SlotDesc{5, 64, 64, 64 * i, 64, H1, 0, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 2) return PackItem{FS_R2A + i, 0, SlotDesc{5, 0, 64, 64 * i, 64, H1, 0, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 2) return PackItem{FS_U2 + i, 0, SlotDesc{6, 0, H1, 32 * i, 32, H2, 0, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 8) return PackItem{FS_U1 + i, 0, SlotDesc{7, 32 * i, 32, 0, H1, H1, 0, NOLIM, NOLIM, 0, 0, -1}};
  i -= 8;
  if (i < 6) return PackItem{FS_U1 + 2 + i, U1_MAIN, SlotDesc{8, 0, NY, 32 * i, 32, H0, 0, NOLIM, d, 0, 0, -1}};
  i -= 6;
  return PackItem{FS_U0, i * PIECE_BYTES, SlotDesc{8, 0, NY, 32 * (6 + i), 32, H0, 0, NOLIM, d, 0, 0, -1}};
}
__host__ __device__ inline uint32_t fwd_slot_bytes(int s) {
  if (s == FS_W0) return 2 * PIECE_BYTES;
  if (s < FS_D2) return (uint32_t)(D1_MAIN + (s - FS_D1 < 6 ? PIECE_BYTES : 0));
  if (s < FS_U1) return 16384u;
  if (s < FS_U1 + 2) return (uint32_t)U1_MAIN;
  if (s < FS_U0) return (uint32_t)(U1_MAIN + PIECE_BYTES);
  return (uint32_t)(2 * PIECE_BYTES);
}

// backward (dgrad) tape, W^T blocks: slot 0: up_0^T pieces 0,1 | 1..8: up_1^T chunk c + up_0^T piece c+2 | 9,10: up_2^T
// (two K = 32 blocks per slot) | 11,12: res_2^T b | 13,14: res_2^T a | 15: down_2^T a | 16: down_2^T b |
// 17..24: down_1^T piece q + Wc^T piece q
constexpr int BWD_SLOTS = 25;
enum BwdSlot { BS_U0T = 0, BS_U1T = 1, BS_U2T = 9, BS_R2TB = 11, BS_R2TA = 13, BS_D2TA = 15, BS_D2TB = 16, BS_D1T = 17 };
// weight scales: 0..9 forward (layer index, WC_LAYER = 9); backward blocks reuse the forward scale of their matrix except
// the two products that ACCUMULATE ONTO another product's accumulator and must arrive in its units:
//   down_2^T onto res_2^T (d_r2):  s_dz3 w_d2T = s_do2 w_r2     |     Wc^T onto down_1^T (d_r1):  s_dy0 w_wcT = s_dz2 w_d1
enum WScale { WS_D2T = 10, WS_WCT = 11, N_WSCALE = 12 };
constexpr int BWD_ITEMS = 44;
__host__ __device__ inline PackItem bwd_item(int d, int i) {
  if (i < 2) return PackItem{BS_U0T, i * PIECE_BYTES, SlotDesc{8, 32 * i, 32, 0, KIN, H0, 1, d, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 8) return PackItem{BS_U1T + i, 0, SlotDesc{7, 0, H1, 32 * i, 32, H1, 1, NOLIM, NOLIM, 0, 0, -1}};
  i -= 8;
  if (i < 6) return PackItem{BS_U1T + i, 16384, SlotDesc{8, 32 * (i + 2), 32, 0, KIN, H0, 1, d, NOLIM, 0, 0, -1}};
  i -= 6;
  if (i < 4) return PackItem{BS_U2T + i / 2, (i % 2) * 8192, SlotDesc{6, 0, H2, 32 * i, 32, H2, 1, NOLIM, NOLIM, 0, 0, -1}};
  i -= 4;
  if (i < 2) return PackItem{BS_R2TB + i, 0, SlotDesc{5, 64, 64, 64 * i, 64, H1, 1, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 2) return PackItem{BS_R2TA + i, 0, SlotDesc{5, 0, 64, 64 * i, 64, H1, 1, NOLIM, NOLIM, 0, 0, -1}};
  i -= 2;
  if (i < 2) return PackItem{BS_D2TA, i * 8192, SlotDesc{2, 0, 64, 32 * i, 32, H1, 1, NOLIM, NOLIM, 0, 0, WS_D2T}};
  i -= 2;
  if (i < 2) return PackItem{BS_D2TB, i * 8192, SlotDesc{2, 64, 64, 32 * i, 32, H1, 1, NOLIM, NOLIM, 0, 0, WS_D2T}};
  i -= 2;
  if (i < 8) return PackItem{BS_D1T + i, 0, SlotDesc{1, 32 * i, 32, 0, H1, H0, 1, NOLIM, NOLIM, 0, 0, -1}};
  i -= 8;
  return PackItem{BS_D1T + i, U1_MAIN, SlotDesc{WC_LAYER, 32 * i, 32, 0, KIN, H0, 1, d, NOLIM, 0, 0, WS_WCT}};
}
__host__ __device__ inline uint32_t bwd_slot_bytes(int s) {
  if (s == BS_U0T) return 2 * PIECE_BYTES;
  if (s < BS_U2T) return (uint32_t)(16384 + (s - BS_U1T < 6 ? PIECE_BYTES : 0));
  if (s < BS_D1T) return 16384u;
  return (uint32_t)(U1_MAIN + PIECE_BYTES);
}

// ---------------------------------------------------------------- small block (floats, read from shared memory)
// layer inputs whose scale is calibrated: 0 xin, 1 r1, 2 r2, 3 r3, 4 o2, 5 y1
enum Act { A_X = 0, A_R1, A_R2, A_R3, A_O2, A_Y1, N_ACT };
struct Small {
  int b_d0, b_d1, b_d2, b_u2, b_r2, b_u1, bc, r0, b_r0;
  int sa;   // [N_ACT]  activation scales (what the producer multiplies by before the fp16 split)
  int inv;  // [10]     1 / (input scale * weight scale) per layer index (WC_LAYER = 9): the epilogue's un-scale factor
  // The biases b_d0, b_d1, b_d2, b_u2, b_r2, b_u1 are stored PRE-MULTIPLIED by the scale of the activation they produce
  // (r1, r2, r3, o2, o2, y1), and `invs` = inv * that scale: relu(acc inv + b) s = relu(acc (inv s) + b s) for s > 0,
  // which saves one multiply per element in the epilogue.
  int invs;  // [10]
  // backward (dgrad chain on row-normalised gradients, see loss_h.cu): operand scales and the accumulator -> operand /
  // accumulator -> true-value factors of the five products
  int sb;    // [N_BACT]  scales of the backward operands d_y0, d_y1, d_o2 (= d_y2), d_z3, d_z2
  int bf;    // [N_BPROD] acc -> next operand:   s_out / (s_in w)
  int bt;    // [N_BPROD] acc -> normalised true value: 1 / (s_in w)
  int sk;    // [N_SCRATCH_T] scale of every tensor of the fp16 scratch (ScratchT), read by K3b's flush (wgrad_h.cu)
  int total;
};
enum BAct { B_DY0 = 0, B_DY1, B_DO2, B_DZ3, B_DZ2, B_DZ1, N_BACT };
// tensors of the K3a -> K3b scratch, in the order of their feature blocks (loss_tc.cuh)
enum ScratchT { T_XIN = 0, T_R1, T_R2, T_R3, T_O2, T_Y1, T_DY0, T_DO0, T_DY1, T_DY2, T_DO2, T_DZ3, T_DZ2, T_DZ1, N_SCRATCH_T };
enum BProd { P_U0T = 0, P_U1T, P_U2T, P_R2T, P_D1T, N_BPROD };   // d_o1, d_o2, d_r3, d_r2 (res_2^T + down_2^T), d_r1
__host__ __device__ inline Small small_layout() {
  Small o;
  int p = 0;
  o.b_d0 = p; p += H0;
  o.b_d1 = p; p += H1;
  o.b_d2 = p; p += H2;
  o.b_u2 = p; p += H1;
  o.b_r2 = p; p += H1;
  o.b_u1 = p; p += H0;
  o.bc = p; p += KIN;
  o.r0 = p; p += KIN * KIN;
  o.b_r0 = p; p += KIN;
  o.sa = p; p += 8;
  o.inv = p; p += 12;
  o.invs = p; p += 12;
  o.sb = p; p += 8;
  o.bf = p; p += 8;
  o.bt = p; p += 8;
  o.sk = p; p += 16;
  o.total = ((p + 3) / 4) * 4;
  return o;
}
// which calibrated activation feeds layer l (index into socm_unet::w; WC_LAYER reads r1 like down_1)
__host__ __device__ inline int act_of_layer(int l) {
  switch (l) {
    case 0: return A_X;
    case 1: return A_R1;
    case 2: return A_R2;
    case 5: return A_R2;
    case 6: return A_R3;
    case 7: return A_O2;
    case 8: return A_Y1;
    case WC_LAYER: return A_R1;
    default: return A_X;
  }
}
// which calibrated activation layer l produces (res_2 and up_2 both feed o2); -1: none (up_0, Wc: fp32 outputs)
__host__ __device__ inline int act_out_of_layer(int l) {
  switch (l) {
    case 0: return A_R1;
    case 1: return A_R2;
    case 2: return A_R3;
    case 5: return A_O2;
    case 6: return A_O2;
    case 7: return A_Y1;
    default: return -1;
  }
}
// workspace: [tape FWD_SLOTS x SLOT_BYTES][small][Wc 16 x 256 floats][max buffer: N_ACT + 10 uint32]
// workspace: [forward tape][backward tape][small][Wc 16 x 256 floats][max buffer 64 uint32]  (the backward tape is always
// reserved; the rollout does not fill it)
constexpr int WC_FLOATS = NY * H0;
__host__ __device__ inline int64_t tape_bytes() { return (int64_t)(FWD_SLOTS + BWD_SLOTS) * SLOT_BYTES; }
__host__ __device__ inline int64_t workspace_bytes() {
  return tape_bytes() + (int64_t)small_layout().total * 4 + (int64_t)WC_FLOATS * 4 + 64 * 4;
}
__host__ __device__ inline float* small_ptr(unsigned char* ws) { return reinterpret_cast<float*>(ws + tape_bytes()); }
__host__ __device__ inline float* wc_ptr(unsigned char* ws) { return small_ptr(ws) + small_layout().total; }
__host__ __device__ inline uint32_t* max_ptr(unsigned char* ws) { return reinterpret_cast<uint32_t*>(wc_ptr(ws) + WC_FLOATS); }
// max buffer slots: [0, N_ACT) forward activations, [N_ACT, N_ACT + 10) weights, [MX_B, MX_B + N_BACT) backward gains
// (max |d_layer| per unit of max |d loss / d nabla_V| of the same point), MX_W = max |w_m| over all paths (exact),
// MX_DIFF = max |nabla_V - target| over the sample points
constexpr int MX_B = 24, MX_W = 40, MX_DIFF = 41;

// power-of-two scale that maps `mx` into [target / 2, target)
__host__ __device__ inline float pow2_scale(float mx, float target) {
  if (!(mx > 0.f) || !(mx < 3.0e38f)) return 1.f;
  int e = (int)floorf(log2f(target / mx));
  e = e < -60 ? -60 : (e > 60 ? 60 : e);
  return exp2f((float)e);
}
constexpr float ACT_TARGET = 64.f, W_TARGET = 1024.f;

// ---------------------------------------------------------------- MMA issue helpers (elected thread)
// one weight block [N][16 * KS] against an A operand in TMEM: k-step ks reads hi columns a_col + a_stride * ks + [0,8),
// lo columns + 8 (the in-place layout of 16-feature groups)
template <int N, int KS>
__device__ __forceinline__ void issue_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_smem, uint32_t kc, bool fresh) {
  constexpr uint32_t id = umma::idesc_f16(TP, N);
  const uint32_t slab = (uint32_t)N * kc * 2, sbo = kc * 16;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint64_t bh = umma::smem_desc(b_smem + ks * W_KSTEP, W_LBO, sbo);
    const uint64_t bl = umma::smem_desc(b_smem + slab + ks * W_KSTEP, W_LBO, sbo);
    umma::mma_ts_f16(d_tmem, a_tmem + 16 * ks, bh, id, (fresh && ks == 0) ? 0u : 1u);
    umma::mma_ts_f16(d_tmem, a_tmem + 16 * ks + 8, bh, id, 1u);
    umma::mma_ts_f16(d_tmem, a_tmem + 16 * ks, bl, id, 1u);
  }
}
// the same with the A operand in two column ranges: k-steps [0, KS/2) at a0, [KS/2, KS) at a1
template <int N, int KS>
__device__ __forceinline__ void issue_ts2(uint32_t d_tmem, uint32_t a0, uint32_t a1, uint32_t b_smem, uint32_t kc, bool fresh) {
  constexpr uint32_t id = umma::idesc_f16(TP, N);
  const uint32_t slab = (uint32_t)N * kc * 2, sbo = kc * 16;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint64_t bh = umma::smem_desc(b_smem + ks * W_KSTEP, W_LBO, sbo);
    const uint64_t bl = umma::smem_desc(b_smem + slab + ks * W_KSTEP, W_LBO, sbo);
    const uint32_t a = ks < KS / 2 ? a0 + 16 * ks : a1 + 16 * (ks - KS / 2);
    umma::mma_ts_f16(d_tmem, a, bh, id, (fresh && ks == 0) ? 0u : 1u);
    umma::mma_ts_f16(d_tmem, a + 8, bh, id, 1u);
    umma::mma_ts_f16(d_tmem, a, bl, id, 1u);
  }
}
// A operand in shared memory (chunk layout; lo part at a_smem + a_lo_off); the block has `slab_rows` rows per slab
template <int N, int KS>
__device__ __forceinline__ void issue_ss(uint32_t d_tmem, uint32_t a_smem, uint32_t a_lo_off, uint32_t b_smem,
                                         uint32_t kc, uint32_t slab_rows, bool fresh) {
  constexpr uint32_t id = umma::idesc_f16(TP, N);
  const uint32_t slab = slab_rows * kc * 2, sbo = kc * 16;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const uint64_t bh = umma::smem_desc(b_smem + ks * W_KSTEP, W_LBO, sbo);
    const uint64_t bl = umma::smem_desc(b_smem + slab + ks * W_KSTEP, W_LBO, sbo);
    const uint64_t ah = umma::smem_desc(a_smem + ks * ACT_KSTEP, ACT_LBO, ACT_SBO);
    const uint64_t al = umma::smem_desc(a_smem + a_lo_off + ks * ACT_KSTEP, ACT_LBO, ACT_SBO);
    umma::mma_ss_f16(d_tmem, ah, bh, id, (fresh && ks == 0) ? 0u : 1u);
    umma::mma_ss_f16(d_tmem, al, bh, id, 1u);
    umma::mma_ss_f16(d_tmem, ah, bl, id, 1u);
  }
}

// ---------------------------------------------------------------- epilogue helpers (thread <-> TMEM lane)
// 16 scaled values -> 8 hi + 8 lo packed registers (pairs of consecutive features; the even feature in the low half)
__device__ __forceinline__ void split16(const float* x, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < 8; ++i) umma::split_h2(x[2 * i], x[2 * i + 1], hi[i], lo[i]);
}
// 16 features (two core-matrix columns q0, q0 + 1) of point p into a shared-memory A chunk
__device__ __forceinline__ void store_chunk16(unsigned char* chunk, int p, int q0, const uint32_t* hi, const uint32_t* lo) {
  unsigned char* base = chunk + (p % 8) * 16 + (p / 8) * 128 + q0 * 2048;
  *reinterpret_cast<uint4*>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + 2048) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
  *reinterpret_cast<uint4*>(base + CHUNK_HALF) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  *reinterpret_cast<uint4*>(base + CHUNK_HALF + 2048) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
}
// in-place A operand: the 16 TMEM columns of a 16-feature group become [hi x 8 | lo x 8]
__device__ __forceinline__ void store_group16(uint32_t taddr, const uint32_t* hi, const uint32_t* lo) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r[i] = hi[i];
    r[8 + i] = lo[i];
  }
  umma::tmem_st16(taddr, r);
}
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}
__device__ __forceinline__ void e_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 epilogue warps

// ---------------------------------------------------------------- per-call setup (rollout_h.cu)
// Calibration sample points: rollout mode (states == nullptr): the first paths' x0 spread by Philox noise at four times;
// loss mode: points of the stored trajectories (+ the loss gradient at them for the backward gains).
struct CalibArgs {
  const float* x0;        // rollout mode [B][d]
  const float* step_tab;  // rollout mode [5][K] (row 4: t_k, row 0: dt_k)
  const float* states;    // loss mode [K+1][B][d]
  const float* ts;        // loss mode [K+1]
  const float* target;    // loss mode [B][ldt]
  const float* w;         // loss mode [B] path weights (importance weights or dF/dS_m of the path functionals)
  int B, K, ldt, n_samples;
  float lmbd, loss_scale;
};
int launch_wgrad_h(const unsigned char* scratch, int n_tiles, int d, const float* scales, float* grad, float* aux,
                   cudaStream_t stream);
int setup_h(const socm_unet* net, unsigned char* ws, const CalibArgs& c, bool with_bwd, cudaStream_t stream);

}  // namespace hx
}  // namespace socm