// unet_tile.cuh -- FFMA tile engine for the default FullyConnectedUNet (hdims [256,128,64]).
//
// One CTA (256 threads) owns a tile of BT = 64 points.  Activations live in shared memory,
// feature-major ([feature][point], pitch LD), so that one layer is a register-tiled
// C[64 x N] += X^T[64 x k] * W^T[k x N] with both operands read as float4 along the
// non-contracted index.  The four big layers (163 840 of the 170 562 weights) do not fit
// shared memory, so they are streamed from L2 every evaluation as a fixed "tape" of 16 KB
// chunks in consumption order, by cp.async.bulk (TMA engine, UBLKCP) into a 3-stage ring
// signalled through mbarriers; the tape is a one-off repack of the nn.Linear weights
// (k-major = transposed) done once per iteration by pack_tape_kernel.
#pragma once
#include "common.cuh"

namespace socm {
namespace tile {

constexpr int BT = 64;    // points per tile
constexpr int LD = 68;    // smem pitch of a feature row (floats): 64 + 4 keeps float4 alignment
constexpr int NT = 256;   // threads per CTA
constexpr int H0 = 256, H1 = 128, H2 = 64;
constexpr int CHUNK = 4096;  // floats per tape chunk (16 KB)
constexpr int STAGES = 3;

// ---- forward tape: W^T ([k][n], n contiguous) of the big layers in consumption order
constexpr int FT_D1 = 0;                    // down_1^T [256][128]
constexpr int FT_D2 = FT_D1 + H0 * H1;      // down_2^T [128][64]
constexpr int FT_U2 = FT_D2 + H1 * H2;      // up_2^T   [64][128]
constexpr int FT_R2 = FT_U2 + H2 * H1;      // res_2^T  [128][128]
constexpr int FT_U1 = FT_R2 + H1 * H1;      // up_1^T   [128][256]
constexpr int FT_R1 = FT_U1 + H1 * H0;      // res_1^T  [256][256]
constexpr int FT_FLOATS = FT_R1 + H0 * H0;  // 163 840
constexpr int FT_CHUNKS = FT_FLOATS / CHUNK;  // 40
static_assert(FT_FLOATS % CHUNK == 0, "tape must be whole chunks");

// ---- backward tape: W ([n][k], native nn.Linear layout = contraction-major for dgrad) in the
// order the backward pass consumes them (see loss_tile.cu)
constexpr int BT_U1 = 0;                    // up_1   [256][128]  d_o2  = W^T d_y1
constexpr int BT_U2 = BT_U1 + H0 * H1;      // up_2   [128][64]   d_r3  = W^T d_y2
constexpr int BT_R2 = BT_U2 + H1 * H2;      // res_2  [128][128]  d_r2  = W^T d_o2
constexpr int BT_D2 = BT_R2 + H1 * H1;      // down_2 [64][128]   d_r2 += W^T d_z3
constexpr int BT_D1 = BT_D2 + H2 * H1;      // down_1 [128][256]  d_r1  = W^T d_z2
constexpr int BT_R1 = BT_D1 + H1 * H0;      // res_1  [256][256]  d_r1 += W^T d_o1
constexpr int BT_FLOATS = BT_R1 + H0 * H0;  // 163 840
constexpr int BT_CHUNKS = BT_FLOATS / CHUNK;
static_assert(BT_FLOATS == FT_FLOATS, "same weights");

// ---- small block (read through L1 with __ldg): everything that depends on d, and biases
struct SmallOff {
  int d0t;   // down_0^T [(d+1)][256]
  int b_d0, b_d1, b_d2, b_u2, b_r2, b_u1, b_r1;
  int u0;    // up_0 [d][256] native
  int b_u0;
  int r0;    // res_0 [d][d+1] native
  int b_r0;
  int total;
};
__host__ __device__ inline SmallOff small_offsets(int d) {
  SmallOff o;
  int p = 0;
  o.d0t = p; p += (d + 1) * H0;
  o.b_d0 = p; p += H0;
  o.b_d1 = p; p += H1;
  o.b_d2 = p; p += H2;
  o.b_u2 = p; p += H1;
  o.b_r2 = p; p += H1;
  o.b_u1 = p; p += H0;
  o.b_r1 = p; p += H0;
  o.u0 = p; p += d * H0;
  o.b_u0 = p; p += ((d + 3) / 4) * 4;
  o.r0 = p; p += ((d * (d + 1) + 3) / 4) * 4;
  o.b_r0 = p; p += ((d + 3) / 4) * 4;
  o.total = p;
  return o;
}
// packed workspace: [fwd tape][bwd tape][small block]
__host__ __device__ inline int64_t packed_floats(int d) { return (int64_t)FT_FLOATS + BT_FLOATS + small_offsets(d).total; }

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- weight-tape pipeline
// All NT threads consume every chunk in lock step; thread 0 is the producer.  A stage is
// re-armed right after the __syncthreads() that ends its consumption.
struct Pipe {
  float* stage;      // STAGES * CHUNK floats, 128-byte aligned
  uint64_t* full;    // STAGES mbarriers
  const float* tape; // global
  int tape_chunks;   // chunks per pass over the tape
  uint32_t cons;     // chunks consumed so far (monotonic, identical in all threads)
  uint32_t total;    // chunks this CTA will consume in its lifetime

  __device__ __forceinline__ void issue(uint32_t idx) {  // thread 0 only
    const int s = idx % STAGES;
    mbar_expect_tx(&full[s], CHUNK * 4);
    bulk_g2s(stage + s * CHUNK, tape + (size_t)(idx % tape_chunks) * CHUNK, CHUNK * 4, &full[s]);
  }
  __device__ __forceinline__ void start(float* stage_, uint64_t* full_, const float* tape_, int tape_chunks_,
                                        uint32_t total_) {
    stage = stage_;
    full = full_;
    tape = tape_;
    tape_chunks = tape_chunks_;
    cons = 0;
    total = total_;
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (uint32_t i = 0; i < STAGES && i < total; ++i) issue(i);
  }
  __device__ __forceinline__ const float* acquire() {
    mbar_wait(&full[cons % STAGES], (cons / STAGES) & 1u);
    return stage + (cons % STAGES) * CHUNK;
  }
  __device__ __forceinline__ void release() {
    __syncthreads();
    if (threadIdx.x == 0 && cons + STAGES < total) issue(cons + STAGES);
    ++cons;
  }
};

// ---------------------------------------------------------------- register-tile micro kernels
// Thread coordinates inside the 64 x N output tile: 8 warps as 2 (points) x 4 (features),
// lanes as 4 (points) x 8 (features); every thread owns 8 points x TN features.
struct Coord {
  int pA;  // first point group  [pA, pA+4), second group at pA+16
  int wn;  // warp column 0..3
  int ln;  // lane column 0..7
  __device__ __forceinline__ Coord() {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pA = (w >> 2) * 32 + (lane >> 3) * 4;
    wn = w & 3;
    ln = lane & 7;
  }
  template <int N>
  __device__ __forceinline__ int n0() const {  // first feature of this thread (TN=8: second group at +32)
    if (N == 256) return wn * 64 + ln * 4;
    if (N == 128) return wn * 32 + ln * 4;
    return wn * 16 + ln * 2;  // N == 64
  }
  // feature index of accumulator column j
  template <int N>
  __device__ __forceinline__ int feat(int j) const {
    if (N == 256) return n0<256>() + (j & 3) + (j >> 2) * 32;
    return n0<N>() + j;
  }
  __device__ __forceinline__ int point(int i) const { return pA + (i & 3) + (i >> 2) * 16; }
};

template <int N>
struct TileAcc {
  static constexpr int TN = N / 32;
  float v[8][TN];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) v[i][j] = 0.f;
  }
};

// acc[p][n] += sum_{k<KC} X[k][p] * Wc[k][n];  X: smem rows (pitch LD), Wc: [KC][N] (smem or global)
template <int N, int KC, bool kGlobalW = false>
__device__ __forceinline__ void mac_chunk(TileAcc<N>& acc, const float* __restrict__ X,
                                          const float* __restrict__ Wc, const Coord& co, int kc_runtime = KC) {
  constexpr int TN = N / 32;
  const float* xp = X + co.pA;
  const float* wp = Wc + co.n0<N>();
  const int kc = kGlobalW ? kc_runtime : KC;
#pragma unroll 4
  for (int k = 0; k < kc; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(xp + k * LD);
    const float4 a1 = *reinterpret_cast<const float4*>(xp + k * LD + 16);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float b[TN];
    if constexpr (TN == 8) {
      const float4 b0 = kGlobalW ? __ldg(reinterpret_cast<const float4*>(wp + k * N))
                                 : *reinterpret_cast<const float4*>(wp + k * N);
      const float4 b1 = kGlobalW ? __ldg(reinterpret_cast<const float4*>(wp + k * N + 32))
                                 : *reinterpret_cast<const float4*>(wp + k * N + 32);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    } else if constexpr (TN == 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(wp + k * N);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
    } else {
      const float2 b0 = *reinterpret_cast<const float2*>(wp + k * N);
      b[0] = b0.x; b[1] = b0.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc.v[i][j] = fmaf(a[i], b[j], acc.v[i][j]);
  }
}

// One streamed layer part: acc += X[KDIM rows] * tape segment (KDIM x N), chunk by chunk.
template <int N, int KDIM>
__device__ __forceinline__ void stream_layer(TileAcc<N>& acc, const float* X, Pipe& pipe, const Coord& co) {
  constexpr int KC = CHUNK / N;
  static_assert(KDIM % KC == 0, "layer must be whole chunks");
  for (int k0 = 0; k0 < KDIM; k0 += KC) {
    const float* wc = pipe.acquire();
    mac_chunk<N, KC>(acc, X + k0 * LD, wc, co);
    pipe.release();
  }
}

// acc[.][j] += bias[feat(j)]
template <int N>
__device__ __forceinline__ void add_bias(TileAcc<N>& acc, const float* __restrict__ bias, const Coord& co) {
#pragma unroll
  for (int j = 0; j < TileAcc<N>::TN; ++j) {
    const float bj = __ldg(bias + co.feat<N>(j));
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i][j] += bj;
  }
}
template <int N>
__device__ __forceinline__ void relu_acc(TileAcc<N>& acc) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TileAcc<N>::TN; ++j) acc.v[i][j] = fmaxf(acc.v[i][j], 0.f);
}
// bit (i*TN + j) of the returned mask = acc > 0  (TN <= 8 -> up to 64 bits)
template <int N>
__device__ __forceinline__ unsigned long long positive_mask(const TileAcc<N>& acc) {
  unsigned long long m = 0ull;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TileAcc<N>::TN; ++j)
      if (acc.v[i][j] > 0.f) m |= 1ull << (i * TileAcc<N>::TN + j);
  return m;
}
template <int N>
__device__ __forceinline__ void apply_mask(TileAcc<N>& acc, unsigned long long m) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TileAcc<N>::TN; ++j)
      if (!((m >> (i * TileAcc<N>::TN + j)) & 1ull)) acc.v[i][j] = 0.f;
}

// Y[feat][point] = acc   (feature-major store, two float4 per feature)
template <int N>
__device__ __forceinline__ void store_tile(const TileAcc<N>& acc, float* Y, const Coord& co) {
#pragma unroll
  for (int j = 0; j < TileAcc<N>::TN; ++j) {
    float* row = Y + co.feat<N>(j) * LD + co.pA;
    *reinterpret_cast<float4*>(row) = make_float4(acc.v[0][j], acc.v[1][j], acc.v[2][j], acc.v[3][j]);
    *reinterpret_cast<float4*>(row + 16) = make_float4(acc.v[4][j], acc.v[5][j], acc.v[6][j], acc.v[7][j]);
  }
}
template <int N>
__device__ __forceinline__ void load_tile(TileAcc<N>& acc, const float* Y, const Coord& co) {
#pragma unroll
  for (int j = 0; j < TileAcc<N>::TN; ++j) {
    const float* row = Y + co.feat<N>(j) * LD + co.pA;
    const float4 a = *reinterpret_cast<const float4*>(row);
    const float4 b = *reinterpret_cast<const float4*>(row + 16);
    acc.v[0][j] = a.x; acc.v[1][j] = a.y; acc.v[2][j] = a.z; acc.v[3][j] = a.w;
    acc.v[4][j] = b.x; acc.v[5][j] = b.y; acc.v[6][j] = b.z; acc.v[7][j] = b.w;
  }
}

// Spill a register tile to this CTA's global scratch, compact feature-major [feat][BT]
// (activations needed again by the backward pass do not fit shared memory).
template <int N>
__device__ __forceinline__ void spill_acc(const TileAcc<N>& acc, float* __restrict__ S, const Coord& co) {
#pragma unroll
  for (int j = 0; j < TileAcc<N>::TN; ++j) {
    float* row = S + co.feat<N>(j) * BT + co.pA;
    __stcg(reinterpret_cast<float4*>(row), make_float4(acc.v[0][j], acc.v[1][j], acc.v[2][j], acc.v[3][j]));
    __stcg(reinterpret_cast<float4*>(row + 16), make_float4(acc.v[4][j], acc.v[5][j], acc.v[6][j], acc.v[7][j]));
  }
}
// Reload `rows` feature rows from the compact scratch into a shared-memory tile (pitch LD).
__device__ __forceinline__ void reload_rows(float* __restrict__ Y, const float* __restrict__ S, int rows) {
  for (int idx = threadIdx.x; idx < rows * (BT / 4); idx += NT) {
    const int r = idx / (BT / 4), c4 = idx % (BT / 4);
    *reinterpret_cast<float4*>(Y + r * LD + c4 * 4) = __ldcg(reinterpret_cast<const float4*>(S + r * BT + c4 * 4));
  }
}

// ---------------------------------------------------------------- the forward pass of one tile
// smem tiles: XIN [(d+1)][LD] rows 0 = t, 1.. = x;  R1 [256][LD], R2 [128][LD], R3 [64][LD],
// V [d][LD] receives nabla_V.  On return R1 holds o1 and R2 holds o2 (forward-only aliasing);
// if kKeep is set the caller passes distinct O2/O1 tiles and receives the ReLU masks of the
// up layers (for the backward pass).
struct FwdMasks {
  unsigned long long y1;  // 64 x 256 tile, TN = 8 -> 64 bits per thread
  unsigned int y2;        // 64 x 128 tile, TN = 4 -> 32 bits per thread
};

template <bool kKeep>
__device__ __forceinline__ void forward_tile(int d, const float* __restrict__ small, const SmallOff& so,
                                             const float* XIN, float* R1, float* R2, float* R3, float* O2,
                                             float* O1, float* V, Pipe& pipe, const Coord& co, FwdMasks* masks,
                                             float* spill_r1 = nullptr, float* spill_r2 = nullptr) {
  // down_0: (d+1) -> 256, weights from L1/L2 (tiny K)
  {
    TileAcc<256> acc;
    acc.zero();
    add_bias<256>(acc, small + so.b_d0, co);
    mac_chunk<256, 1, true>(acc, XIN, small + so.d0t, co, d + 1);
    relu_acc<256>(acc);
    store_tile<256>(acc, R1, co);
    if (kKeep) spill_acc<256>(acc, spill_r1, co);
  }
  __syncthreads();
  // down_1: 256 -> 128
  {
    TileAcc<128> acc;
    acc.zero();
    stream_layer<128, H0>(acc, R1, pipe, co);
    add_bias<128>(acc, small + so.b_d1, co);
    relu_acc<128>(acc);
    store_tile<128>(acc, R2, co);
    if (kKeep) spill_acc<128>(acc, spill_r2, co);
  }
  __syncthreads();
  // down_2: 128 -> 64
  {
    TileAcc<64> acc;
    acc.zero();
    stream_layer<64, H1>(acc, R2, pipe, co);
    add_bias<64>(acc, small + so.b_d2, co);
    relu_acc<64>(acc);
    store_tile<64>(acc, R3, co);
  }
  __syncthreads();
  // o2 = relu(up_2 r3 + b) + res_2 r2 + b
  {
    TileAcc<128> acc;
    acc.zero();
    stream_layer<128, H2>(acc, R3, pipe, co);
    add_bias<128>(acc, small + so.b_u2, co);
    if (kKeep) masks->y2 = (unsigned int)positive_mask<128>(acc);
    relu_acc<128>(acc);
    stream_layer<128, H1>(acc, R2, pipe, co);
    add_bias<128>(acc, small + so.b_r2, co);
    // the last release() of stream_layer synchronised all reads of R2
    store_tile<128>(acc, O2, co);
  }
  __syncthreads();
  // o1 = relu(up_1 o2 + b) + res_1 r1 + b
  {
    TileAcc<256> acc;
    acc.zero();
    stream_layer<256, H1>(acc, O2, pipe, co);
    add_bias<256>(acc, small + so.b_u1, co);
    if (kKeep) masks->y1 = positive_mask<256>(acc);
    relu_acc<256>(acc);
    stream_layer<256, H0>(acc, R1, pipe, co);
    add_bias<256>(acc, small + so.b_r1, co);
    store_tile<256>(acc, O1, co);
  }
  __syncthreads();
  // o0 = relu(up_0 o1 + b) + res_0 xin + b : thread -> (point p, features j = jg, jg+4, ...)
  {
    const int p = threadIdx.x & (BT - 1), jg = threadIdx.x >> 6;
    float au[kMaxDim / 4], ar[kMaxDim / 4];
#pragma unroll
    for (int i = 0; i < kMaxDim / 4; ++i) {
      const int j = jg + 4 * i;
      au[i] = j < d ? __ldg(small + so.b_u0 + j) : 0.f;
      ar[i] = j < d ? __ldg(small + so.b_r0 + j) : 0.f;
    }
    const float* wu = small + so.u0;
    for (int k = 0; k < H0; ++k) {
      const float o = O1[k * LD + p];
#pragma unroll
      for (int i = 0; i < kMaxDim / 4; ++i) {
        const int j = jg + 4 * i;
        if (j < d) au[i] = fmaf(__ldg(wu + j * H0 + k), o, au[i]);
      }
    }
    const float* wr = small + so.r0;
    for (int k = 0; k <= d; ++k) {
      const float xv = XIN[k * LD + p];
#pragma unroll
      for (int i = 0; i < kMaxDim / 4; ++i) {
        const int j = jg + 4 * i;
        if (j < d) ar[i] = fmaf(__ldg(wr + j * (d + 1) + k), xv, ar[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxDim / 4; ++i) {
      const int j = jg + 4 * i;
      if (j < d) {
        V[j * LD + p] = fmaxf(au[i], 0.f) + ar[i];
        if (kKeep) V[(kMaxDim + j) * LD + p] = au[i];  // pre-ReLU y0 for the backward mask
      }
    }
  }
  __syncthreads();
}

// shared-memory carve-up (floats) of the forward-only (rollout) kernel
constexpr int XIN_ROWS = kMaxDim + 1;
constexpr int SM_XIN = 0;
constexpr int SM_V = SM_XIN + XIN_ROWS * LD;
constexpr int SM_R1 = SM_V + kMaxDim * LD;
constexpr int SM_R2 = SM_R1 + H0 * LD;
constexpr int SM_R3 = SM_R2 + H1 * LD;
constexpr int SM_STAGE = ((SM_R3 + H2 * LD + 31) / 32) * 32;  // 128-byte aligned
constexpr int SM_BAR = SM_STAGE + STAGES * CHUNK;
constexpr int SM_FWD_FLOATS = SM_BAR + 2 * STAGES;  // mbarriers are 8 bytes each
constexpr int SM_FWD_BYTES = SM_FWD_FLOATS * 4;

}  // namespace tile
}  // namespace socm
