// loss_tile.cu -- K3 for the default UNet (hdims [256,128,64]): forward at 64 trajectory points,
// weighted loss, dgrad and wgrad in one persistent CTA per SM, FP32 FFMA (parity path).
//
// Per tile (time i, 64 consecutive paths) the CTA
//   * runs the forward of unet_tile.cuh, keeping the ReLU masks of the up layers in registers and
//     spilling r1, r2 (needed again by wgrad) to a private, L2-resident scratch;
//   * evaluates the per-point loss / d loss / d nabla_V (loss_common.cuh) and writes G;
//   * walks the network backwards with two 256-row shared-memory buffers PA/PB:
//       dgrad  = the same register-tile kernel as forward, fed by the backward weight tape
//                (native nn.Linear layout, streamed with cp.async.bulk);
//       wgrad  = register-tiled  gW[n][c] += sum_p delta[n][p] act[c][p]  over the 64 points,
//                accumulated by plain read-modify-write into a gradient buffer PRIVATE to the
//                CTA (no atomics, deterministic), reduced over CTAs by reduce_grad_kernel.
// Buffer schedule (rows of PA | PB), see DESIGN.md "K3":
//   fwd : r1 -> PA          r2 -> PB[0:128]  r3 -> PB[128:192]   o2 -> PB[0:128]   o1 -> PA
//   bwd : d_y1 -> PA        d_o2 -> PB[0:128], d_y2 -> PA[0:128] d_z3 -> PB[192:256]
//         r2 -> PA[128:256] d_z2 -> PA[0:128]  r1 -> PB          d_o1 -> PA        d_z1 -> PA
#include "kernels.h"
#include "loss_common.cuh"
#include "unet_tile.cuh"

namespace socm {
using namespace tile;

// ---------------------------------------------------------------- shared-memory carve-up (floats)
constexpr int V_ROWS = 2 * kMaxDim;  // rows [0,32): nabla_V then d_o0; rows [32,64): y0 then d_y0
constexpr int LS_XIN = 0;
constexpr int LS_V = LS_XIN + XIN_ROWS * LD;
constexpr int LS_PA = LS_V + V_ROWS * LD;
constexpr int LS_PB = LS_PA + H0 * LD;
constexpr int LS_STAGE = ((LS_PB + H0 * LD + 31) / 32) * 32;
constexpr int LS_BAR = LS_STAGE + STAGES * CHUNK;
constexpr int LS_FLOATS = LS_BAR + 2 * STAGES + 2;
constexpr int LS_BYTES = LS_FLOATS * 4;
static_assert(LS_BYTES <= 227 * 1024, "K3 shared memory budget");

// per-CTA global scratch (floats): r1 [256][64], r2 [128][64]
constexpr int SCR_R1 = 0;
constexpr int SCR_R2 = SCR_R1 + H0 * BT;
constexpr int SCR_FLOATS = SCR_R2 + H1 * BT;

// flat gradient layout (= socm_unet layer order, w then b)
struct GradOff {
  int w[9], b[9], total;
};
__host__ __device__ inline GradOff grad_offsets(int d) {
  const int nout[9] = {H0, H1, H2, d, H0, H1, H1, H0, d};
  const int nin[9] = {d + 1, H0, H1, d + 1, H0, H1, H2, H1, H0};
  GradOff g;
  int p = 0;
  for (int l = 0; l < 9; ++l) {
    g.w[l] = p;
    p += nout[l] * nin[l];
    g.b[l] = p;
    p += nout[l];
  }
  g.total = p;
  return g;
}

// ---------------------------------------------------------------- wgrad micro kernel
// gw[n][c] += sum_p D[n][p] * A[c][p]   (D: N rows, A: C rows of the shared tiles, p = 64 points)
// warp tile 32 n x 64 c (lanes 4 x 8, 8 x 8 outputs per thread, rows interleaved so that the float4
// reads along p are bank-conflict free); warps tile [32*WN] x [64*WC].
constexpr int PASS_FLOATS = NT * 64;  // one wgrad pass = 64 accumulators per thread

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Where element (n, c) of an N x C weight gradient lives inside its thread-major private block
// (inverse of the mapping used by wgrad_tile); shared with reduce_grad_kernel.
template <int N, int C>
__host__ __device__ inline int wgrad_priv_index(int n, int c) {
  constexpr int WC = C / 64, WN = 8 / WC;
  const int pass = n / (32 * WN), nr = n % (32 * WN);
  const int wn = nr / 32, ln = nr % 4, i = (nr % 32) / 4;
  const int wc = c / 64, lc = c % 8, j = (c % 64) / 8;
  const int t = (wn * WC + wc) * 32 + ln * 8 + lc;
  const int k4 = 2 * i + j / 4;
  return pass * PASS_FLOATS + (k4 * NT + t) * 4 + (j % 4);
}
template <int N, int C>
__host__ __device__ constexpr int wgrad_priv_floats() {
  return ((N + 32 * (8 / (C / 64)) - 1) / (32 * (8 / (C / 64)))) * PASS_FLOATS;
}

// CTA-private accumulation buffer: [flat part in GradOff layout (small layers, biases; the slots
// of the six big weight matrices are unused)] [six thread-major wgrad blocks].
struct PrivOff {
  int d1, d2, r1, r2, u2, u1, total;
};
__host__ __device__ inline PrivOff priv_offsets(int d) {
  PrivOff o;
  int p = ((grad_offsets(d).total + 31) / 32) * 32;
  o.d1 = p; p += wgrad_priv_floats<H1, H0>();
  o.d2 = p; p += wgrad_priv_floats<H2, H1>();
  o.r1 = p; p += wgrad_priv_floats<H0, H0>();
  o.r2 = p; p += wgrad_priv_floats<H1, H1>();
  o.u2 = p; p += wgrad_priv_floats<H1, H2>();
  o.u1 = p; p += wgrad_priv_floats<H0, H1>();
  o.total = p;
  return o;
}
// per-CTA pitch of the private buffers
__host__ __device__ inline int grad_stride(int d) { return priv_offsets(d).total; }
// private offset of flat gradient element i
__device__ __forceinline__ int priv_slot(const GradOff& go, const PrivOff& po, int i) {
  if (i >= go.w[1] && i < go.b[1]) { const int e = i - go.w[1]; return po.d1 + wgrad_priv_index<H1, H0>(e / H0, e % H0); }
  if (i >= go.w[2] && i < go.b[2]) { const int e = i - go.w[2]; return po.d2 + wgrad_priv_index<H2, H1>(e / H1, e % H1); }
  if (i >= go.w[4] && i < go.b[4]) { const int e = i - go.w[4]; return po.r1 + wgrad_priv_index<H0, H0>(e / H0, e % H0); }
  if (i >= go.w[5] && i < go.b[5]) { const int e = i - go.w[5]; return po.r2 + wgrad_priv_index<H1, H1>(e / H1, e % H1); }
  if (i >= go.w[6] && i < go.b[6]) { const int e = i - go.w[6]; return po.u2 + wgrad_priv_index<H1, H2>(e / H2, e % H2); }
  if (i >= go.w[7] && i < go.b[7]) { const int e = i - go.w[7]; return po.u1 + wgrad_priv_index<H0, H1>(e / H1, e % H1); }
  return i;
}

template <int N, int C>
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ D, const float* __restrict__ A,
                                           float* __restrict__ gw) {
  constexpr int WC = C / 64;   // warps along c
  constexpr int WN = 8 / WC;   // warps along n
  static_assert(C == 64 || C == 128 || C == 256, "unsupported C");
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wc = w % WC, wn = w / WC;
  const int ln = lane >> 3, lc = lane & 7;
  const int c0 = wc * 64 + lc;  // columns c0 + 8 j
  for (int nb = wn * 32; nb < N; nb += 32 * WN) {
    const int n0 = nb + ln;     // rows n0 + 4 i
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
    for (int p = 0; p < BT; p += 4) {
      float4 av[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) av[j] = *reinterpret_cast<const float4*>(A + (c0 + 8 * j) * LD + p);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 dv = *reinterpret_cast<const float4*>(D + (n0 + 4 * i) * LD + p);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][j] = fmaf(dv.x, av[j].x, acc[i][j]);
          acc[i][j] = fmaf(dv.y, av[j].y, acc[i][j]);
          acc[i][j] = fmaf(dv.z, av[j].z, acc[i][j]);
          acc[i][j] = fmaf(dv.w, av[j].w, acc[i][j]);
        }
      }
    }
    // accumulate into the CTA-private buffer: thread-major layout (float4 k4 of thread t at
    // (k4 * 256 + t) * 4), so every warp-wide REDG.F32x4 is one contiguous 512-byte request and
    // nothing is loaded back (fire-and-forget reduction at L2, no scoreboard stall)
    float* base = gw + (size_t)((nb - wn * 32) / (32 * WN)) * PASS_FLOATS + threadIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h)
        red_add_v4(base + (2 * i + h) * (NT * 4), acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2],
                   acc[i][4 * h + 3]);
  }
}

// gb[r] += sum_p D[r][p]
template <int N>
__device__ __forceinline__ void bias_grad(const float* __restrict__ D, float* __restrict__ gb) {
  for (int r = threadIdx.x; r < N; r += NT) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < BT; p += 4) {
      const float4 v = *reinterpret_cast<const float4*>(D + r * LD + p);
      s += (v.x + v.y) + (v.z + v.w);
    }
    __stcg(gb + r, __ldcg(gb + r) + s);
  }
}

// acc[i][j] = 0 where the activation R[feat][point] <= 0   (ReLU mask of the down layers)
template <int N>
__device__ __forceinline__ void mask_by_activation(TileAcc<N>& acc, const float* __restrict__ R, const Coord& co) {
#pragma unroll
  for (int j = 0; j < TileAcc<N>::TN; ++j) {
    const float* row = R + co.feat<N>(j) * LD + co.pA;
    const float4 a = *reinterpret_cast<const float4*>(row);
    const float4 b = *reinterpret_cast<const float4*>(row + 16);
    const float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (!(r[i] > 0.f)) acc.v[i][j] = 0.f;
  }
}

// ---------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NT, 1) loss_tile_kernel(LossArgs a, const float* __restrict__ packed,
                                                          float* __restrict__ priv_all, float* __restrict__ scratch_all) {
  extern __shared__ __align__(128) float smem[];
  float* XIN = smem + LS_XIN;
  float* V = smem + LS_V;
  float* PA = smem + LS_PA;
  float* PB = smem + LS_PB;
  const int d = a.st.d, K = a.K, B = a.B;
  const SmallOff so = small_offsets(d);
  const GradOff go = grad_offsets(d);
  const PrivOff po = priv_offsets(d);
  const float* small = packed + FT_FLOATS + BT_FLOATS;
  float* priv = priv_all + (size_t)blockIdx.x * grad_stride(d);
  float* scr = scratch_all + (size_t)blockIdx.x * SCR_FLOATS;
  const int n_mblk = (B + BT - 1) / BT;
  const int n_tiles = (K + 1) * n_mblk;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tid = threadIdx.x;

  for (int i = tid; i < po.total; i += NT) __stcg(priv + i, 0.f);  // private gradient accumulator
  Pipe pipe;  // forward tape followed by backward tape = one contiguous 80-chunk tape
  pipe.start(smem + LS_STAGE, reinterpret_cast<uint64_t*>(smem + LS_BAR), packed, FT_CHUNKS + BT_CHUNKS,
             (uint32_t)my_tiles * (uint32_t)(FT_CHUNKS + BT_CHUNKS));
  const Coord co;
  double loss_acc = 0.0;

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int ti = t / n_mblk, m0 = (t - ti * n_mblk) * BT;
    // ---- inputs: XIN row 0 = t_i, rows 1..d = x
    if (tid < BT) XIN[tid] = __ldg(a.ts + ti);
    for (int idx = tid; idx < BT * d; idx += NT) {
      const int p = idx / d, j = idx - p * d;
      const int m = m0 + p;
      XIN[(1 + j) * LD + p] = m < B ? __ldg(a.states + ((size_t)ti * B + m) * d + j) : 0.f;
    }
    __syncthreads();

    // ---- forward (r1 -> PA, r2 -> PB[0:128], r3 -> PB[128:192], o2 -> PB[0:128], o1 -> PA)
    FwdMasks masks;
    forward_tile<true>(d, small, so, XIN, PA, PB, PB + H1 * LD, PB, PA, V, pipe, co, &masks, scr + SCR_R1,
                       scr + SCR_R2);

    // ---- loss, d_o0 (V rows 0..d), d_y0 (V rows 32..32+d), G
    if (tid < BT) {
      const int m = m0 + tid;
      float dv[kMaxDim];
      if (m < B) {
        loss_acc += (double)point_loss(a, ti, m, XIN + LD + tid, LD, V + tid, LD, dv);
      } else {
        for (int j = 0; j < d; ++j) dv[j] = 0.f;
      }
      for (int j = 0; j < d; ++j) {
        const float y0 = V[(kMaxDim + j) * LD + tid];
        V[j * LD + tid] = dv[j];
        V[(kMaxDim + j) * LD + tid] = y0 > 0.f ? dv[j] : 0.f;
      }
    }
    __syncthreads();
    const float* DO0 = V;
    const float* DY0 = V + kMaxDim * LD;

    // ---- b0: small wgrads that need o1 (PA), XIN
    {
      // up_0: gW[j][c] += sum_p d_y0[j][p] o1[c][p]; thread <-> feature c
      float accj[kMaxDim];
#pragma unroll
      for (int j = 0; j < kMaxDim; ++j) accj[j] = 0.f;
      const float* orow = PA + tid * LD;
      for (int p = 0; p < BT; ++p) {
        const float o = orow[p];
#pragma unroll
        for (int j = 0; j < kMaxDim; ++j)
          if (j < d) accj[j] = fmaf(DY0[j * LD + p], o, accj[j]);
      }
#pragma unroll
      for (int j = 0; j < kMaxDim; ++j)
        if (j < d) {
          float* g = priv + go.w[8] + j * H0 + tid;
          __stcg(g, __ldcg(g) + accj[j]);
        }
      // res_0: gW[j][k] += sum_p d_o0[j][p] xin[k][p]; biases of up_0 / res_0
      for (int o = tid; o < d * (d + 1); o += NT) {
        const int j = o / (d + 1), k = o - j * (d + 1);
        float s = 0.f;
        for (int p = 0; p < BT; ++p) s = fmaf(DO0[j * LD + p], XIN[k * LD + p], s);
        float* g = priv + go.w[3] + o;
        __stcg(g, __ldcg(g) + s);
      }
      if (tid < 2 * d) {
        const int j = tid % d, which = tid / d;  // 0: up_0 bias (d_y0), 1: res_0 bias (d_o0)
        const float* src = which == 0 ? DY0 : DO0;
        float s = 0.f;
        for (int p = 0; p < BT; ++p) s += src[j * LD + p];
        float* g = priv + (which == 0 ? go.b[8] : go.b[3]) + j;
        __stcg(g, __ldcg(g) + s);
      }
    }
    __syncthreads();  // all reads of o1 done

    // ---- b1: d_y1 = mask_y1 . (up_0^T d_y0) -> PA
    {
      TileAcc<256> t;
      t.zero();
      mac_chunk<256, 1, true>(t, DY0, small + so.u0, co, d);
      apply_mask<256>(t, masks.y1);
      store_tile<256>(t, PA, co);
    }
    __syncthreads();
    // ---- b2: up_1 wgrad: gW[256][128] += d_y1 (PA) x o2 (PB[0:128])
    wgrad_tile<256, 128>(PA, PB, priv + po.u1);
    bias_grad<256>(PA, priv + go.b[7]);
    // ---- b3: d_o2 = up_1^T d_y1 -> PB[0:128]; d_y2 = mask_y2 . d_o2 -> PA[0:128]
    {
      TileAcc<128> t;
      t.zero();
      stream_layer<128, H0>(t, PA, pipe, co);  // last release() = all reads of PA and (b2) PB done
      store_tile<128>(t, PB, co);
      apply_mask<128>(t, (unsigned long long)masks.y2);
      store_tile<128>(t, PA, co);
    }
    __syncthreads();
    // ---- b4: up_2 wgrad: gW[128][64] += d_y2 (PA[0:128]) x r3 (PB[128:192]); res_2 bias
    wgrad_tile<128, 64>(PA, PB + H1 * LD, priv + po.u2);
    bias_grad<128>(PA, priv + go.b[6]);
    bias_grad<128>(PB, priv + go.b[5]);
    // ---- b5: d_z3 = relu'(r3) . (up_2^T d_y2) -> PB[192:256]
    {
      TileAcc<64> t;
      t.zero();
      stream_layer<64, H1>(t, PA, pipe, co);
      mask_by_activation<64>(t, PB + H1 * LD, co);
      store_tile<64>(t, PB + (H1 + H2) * LD, co);
    }
    // ---- b6: r2 -> PA[128:256]; res_2 / down_2 wgrads
    reload_rows(PA + H1 * LD, scr + SCR_R2, H1);
    __syncthreads();
    wgrad_tile<128, 128>(PB, PA + H1 * LD, priv + po.r2);
    wgrad_tile<64, 128>(PB + (H1 + H2) * LD, PA + H1 * LD, priv + po.d2);
    bias_grad<64>(PB + (H1 + H2) * LD, priv + go.b[2]);
    // ---- b7: d_z2 = relu'(r2) . (res_2^T d_o2 + down_2^T d_z3) -> PA[0:128]
    {
      TileAcc<128> t;
      t.zero();
      stream_layer<128, H1>(t, PB, pipe, co);
      stream_layer<128, H2>(t, PB + (H1 + H2) * LD, pipe, co);
      mask_by_activation<128>(t, PA + H1 * LD, co);
      store_tile<128>(t, PA, co);  // d_y2 is dead: every thread passed the last release()
    }
    // ---- b8: r1 -> PB; down_1 wgrad
    reload_rows(PB, scr + SCR_R1, H0);
    __syncthreads();
    wgrad_tile<128, 256>(PA, PB, priv + po.d1);
    bias_grad<128>(PA, priv + go.b[1]);
    // ---- b9..b12: d_z1 = relu'(r1) . (down_1^T d_z2 + res_1^T d_o1)
    {
      TileAcc<256> t;
      t.zero();
      stream_layer<256, H1>(t, PA, pipe, co);  // down_1^T d_z2; the last release() frees PA
      {
        TileAcc<256> o;  // d_o1 = up_0^T d_y0 recomputed (unmasked this time)
        o.zero();
        mac_chunk<256, 1, true>(o, DY0, small + so.u0, co, d);
        store_tile<256>(o, PA, co);
      }
      __syncthreads();
      wgrad_tile<256, 256>(PA, PB, priv + po.r1);
      bias_grad<256>(PA, priv + go.b[4]);
      stream_layer<256, H0>(t, PA, pipe, co);  // + res_1^T d_o1
      mask_by_activation<256>(t, PB, co);
      store_tile<256>(t, PA, co);
    }
    __syncthreads();
    // ---- b13: down_0 wgrad: gW[n][k] += sum_p d_z1[n][p] xin[k][p]; thread <-> feature n
    {
      float acck[kMaxDim + 1];
#pragma unroll
      for (int k = 0; k <= kMaxDim; ++k) acck[k] = 0.f;
      float bsum = 0.f;
      const float* zrow = PA + tid * LD;
      for (int p = 0; p < BT; ++p) {
        const float z = zrow[p];
        bsum += z;
#pragma unroll
        for (int k = 0; k <= kMaxDim; ++k)
          if (k <= d) acck[k] = fmaf(z, XIN[k * LD + p], acck[k]);
      }
#pragma unroll
      for (int k = 0; k <= kMaxDim; ++k)
        if (k <= d) {
          float* g = priv + go.w[0] + tid * (d + 1) + k;
          __stcg(g, __ldcg(g) + acck[k]);
        }
      float* gb = priv + go.b[0] + tid;
      __stcg(gb, __ldcg(gb) + bsum);
    }
    __syncthreads();  // XIN / V / PA are rewritten by the next tile
  }

  // ---- loss: warp-shuffle block reduction, one fp64 atomic per CTA
  __shared__ double red[NT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = loss_acc;
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int wv = 0; wv < NT / 32; ++wv) s += red[wv];
    if (s != 0.0) atomicAdd(a.loss_sums, s);
  }
}

// grad[i] += sum over CTAs of their private slot of element i   (fixed order: deterministic)
__global__ void __launch_bounds__(256) reduce_grad_kernel(const float* __restrict__ priv_all, int n_cta, int d,
                                                          float* __restrict__ grad) {
  const GradOff go = grad_offsets(d);
  const PrivOff po = priv_offsets(d);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < go.total; i += gridDim.x * blockDim.x) {
    const int slot = priv_slot(go, po, i);
    float s = 0.f;
    for (int c = 0; c < n_cta; ++c) s += __ldcg(priv_all + (size_t)c * po.total + slot);
    grad[i] += s;
  }
}

static int loss_tile_grid(int B, int K) {
  const long long n_tiles = (long long)(K + 1) * ((B + BT - 1) / BT);
  const int sms = sm_count();
  return (int)(n_tiles < sms ? n_tiles : sms);
}

int64_t loss_tile_workspace_bytes(int d, int B, int K) {
  (void)B;
  (void)K;
  const int64_t ctas = sm_count();  // upper bound of the grid
  return (packed_floats(d) + ctas * (grad_stride(d) + SCR_FLOATS) + 64) * (int64_t)sizeof(float);
}

int launch_loss_tile(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream) {
  const int d = a.st.d;
  SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_loss_workspace_bytes)");
  float* packed = static_cast<float*>(workspace);
  const int64_t pf = ((packed_floats(d) + 31) / 32) * 32;
  const int grid = loss_tile_grid(a.B, a.K);
  float* priv = packed + pf;
  float* scratch = priv + (size_t)sm_count() * grad_stride(d);
  if (int rc = pack_tape(net, packed, stream)) return rc;
  SOCM_CUDA(cudaFuncSetAttribute(loss_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_BYTES));
  loss_tile_kernel<<<grid, NT, LS_BYTES, stream>>>(a, packed, priv, scratch);
  SOCM_LAUNCH_CHECK();
  const int n = grad_offsets(d).total;
  reduce_grad_kernel<<<(n + 255) / 256, 256, 0, stream>>>(priv, grid, d, grad);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

}  // namespace socm
