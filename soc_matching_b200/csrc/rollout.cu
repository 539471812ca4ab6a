// rollout.cu -- K1: batched Euler-Maruyama rollout with the control network in the loop.
// Replaces utils.stochastic_trajectories (utils.py:17-128) + NeuralSDE.control (method.py:58-80).
//
//  rollout_tile_kernel    persistent, one CTA per 64-path tile, all K steps inside the kernel:
//                         states and UNet activations stay in shared memory / registers, the
//                         weights are streamed from L2 (unet_tile.cuh), the SDE step, the
//                         stopping logic, the log-weight accumulation and the Philox noise are
//                         fused into every step.  Default hdims only.
//  rollout_generic_kernel one warp per path, any hdims.
#include "kernels.h"
#include "rollout_common.cuh"
#include "unet_generic.cuh"
#include "unet_tile.cuh"

namespace socm {

// ---------------------------------------------------------------- weight repack (once per call)
__global__ void pack_tape_kernel(socm_unet net, float* __restrict__ packed) {
  using namespace tile;
  const int d = net.d;
  const SmallOff so = small_offsets(d);
  float* ft = packed;
  float* bt = packed + FT_FLOATS;
  float* sm = packed + FT_FLOATS + BT_FLOATS;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // transposed copies: dst[k][n] = W[n][k]
  auto transpose = [&](float* dst, const float* W, int nout, int nin) {
    for (int i = tid; i < nout * nin; i += nth) {
      const int k = i / nout, n = i - k * nout;
      dst[i] = W[n * nin + k];
    }
  };
  auto copy = [&](float* dst, const float* src, int n) {
    for (int i = tid; i < n; i += nth) dst[i] = src[i];
  };
  transpose(ft + FT_D1, net.w[1], H1, H0);
  transpose(ft + FT_D2, net.w[2], H2, H1);
  transpose(ft + FT_U2, net.w[6], H1, H2);
  transpose(ft + FT_R2, net.w[5], H1, H1);
  transpose(ft + FT_U1, net.w[7], H0, H1);
  transpose(ft + FT_R1, net.w[4], H0, H0);
  copy(bt + BT_U1, net.w[7], H0 * H1);
  copy(bt + BT_R1, net.w[4], H0 * H0);
  copy(bt + BT_U2, net.w[6], H1 * H2);
  copy(bt + BT_R2, net.w[5], H1 * H1);
  copy(bt + BT_D2, net.w[2], H2 * H1);
  copy(bt + BT_D1, net.w[1], H1 * H0);
  transpose(sm + so.d0t, net.w[0], H0, d + 1);
  copy(sm + so.b_d0, net.b[0], H0);
  copy(sm + so.b_d1, net.b[1], H1);
  copy(sm + so.b_d2, net.b[2], H2);
  copy(sm + so.b_u2, net.b[6], H1);
  copy(sm + so.b_r2, net.b[5], H1);
  copy(sm + so.b_u1, net.b[7], H0);
  copy(sm + so.b_r1, net.b[4], H0);
  copy(sm + so.u0, net.w[8], d * H0);
  copy(sm + so.b_u0, net.b[8], d);
  copy(sm + so.r0, net.w[3], d * (d + 1));
  copy(sm + so.b_r0, net.b[3], d);
}

int pack_tape(const socm_unet* net, float* packed, cudaStream_t stream) {
  pack_tape_kernel<<<64, 256, 0, stream>>>(*net, packed);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

// ---------------------------------------------------------------- generic kernel (warp per path)
__global__ void __launch_bounds__(256) rollout_generic_kernel(RolloutArgs a, socm_unet net) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * (blockDim.x >> 5) + warp;
  if (m >= a.B) return;
  const int d = a.st.d, K = a.K;
  const int per_warp = generic::fwd_floats(d, net.h0, net.h1, net.h2);
  generic::FwdBuf b = generic::carve_fwd(smem + (size_t)warp * per_warp, d, net.h0, net.h1, net.h2);
  PathAcc acc{1.f, 0.f, 0.f};
  if (lane == 0) {
    for (int j = 0; j < d; ++j) {
      const float v = __ldg(a.x0 + (size_t)m * d + j);
      b.xin[1 + j] = v;
      if (a.states) a.states[(size_t)m * d + j] = v;
    }
    if (a.stop) a.stop[m] = 1.f;
  }
  for (int k = 0; k < K; ++k) {
    if (lane == 0) b.xin[0] = __ldg(a.step_tab + 4 * K + k);
    __syncwarp();
    generic::forward(net, b, lane);
    if (lane == 0) path_step(a, m, k, b.xin + 1, 1, b.o0, 1, acc);
    __syncwarp();
  }
  if (lane == 0) path_finish(a, m, b.xin + 1, 1, acc);
}

// ---------------------------------------------------------------- tiled persistent kernel
__global__ void __launch_bounds__(tile::NT, 1) rollout_tile_kernel(RolloutArgs a, const float* __restrict__ packed) {
  using namespace tile;
  extern __shared__ __align__(128) float smem[];
  float* XIN = smem + SM_XIN;
  float* V = smem + SM_V;
  float* R1 = smem + SM_R1;
  float* R2 = smem + SM_R2;
  float* R3 = smem + SM_R3;
  const int d = a.st.d, K = a.K, B = a.B;
  const SmallOff so = small_offsets(d);
  const float* small = packed + FT_FLOATS + BT_FLOATS;
  const int n_tiles = (B + BT - 1) / BT;
  const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  Pipe pipe;
  pipe.start(smem + SM_STAGE, reinterpret_cast<uint64_t*>(smem + SM_BAR), packed, FT_CHUNKS,
             (uint32_t)my_tiles * (uint32_t)K * (uint32_t)FT_CHUNKS);
  const Coord co;
  const int p = threadIdx.x;  // path slot of the stepping threads (p < BT)

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int m = t * BT + p;
    const bool live = p < BT && m < B;
    PathAcc acc{1.f, 0.f, 0.f};
    if (p < BT) {
      for (int j = 0; j < d; ++j) {
        const float v = live ? __ldg(a.x0 + (size_t)m * d + j) : 0.f;
        XIN[(1 + j) * LD + p] = v;
        if (live && a.states) a.states[(size_t)m * d + j] = v;
      }
      if (live && a.stop) a.stop[m] = 1.f;
    }
    for (int k = 0; k < K; ++k) {
      if (p < BT) XIN[p] = __ldg(a.step_tab + 4 * K + k);
      __syncthreads();
      forward_tile<false>(d, small, so, XIN, R1, R2, R3, /*O2=*/R2, /*O1=*/R1, V, pipe, co, nullptr);
      if (live) path_step(a, m, k, XIN + LD + p, LD, V + p, LD, acc);
      // the __syncthreads() at the top of the next step orders these XIN writes
    }
    if (live) path_finish(a, m, XIN + LD + p, LD, acc);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- rollout under a TABULATED control (no network)
// The ground-truth / baseline controls of models.py:10-150, consumed by control_objective and normalization_constant
// (utils.py:131-231, main.py:117-153) through the `else` branch of NeuralSDE.control (method.py:103-107):
//   affine  u_k(x) = A_k x + c_k   LinearControl (models.py:10-39: A_k = u[floor((n-1) t_k / T)], c = 0) and
//                                  ConstantControlLinear (models.py:61-81: A = 0, c_k = ut[floor(n t_k / T)]);
//                                  the caller tabulates A_k, c_k on the grid times with the reference's index rule
//   lookup  u_j(t_k, x) = ut[it_k][clamp(floor((x_j + xb) / dx), 0, nx - 1)][j]      LowDimControl (models.py:84-150),
//                                  it_k = ceil(t_k / delta_t) tabulated by the caller
// One thread per path, state in registers / local memory; the SDE step, stopping logic and weight accumulation are the
// shared sde_step of common.cuh, so every output has the semantics of utils.py:17-128.
__global__ void __launch_bounds__(128) rollout_tab_kernel(RolloutArgs a, socm_tab_control c) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.B) return;
  const int d = a.st.d, K = a.K;
  float x[kMaxDim], u[kMaxDim], eps[kMaxDim];
  PathAcc acc{1.f, 0.f, 0.f};
  for (int j = 0; j < d; ++j) {
    x[j] = __ldg(a.x0 + (size_t)m * d + j);
    if (a.states) a.states[(size_t)m * d + j] = x[j];
  }
  if (a.stop) a.stop[m] = 1.f;
  for (int k = 0; k < K; ++k) {
    if (c.kind == SOCM_CONTROL_AFFINE) {
      for (int i = 0; i < d; ++i) {
        float v = c.c ? __ldg(c.c + (size_t)k * d + i) : 0.f;
        if (c.A)
          for (int j = 0; j < d; ++j) v = fmaf(__ldg(c.A + ((size_t)k * d + i) * d + j), x[j], v);
        u[i] = v;
      }
    } else {
      const int it = __ldg(c.idx_t + k);
      for (int j = 0; j < d; ++j) {
        // floor((x + xb) / delta_x) clamped to the table (models.py:97-105), same fp32 operations
        int ix = (int)floorf(__fdiv_rn(__fadd_rn(x[j], c.xb), c.dx));
        ix = ix < 0 ? 0 : (ix > c.nx - 1 ? c.nx - 1 : ix);
        u[j] = __ldg(c.ut + ((size_t)it * c.nx + ix) * d + j);
      }
    }
    draw_noise(a, m, k, eps);
    const float dt = __ldg(a.step_tab + k), sq_ldt = __ldg(a.step_tab + K + k);
    const float dt_l = __ldg(a.step_tab + 2 * K + k), sq_dtl = __ldg(a.step_tab + 3 * K + k);
    const float eff = sde_step(a.st, nullptr, nullptr, x, 1, nullptr, 1, eps, u, dt, sq_ldt, dt_l, sq_dtl, acc);
    const size_t row = (size_t)k * a.B + m;
    if (a.states) {
      for (int j = 0; j < d; ++j) {
        a.states[(row + a.B) * d + j] = x[j];
        a.controls[row * d + j] = u[j];
      }
      if (a.noise_in == nullptr)
        for (int j = 0; j < d; ++j) a.noises[row * d + j] = eps[j];
      a.stop[row + a.B] = acc.alive;
      a.eff_dt[row] = eff;
    }
  }
  path_finish(a, m, x, 1, acc);
}

// ---------------------------------------------------------------- Philox noise as a stand-alone op
__global__ void philox_normal_kernel(uint64_t seed, uint64_t path_offset, int B, int K, int d, float* out) {
  const int nblk = (d + 3) / 4;
  const size_t total = (size_t)K * B * nblk;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int blk = (int)(i % nblk);
    const size_t km = i / nblk;
    const int m = (int)(km % B), k = (int)(km / B);
    float z[4];
    philox_normal4(seed, path_offset + (uint64_t)m, (uint32_t)k, (uint32_t)blk, z);
    for (int j = 0; j < 4 && blk * 4 + j < d; ++j) out[km * d + blk * 4 + j] = z[j];
  }
}

// ---------------------------------------------------------------- UNet forward at n points (generic)
__global__ void __launch_bounds__(256) unet_forward_generic_kernel(socm_unet net, const float* __restrict__ tx,
                                                                  int n, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, d = net.d;
  const int per_warp = generic::fwd_floats(d, net.h0, net.h1, net.h2);
  generic::FwdBuf b = generic::carve_fwd(smem + (size_t)warp * per_warp, d, net.h0, net.h1, net.h2);
  for (int i = blockIdx.x * (blockDim.x >> 5) + warp; i < n; i += gridDim.x * (blockDim.x >> 5)) {
    for (int j = lane; j <= d; j += 32) b.xin[j] = __ldg(tx + (size_t)i * (d + 1) + j);
    __syncwarp();
    generic::forward(net, b, lane);
    for (int j = lane; j < d; j += 32) out[(size_t)i * d + j] = b.o0[j];
    __syncwarp();
  }
}

}  // namespace socm

// ================================================================ C ABI
using namespace socm;

extern "C" int64_t socm_rollout_workspace_bytes(const socm_unet* net) {
  if (!net) return -1;
  const int64_t ffma = tile::packed_floats(net->d) * (int64_t)sizeof(float);
  int64_t tcb = tc::rollout_tc_supported(net) ? tc::rollout_tc_workspace_bytes(net->d) : 0;
  if (hx::rollout_h_supported(net) && hx::rollout_h_workspace_bytes() > tcb) tcb = hx::rollout_h_workspace_bytes();
  return ffma > tcb ? ffma : tcb;
}

extern "C" int socm_rollout_f32(const socm_setting* st, const socm_unet* net, const socm_warm_table* warm,
                                const float* x0, const float* step_tab, const float* noise_in, uint64_t seed,
                                uint64_t path_offset, int32_t B, int32_t K, float* states, float* noises,
                                float* controls, float* stop, float* eff_dt, float* logw_det, float* logw_sto,
                                float* logw_term, void* workspace, uint32_t flags, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = validate_setting(st)) return rc;
  if (int rc = validate_unet(net, st->d)) return rc;
  SOCM_CHECK_ARG(B >= 0 && K >= 1, "bad sizes B=%d K=%d", B, K);
  SOCM_CHECK_ARG(x0 && step_tab && logw_det && logw_sto && logw_term, "required pointer is NULL");
  const bool no_traj = flags & SOCM_ROLLOUT_NO_TRAJ;
  SOCM_CHECK_ARG(no_traj || (states && controls && stop && eff_dt && (noises || noise_in)),
                 "trajectory outputs are NULL (pass SOCM_ROLLOUT_NO_TRAJ for weights-only mode)");
  SOCM_CHECK_ARG(!warm || (warm->A && warm->c), "warm-start table has NULL members");
  if (B == 0) return SOCM_OK;
  RolloutArgs a;
  a.st = *st;
  a.warmA = warm ? warm->A : nullptr;
  a.warmc = warm ? warm->c : nullptr;
  a.x0 = x0;
  a.step_tab = step_tab;
  a.noise_in = noise_in;
  a.seed = seed;
  a.path_offset = path_offset;
  a.B = B;
  a.K = K;
  a.states = states;
  a.noises = noises;
  a.controls = controls;
  a.stop = stop;
  a.eff_dt = eff_dt;
  a.lw_det = logw_det;
  a.lw_sto = logw_sto;
  a.lw_term = logw_term;

  // fp16-split engine wherever it applies (default architecture, d <= 15): two CTAs per SM at large batches, and the
  // shorter step even for a single resident tile (measured at B = 128, K = 200: 4.25 ms vs 4.92 ms on the 3xTF32 engine).
  // SOCM_F16=0 in the environment or SOCM_ROLLOUT_TF32 select the 3xTF32 engine.
  const bool want_f16 = (flags & SOCM_ROLLOUT_F16) || f16_default() != 0;
  if (want_f16 && hx::rollout_h_supported(net) && !(flags & (SOCM_ROLLOUT_FORCE_GENERIC | SOCM_ROLLOUT_FORCE_FFMA | SOCM_ROLLOUT_TF32))) {
    SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_rollout_workspace_bytes)");
    if (int rc = hx::launch_rollout_h(a, net, workspace, stream)) return rc;
  } else if (tc::rollout_tc_supported(net) && !(flags & (SOCM_ROLLOUT_FORCE_GENERIC | SOCM_ROLLOUT_FORCE_FFMA))) {
    SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_rollout_workspace_bytes)");
    if (int rc = tc::launch_rollout_tc(a, net, workspace, stream)) return rc;
  } else if (is_default_arch(net) && !(flags & SOCM_ROLLOUT_FORCE_GENERIC)) {
    SOCM_CHECK_ARG(workspace != nullptr, "workspace is NULL (socm_rollout_workspace_bytes)");
    float* packed = static_cast<float*>(workspace);
    if (int rc = pack_tape(net, packed, stream)) return rc;
    SOCM_CUDA(cudaFuncSetAttribute(rollout_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   tile::SM_FWD_BYTES));
    const int n_tiles = (B + tile::BT - 1) / tile::BT;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    rollout_tile_kernel<<<grid, tile::NT, tile::SM_FWD_BYTES, stream>>>(a, packed);
    SOCM_LAUNCH_CHECK();
  } else {
    const int warps = 8;
    const size_t smem = (size_t)warps * generic::fwd_floats(st->d, net->h0, net->h1, net->h2) * sizeof(float);
    SOCM_CHECK_ARG(smem <= 200 * 1024, "hidden sizes too large for the generic kernel");
    SOCM_CUDA(cudaFuncSetAttribute(rollout_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rollout_generic_kernel<<<(B + warps - 1) / warps, warps * 32, smem, stream>>>(a, *net);
    SOCM_LAUNCH_CHECK();
  }
  return SOCM_OK;
}

extern "C" int socm_rollout_tabulated_f32(const socm_setting* st, const socm_tab_control* ctrl, const float* x0,
                                          const float* step_tab, const float* noise_in, uint64_t seed,
                                          uint64_t path_offset, int32_t B, int32_t K, float* states, float* noises,
                                          float* controls, float* stop, float* eff_dt, float* logw_det,
                                          float* logw_sto, float* logw_term, uint32_t flags, void* stream_) {
  if (int rc = validate_setting(st)) return rc;
  SOCM_CHECK_ARG(ctrl != nullptr, "control table is NULL");
  SOCM_CHECK_ARG(B >= 0 && K >= 1, "bad sizes B=%d K=%d", B, K);
  SOCM_CHECK_ARG(x0 && step_tab && logw_det && logw_sto && logw_term, "required pointer is NULL");
  const bool no_traj = flags & SOCM_ROLLOUT_NO_TRAJ;
  SOCM_CHECK_ARG(no_traj || (states && controls && stop && eff_dt && (noises || noise_in)),
                 "trajectory outputs are NULL (pass SOCM_ROLLOUT_NO_TRAJ for weights-only mode)");
  if (ctrl->kind == SOCM_CONTROL_AFFINE) {
    SOCM_CHECK_ARG(ctrl->A || ctrl->c, "affine control needs A and / or c");
  } else if (ctrl->kind == SOCM_CONTROL_LOOKUP) {
    SOCM_CHECK_ARG(ctrl->ut && ctrl->idx_t && ctrl->nx >= 1 && ctrl->dx > 0.f, "lookup control needs ut, idx_t, nx, dx");
  } else {
    set_error("unknown control kind %d", ctrl->kind);
    return SOCM_ERR_UNSUPPORTED;
  }
  if (B == 0) return SOCM_OK;
  RolloutArgs a;
  a.st = *st;
  a.warmA = a.warmc = nullptr;
  a.x0 = x0;
  a.step_tab = step_tab;
  a.noise_in = noise_in;
  a.seed = seed;
  a.path_offset = path_offset;
  a.B = B;
  a.K = K;
  a.states = no_traj ? nullptr : states;
  a.noises = noises;
  a.controls = controls;
  a.stop = no_traj ? nullptr : stop;
  a.eff_dt = eff_dt;
  a.lw_det = logw_det;
  a.lw_sto = logw_sto;
  a.lw_term = logw_term;
  rollout_tab_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream_)>>>(a, *ctrl);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_philox_normal_f32(uint64_t seed, uint64_t path_offset, int32_t B, int32_t K, int32_t d,
                                      float* out, void* stream_) {
  SOCM_CHECK_ARG(out && B >= 0 && K >= 0 && d >= 1, "bad arguments");
  if (B == 0 || K == 0) return SOCM_OK;
  const size_t total = (size_t)K * B * ((d + 3) / 4);
  const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  philox_normal_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(seed, path_offset, B, K, d, out);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}

extern "C" int socm_unet_forward_f32(const socm_unet* net, const float* tx, int32_t n, float* out, void* stream_) {
  SOCM_CHECK_ARG(net && tx && out && n >= 0, "bad arguments");
  if (int rc = validate_unet(net, net->d)) return rc;
  if (n == 0) return SOCM_OK;
  const int warps = 8;
  const size_t smem = (size_t)warps * generic::fwd_floats(net->d, net->h0, net->h1, net->h2) * sizeof(float);
  SOCM_CHECK_ARG(smem <= 200 * 1024, "hidden sizes too large for the generic kernel");
  SOCM_CUDA(cudaFuncSetAttribute(unet_forward_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (n + warps - 1) / warps;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  unet_forward_generic_kernel<<<grid, warps * 32, smem, static_cast<cudaStream_t>(stream_)>>>(*net, tx, n, out);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
