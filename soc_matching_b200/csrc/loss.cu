// loss.cu -- K3: UNet forward at all (K+1)B trajectory points, importance-weighted squared
// error against the SOCM target, and the backward pass, in one kernel.
// Replaces method.py:272-287 (nabla_V at all points, warm-start correction), 692-720 (loss)
// and the autograd backward of main.py:323 for the UNet parameters and for the target.
//
//   loss_generic_kernel  one warp per point, any hdims; parameter gradients by global atomics
//   (the tiled kernel for the default hdims lives in loss_tile.cu)
#include "kernels.h"
#include "loss_common.cuh"
#include "unet_generic.cuh"

namespace socm {

__global__ void __launch_bounds__(128) loss_generic_kernel(LossArgs a, socm_unet net, float* __restrict__ grad) {
  extern __shared__ __align__(128) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int d = a.st.d, h0 = net.h0, h1 = net.h1, h2 = net.h2;
  const int per_warp = generic::fwd_floats(d, h0, h1, h2) + generic::bwd_floats(d, h0, h1, h2);
  float* base = smem + (size_t)warp * per_warp;
  generic::FwdBuf b = generic::carve_fwd(base, d, h0, h1, h2);
  float* bw = base + generic::fwd_floats(d, h0, h1, h2);
  const generic::GradPtrs g = generic::grad_ptrs(grad, d, h0, h1, h2);
  const size_t total = (size_t)(a.K + 1) * a.B;
  double loss_acc = 0.0;
  for (size_t q = (size_t)blockIdx.x * nwarp + warp; q < total; q += (size_t)gridDim.x * nwarp) {
    const int i = (int)(q / a.B), m = (int)(q - (size_t)i * a.B);
    if (lane == 0) b.xin[0] = __ldg(a.ts + i);
    for (int j = lane; j < d; j += 32) b.xin[1 + j] = __ldg(a.states + ((size_t)i * a.B + m) * d + j);
    __syncwarp();
    generic::forward(net, b, lane);
    if (lane == 0) loss_acc += (double)point_loss(a, i, m, b.xin + 1, 1, b.o0, 1, bw);
    __syncwarp();
    generic::backward(net, b, bw, g, lane);
  }
  if (lane == 0 && loss_acc != 0.0) atomicAdd(a.loss_sums, loss_acc);
}

int launch_loss_tile(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream);
int64_t loss_tile_workspace_bytes(int d, int B, int K);
namespace tc {
int launch_loss_tc(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream);
}
namespace hx {
int launch_loss_h(const LossArgs& a, const socm_unet* net, float* grad, void* workspace, cudaStream_t stream);
}

}  // namespace socm

// ================================================================ C ABI
using namespace socm;

extern "C" int64_t socm_loss_workspace_bytes(const socm_unet* net, int32_t B, int32_t K) {
  if (!net) return -1;
  if (is_default_arch(net)) {
    const int64_t ffma = loss_tile_workspace_bytes(net->d, B, K);
    int64_t tcb = tc::loss_tc_supported(net) ? tc::loss_tc_workspace_bytes(net->d, B, K) : 0;
    if (hx::loss_h_supported(net) && hx::loss_h_workspace_bytes(B, K) > tcb) tcb = hx::loss_h_workspace_bytes(B, K);
    return ffma > tcb ? ffma : tcb;
  }
  return 256;  // the generic kernel needs no workspace
}

extern "C" int socm_unet_loss_fwdbwd_f32(const socm_setting* st, const socm_unet* net, const socm_warm_table* warm,
                                         const float* ts, const float* states, const float* target, int32_t ldt,
                                         const float* w, const float* stop, float scale, int32_t B, int32_t K,
                                         float* G, float* grad, double* loss_sums, void* workspace, uint32_t flags,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (int rc = validate_setting(st)) return rc;
  if (int rc = validate_unet(net, st->d)) return rc;
  SOCM_CHECK_ARG(ts && states && target && w && G && grad && loss_sums, "required pointer is NULL");
  SOCM_CHECK_ARG(B >= 0 && K >= 1 && ldt >= (K + 1) * st->d, "bad sizes B=%d K=%d ldt=%d", B, K, ldt);
  SOCM_CHECK_ARG(!warm || (warm->A && warm->c), "warm-start table has NULL members");
  if (B == 0) return SOCM_OK;
  LossArgs a;
  a.st = *st;
  a.warmA = warm ? warm->A : nullptr;
  a.warmc = warm ? warm->c : nullptr;
  a.ts = ts;
  a.states = states;
  a.target = target;
  a.w = w;
  a.stop = stop;
  a.scale = scale;
  a.B = B;
  a.K = K;
  a.ldt = ldt;
  a.G = G;
  a.loss_sums = loss_sums;
  // tcgen05 kernels for large batches; the fp32 FFMA tile kernel below that (the
  // 3xTF32 forward differs from fp32 by ~2e-6, enough to flip a ReLU mask about once per 5e5
  // pre-activations, which is visible in the gradients of small problems)
  const bool want_tc = (flags & SOCM_LOSS_FORCE_TC) || (int64_t)(K + 1) * B >= SOCM_LOSS_TC_MIN_POINTS;
  // fp16-split engine (two CTAs per SM, csrc/loss_h.cu) wherever it applies; SOCM_F16=0 / SOCM_LOSS_TF32 select the 3xTF32 one
  const bool want_f16 = (flags & SOCM_LOSS_F16) || f16_default() != 0;
  if (hx::loss_h_supported(net) && want_tc && want_f16 &&
      !(flags & (SOCM_LOSS_FORCE_GENERIC | SOCM_LOSS_FORCE_FFMA | SOCM_LOSS_TF32)))
    return hx::launch_loss_h(a, net, grad, workspace, stream);
  if (tc::loss_tc_supported(net) && want_tc && !(flags & (SOCM_LOSS_FORCE_GENERIC | SOCM_LOSS_FORCE_FFMA)))
    return tc::launch_loss_tc(a, net, grad, workspace, stream);
  if (is_default_arch(net) && !(flags & SOCM_LOSS_FORCE_GENERIC)) return launch_loss_tile(a, net, grad, workspace, stream);
  const int warps = 4;
  const size_t smem = (size_t)warps *
                      (generic::fwd_floats(st->d, net->h0, net->h1, net->h2) +
                       generic::bwd_floats(st->d, net->h0, net->h1, net->h2)) *
                      sizeof(float);
  SOCM_CHECK_ARG(smem <= 200 * 1024, "hidden sizes too large for the generic kernel");
  SOCM_CUDA(cudaFuncSetAttribute(loss_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t total = (size_t)(K + 1) * B;
  size_t grid = (total + warps - 1) / warps;
  const size_t cap = (size_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  loss_generic_kernel<<<(int)grid, warps * 32, smem, stream>>>(a, *net, grad);
  SOCM_LAUNCH_CHECK();
  return SOCM_OK;
}
