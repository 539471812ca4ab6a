// kernels.h -- host-side declarations shared between translation units.
#pragma once
#include "common.cuh"

namespace socm {
// Repack the nn.Linear weights of the default-arch UNet into the forward / backward tapes and
// the small block (unet_tile.cuh); `packed` has tile::packed_floats(d) floats.
int pack_tape(const socm_unet* net, float* packed, cudaStream_t stream);
struct RolloutArgs;
namespace tc {
// tcgen05 rollout (rollout_tc.cu): default hdims and d <= 23
bool rollout_tc_supported(const socm_unet* net);
int64_t rollout_tc_workspace_bytes(int d);
int launch_rollout_tc(const RolloutArgs& a, const socm_unet* net, void* workspace, cudaStream_t stream);
// split (3xTF32 hi/lo) weight tape(s) + small fp32 block of unet_tc.cuh
int pack_tc(const socm_unet* net, unsigned char* tape, bool with_bwd, cudaStream_t stream);
// tcgen05 K3: loss_tc.cu (forward + loss + dgrad) and wgrad_tc.cu (weight gradients)
// `aux`: AUX_FLOATS accumulators of loss_tc.cuh (S = d_y0^T r1, sb = sum d_y0), zeroed by the caller
int launch_wgrad_tc(const unsigned char* scratch, int n_tiles, int d, float* grad, float* aux, cudaStream_t stream);
bool loss_tc_supported(const socm_unet* net);
int64_t loss_tc_workspace_bytes(int d, int B, int K);
}  // namespace tc
namespace hx {
// fp16-split tcgen05 engine, two CTAs per SM (unet_h.cuh): default hdims and d <= 15
bool rollout_h_supported(const socm_unet* net);
int64_t rollout_h_workspace_bytes();
int launch_rollout_h(const RolloutArgs& a, const socm_unet* net, void* workspace, cudaStream_t stream);
// K3a / K3b on the same engine (loss_h.cu, wgrad_h.cu)
bool loss_h_supported(const socm_unet* net);
int64_t loss_h_workspace_bytes(int B, int K);
// K2 forward on kind::f16 (target_h.cu); workspace of socm_target_gemm_tc_workspace_bytes
int64_t target_h_workspace_bytes(int K, int d);
int launch_target_h(const float* L, const float* R, int B, int K, int d, int ldr, float* target, int ldt, void* workspace,
                    cudaStream_t stream);
// K2 backward on kind::f16 (target_bwd_h.cu); workspace of socm_target_gemm_bwd_tc_workspace_bytes
int launch_target_bwd_h(const float* G, const float* R, int B, int K, int d, int ldr, int ldt, float* dL, void* workspace,
                        cudaStream_t stream);
}  // namespace hx
}  // namespace socm
