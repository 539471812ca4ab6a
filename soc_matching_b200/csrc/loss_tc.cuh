// loss_tc.cuh -- shared definitions of the tensor-core K3 (loss_tc.cu = K3a: forward + loss + dgrad,
// wgrad_tc.cu = K3b: weight gradients).
//
// K3a hands K3b every operand of the weight-gradient GEMMs  dW[out][in] = sum_p dY[p][out] Act[p][in]
// through a global scratch in exactly the shared-memory layout tcgen05.mma wants for a K-major tf32
// operand with the 128-byte swizzle (K = trajectory points):
//   a "feature block" = 32 consecutive features of one tensor for 32 consecutive points (a TMEM lane
//   quarter = one warp of K3a) = 32 rows of 128 bytes; row f holds the 32 points of feature f, its
//   eight 16-byte chunks XOR-permuted by (f & 7).  With thread <-> point, one warp-wide 4-byte store
//   of a feature is exactly one 128-byte line (fully coalesced), and one cp.async.bulk per operand
//   lands a run of feature blocks in shared memory ready for the MMA.  Values are plain fp32; K3b
//   splits them into hi/lo in shared memory and runs the GEMMs as 3xTF32 (a single-pass TF32 weight
//   gradient was measured at 1e-4 relative error on a 303-point case: not enough margin).
//   tile (128 points) -> 4 quarters -> NFB feature blocks of 4 KB.
#pragma once
#include "common.cuh"

namespace socm {
namespace tc {

// feature-block index of each tensor inside a quarter
enum Fb {
  FB_XIN = 0,   // [t, x_0..x_{d-1}, 0.., 1 at feature 31]
  FB_R1 = 1,    // 8 blocks
  FB_R2 = 9,    // 4
  FB_R3 = 13,   // 2
  FB_O2 = 15,   // 4
  FB_Y1 = 19,   // 8   y1 = relu(up_1 o2): res_1 is folded into up_0 (unet_tc.cuh), o1 / d_o1 are never formed
  FB_DY0 = 27,  // d_y0 (features >= d are zero)
  FB_DO0 = 28,  // d_o0
  FB_DY1 = 29,  // 8
  FB_DY2 = 37,  // 4
  FB_DO2 = 41,  // 4
  FB_DZ3 = 45,  // 2
  FB_DZ2 = 47,  // 4
  FB_DZ1 = 51,  // 8
  NFB = 59
};
constexpr int FB_BYTES = 4096;
constexpr int QUARTER_BYTES = NFB * FB_BYTES;
constexpr int64_t TILE_BYTES = 4LL * QUARTER_BYTES;  // 966 656
constexpr int ONES_FEATURE = 31;                     // constant-1 feature of the FB_XIN block (bias gradients)
constexpr int MAX_D_TC_LOSS = 30;

// byte offset of (feature f, point r), both 0..31, inside a feature block
__host__ __device__ inline int fb_off(int r, int f) { return f * 128 + ((((r >> 2) ^ (f & 7)) & 7) << 4) + (r & 3) * 4; }

// K3b accumulates the two small products the folding needs into an aux block of the workspace
//   S[j][g] = sum_p d_y0[p][j] r1[p][g]  (32 x 256, rows >= d zero)   and   sb[j] = sum_p d_y0[p][j]  (32)
// and fold_finish_kernel (loss_tc.cu) turns them into the gradients of res_1 and the S-part of up_0.
constexpr int AUX_S = 0, AUX_SB = 32 * 256, AUX_FLOATS = 32 * 256 + 32;

// flat gradient layout (= socm_unet layer order, w then b per layer; same as loss_tile.cu)
struct GradOffTc {
  int w[9], b[9], total;
};
__host__ __device__ inline GradOffTc grad_offsets_tc(int d) {
  const int nout[9] = {256, 128, 64, d, 256, 128, 128, 256, d};
  const int nin[9] = {d + 1, 256, 128, d + 1, 256, 128, 64, 128, 256};
  GradOffTc g;
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

}  // namespace tc
}  // namespace socm
