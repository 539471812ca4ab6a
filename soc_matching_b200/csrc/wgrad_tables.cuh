// wgrad_tables.cuh -- work description of the weight-gradient kernels: the record types and the stage / product /
// output tables of wgrad_tc.cu (fp32 scratch, 3xTF32).  wgrad_h.cu (fp16 hi|lo scratch, kind::f16) uses the same
// record types and region offsets with its own tables (one stage fewer).  Both scratches have the same block geometry
// (loss_tc.cuh): tile -> 4 quarters of 32 points -> 59 feature blocks of 32 features = 4 KB; they differ inside a block
// (fp32 K-major rows with the 128-byte swizzle here, fp16 planes in MN-major core matrices there).
#pragma once
#include "loss_tc.cuh"

namespace socm {
namespace tc {

// ---------------------------------------------------------------- work description
// A stage = the operands of one (tile, quarter) that a group of products shares, copied once into shared
// memory: up to four regions  A (16 KB = 128 feature rows) | B (32 KB = 256 rows) | X (4 KB) | X2 (4 KB).
// Every region can serve as the M = 128 operand or as the N operand of a product (the K-major swizzled
// layout is the same for both), so e.g. r1 is at once the N operand of down_1 and, in two halves, the M
// operand of S^T = r1^T d_y0.  Small products ride on the stages of the big ones instead of paying a
// pipeline stage of their own (the kernel is bound by the per-stage latency, not by bytes or MMAs).
struct Load {
  int dst, fb, bytes;  // offset inside the raw stage, first feature block of the scratch quarter, bytes
};
struct Mma {
  int a_off, b_off, N, col;  // M = 128 rows at a_off  x  N rows at b_off  ->  TMEM columns [col, col + N)
};
struct StageDesc {
  int n_load;
  Load ld[4];
  int n_mma;
  Mma mma[4];
};
// One output block of a pass: TMEM columns [col, col + N), lane = output row
struct OutDesc {
  int col, N, kind, layer, row0, rows;
};
enum { OUT_DIRECT = 0, OUT_TRANSPOSED = 1, OUT_XIN = 2, OUT_SMALL = 3, OUT_BIAS = 4, OUT_AUX_S = 5 };

constexpr int WG_A = 0, WG_B = 16384, WG_X = 49152, WG_X2 = 53248;
constexpr int WG_RAW_BYTES = 57344;       // A | B | X | X2 as copied from the scratch
constexpr int WG_LO = WG_RAW_BYTES;       // offset of the "lo" copy inside a stage
constexpr int WG_STAGE_BYTES = 2 * WG_RAW_BYTES;
constexpr int WG_STAGES = 2;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;  // + alignment slack + barriers
constexpr int WG_NT = 320;  // warps 0-3 split + flush, 4 MMA, 5 producer, 6-9 split
constexpr int WG_PREFETCH = 6;  // stages of L2 prefetch lookahead
// The tensor core adds into its fp32 accumulator with truncation: 2640 accumulation steps (all 55
// tiles of a CTA) lose 5e-5 relative, measured against the fp32 FFMA kernel on identical inputs.  The accumulators
// are therefore flushed (red.global.add: round-to-nearest adds in L2) every WG_SEG tiles = 384 steps (7e-6).
constexpr int WG_SEG = 8;
constexpr uint64_t DESC_SW128 = 2ull << 61;  // layout type SWIZZLE_128B
constexpr int FBB = FB_BYTES;

// The stages are grouped into passes whose accumulators fit the 512 TMEM columns together; within a pass the
// loop order is (tile, quarter) outer, stage inner; every accumulator sums over ALL tiles of the CTA and is
// flushed once, at the end of its pass, with red.global.add.
constexpr int N_STAGE_DESC = 7, N_OUT = 19, N_PASS = 3;
static __constant__ int c_pass_stage[N_PASS + 1] = {0, 3, 5, 7};
static __constant__ int c_pass_out[N_PASS + 1] = {0, 8, 14, 19};
static __constant__ StageDesc c_stage[N_STAGE_DESC] = {
    // ---- pass 0: down_1 (+bias), S^T = r1^T d_y0, (d_y0^T y1)^T, down_0
    {4, {{WG_A, FB_DZ2, 4 * FBB}, {WG_B, FB_R1, 8 * FBB}, {WG_X, FB_XIN, FBB}, {WG_X2, FB_DY0, FBB}},
     4, {{WG_A, WG_B, 256, 0}, {WG_A, WG_X, 32, 256}, {WG_B, WG_X2, 32, 288}, {WG_B + 16384, WG_X2, 32, 320}}},
    {2, {{WG_B, FB_Y1, 8 * FBB}, {WG_X2, FB_DY0, FBB}, {0, 0, 0}, {0, 0, 0}},
     2, {{WG_B, WG_X2, 32, 352}, {WG_B + 16384, WG_X2, 32, 384}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    {2, {{WG_B, FB_DZ1, 8 * FBB}, {WG_X, FB_XIN, FBB}, {0, 0, 0}, {0, 0, 0}},
     2, {{WG_B, WG_X, 32, 416}, {WG_B + 16384, WG_X, 32, 448}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    // ---- pass 1: up_1 (both row halves; o2 sits in the A region as the N operand), res_2
    // (an N operand and the XIN block are placed back to back wherever the sum stays <= 256 rows: one MMA of
    //  N + 32 columns reads the M operand once instead of twice)
    {3, {{20480, FB_DY1, 8 * FBB}, {0, FB_O2, 4 * FBB}, {16384, FB_XIN, FBB}, {0, 0, 0}},
     2, {{20480, 0, 160, 0}, {20480 + 16384, 0, 160, 160}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    {3, {{WG_A, FB_DO2, 4 * FBB}, {16384, FB_R2, 4 * FBB}, {32768, FB_XIN, FBB}, {0, 0, 0}},
     1, {{WG_A, 16384, 160, 320}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    // ---- pass 2: up_2, down_2 (D[in][out]) + its bias, and the d-sized layers
    {3, {{WG_A, FB_DY2, 4 * FBB}, {16384, FB_R3, 2 * FBB}, {24576, FB_XIN, FBB}, {0, 0, 0}},
     1, {{WG_A, 16384, 96, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}}},
    // M is always 128 (an M = 64 accumulator is spread over 16 lanes per TMEM quarter): the d_z3 and d_y0 row
    // blocks are loaded with the neighbouring tensors of the scratch, whose rows the flush ignores
    {4, {{WG_A, FB_R2, 4 * FBB}, {WG_B, FB_DZ3, 4 * FBB}, {WG_B + 16384, FB_DY0, 4 * FBB}, {WG_X, FB_XIN, FBB}},
     3, {{WG_A, WG_B, 64, 96}, {WG_B, WG_X, 32, 160}, {WG_B + 16384, WG_X, 32, 192}, {0, 0, 0, 0}}},
};
static __constant__ OutDesc c_out[N_OUT] = {
    {0, 256, OUT_DIRECT, 1, 0, 128},    {256, 32, OUT_BIAS, 1, 0, 128},        // down_1
    {288, 32, OUT_AUX_S, 8, 0, 128},    {320, 32, OUT_AUX_S, 8, 128, 128},     // S^T -> aux (res_1 / up_0 via fold_finish_kernel)
    {352, 32, OUT_TRANSPOSED, 8, 0, 128}, {384, 32, OUT_TRANSPOSED, 8, 128, 128},  // up_0, y1 part
    {416, 32, OUT_XIN, 0, 0, 128},      {448, 32, OUT_XIN, 0, 128, 128},       // down_0 (+ bias via the ones feature)
    {0, 128, OUT_DIRECT, 7, 0, 128},    {128, 32, OUT_BIAS, 7, 0, 128},        // up_1 rows 0..127
    {160, 128, OUT_DIRECT, 7, 128, 128}, {288, 32, OUT_BIAS, 7, 128, 128},     // up_1 rows 128..255
    {320, 128, OUT_DIRECT, 5, 0, 128},  {448, 32, OUT_BIAS, 5, 0, 128},        // res_2
    {0, 64, OUT_DIRECT, 6, 0, 128},     {64, 32, OUT_BIAS, 6, 0, 128},         // up_2
    {96, 64, OUT_TRANSPOSED, 2, 0, 128},                                       // down_2
    {160, 32, OUT_BIAS, 2, 0, 64},                                             // bias of down_2 (rows = d_z3 features)
    {192, 32, OUT_SMALL, 3, 0, 64},     // rows 0..31 d_y0 -> b(up_0), aux sb; rows 32..63 d_o0 -> res_0, b(res_0)
};


}  // namespace tc
}  // namespace socm
