/*
 * socm_b200.h -- C ABI of the B200-native SOC-matching hot path (libsocm_b200.so).
 *
 * This is the drop-in boundary.  The reference (facebookresearch/SOC-matching) is pure
 * Python/PyTorch and has no FFI of its own; the entry points below are what a binding for
 * its hot path has to call, one per reference call site (file:line relative to the
 * reference tree).  INTEGRATION.md shows the ctypes stub a maintainer of the reference
 * would add in SOC_matching/utils.py and SOC_matching/method.py.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers (fp32 unless noted), sizes, a cudaStream_t passed as
 *     void*; no torch types.  Every launcher is asynchronous on `stream`, never allocates,
 *     never synchronises; workspaces are passed in by the caller.
 *   - return value 0 = ok, otherwise a socm_status code; socm_last_error() gives the text.
 *   - all matrices row-major.  "paths" = trajectories (B), "steps" = K = num_steps,
 *     d = state dimension, grid times ts[0..K].
 *   - network weights are in torch.nn.Linear layout: W[out][in], b[out].
 */
#ifndef SOCM_B200_H_
#define SOCM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOCM_MAX_DIM 32 /* largest supported state dimension d */

typedef enum {
  SOCM_OK = 0,
  SOCM_ERR_INVALID = 1,     /* bad argument (null pointer, d > SOCM_MAX_DIM, ...) */
  SOCM_ERR_UNSUPPORTED = 2, /* setting / shape the kernels do not cover */
  SOCM_ERR_CUDA = 3         /* a CUDA runtime call failed */
} socm_status;

/* experiment_settings/{OU_quadratic,OU_linear,double_well,molecular_dynamics}.py */
typedef enum {
  SOCM_OU_QUADRATIC = 0,      /* b=Ax, f=x'Px, g=x'Qx            OU_quadratic.py:51-83 */
  SOCM_OU_LINEAR = 1,         /* b=Ax, f=0,    g=omega.x         OU_linear.py:43-96    */
  SOCM_DOUBLE_WELL = 2,       /* b=-4k x(x^2-1), f=0, g=sum nu(x^2-1)^2   double_well.py:43-97 */
  SOCM_MOLECULAR_DYNAMICS = 3 /* double-well drift, f=1, g=0, Phi=-x_0    molecular_dynamics.py:49-95 */
} socm_kind;

/* Closed-form problem data (what the reference keeps as attributes of its NeuralSDE
 * subclass, method.py:15-56).  Unused pointers may be NULL. */
typedef struct {
  int32_t kind;              /* socm_kind */
  int32_t d;
  int32_t sigma_is_identity; /* 1: skip the d x d products with sigma (results identical) */
  float lmbd;
  const float* sigma;       /* [d][d] */
  const float* sigma_inv;   /* [d][d]  torch.inverse(sigma) */
  const float* A;           /* [d][d]  OU settings */
  const float* P;           /* [d][d]  OU_quadratic */
  const float* Q;           /* [d][d]  OU_quadratic */
  const float* omega;       /* [d]     OU_linear */
  const float* kappa;       /* [d]     double_well, molecular_dynamics */
  const float* nu;          /* [d]     double_well */
} socm_setting;

/* FullyConnectedUNet (models.py:202-242).  Layer order of the arrays:
 *   0 down_0 (d+1->h0)  1 down_1 (h0->h1)  2 down_2 (h1->h2)
 *   3 res_0  (d+1->d)   4 res_1  (h0->h0)  5 res_2  (h1->h1)
 *   6 up_2   (h2->h1)   7 up_1   (h1->h0)  8 up_0   (h0->d)          (= named_parameters order) */
typedef struct {
  int32_t d, h0, h1, h2;
  const float* w[9];
  const float* b[9];
} socm_unet;

/* Warm-start control as a per-grid-time affine table (models.py:163-199 evaluated on the
 * frozen Gaussian-path spline, gsbm_lib.py:227-306):
 *     u_ws(t_k, x) = sigma^{-1} ( c_k + A_k x - b(x) ). */
typedef struct {
  const float* A; /* [rows][d][d] */
  const float* c; /* [rows][d]    */
} socm_warm_table;

/* --- bookkeeping ----------------------------------------------------------------------- */
int socm_abi_version(void);
const char* socm_last_error(void);
/* SM count / opt-in shared memory of the current device (used to size persistent grids). */
int socm_device_info(int* sm_count, int* smem_optin_bytes);
/* Tensor-core engine of the default-width kernels where the call has no flag for it (the target GEMMs) and the default of
 * the others: -1 = the SOCM_F16 environment variable, else fp16 hi/lo split on kind::f16 wherever it applies;
 * 0 = 3xTF32; 1 = fp16 split.  Process-wide; the soc_matching_b200.simulate.ENGINE switch of the Python mirror sets it. */
int socm_set_default_engine(int32_t engine);

/* --- K1: Euler-Maruyama rollout  (replaces utils.stochastic_trajectories, utils.py:17-128,
 *     including NeuralSDE.control, method.py:58-80) ------------------------------------- */
#define SOCM_ROLLOUT_FORCE_GENERIC 1u /* use the shape-generic kernel even for the default net */
#define SOCM_ROLLOUT_NO_TRAJ 2u       /* weights-only mode: states/noises/controls may be NULL */
#define SOCM_ROLLOUT_FORCE_FFMA 4u    /* default net: use the fp32 FFMA tile kernel instead of tcgen05 (3xTF32) */
#define SOCM_ROLLOUT_F16 8u           /* default net, d <= 15: fp16-split tcgen05 engine, two CTAs per SM (unet_h.cuh) */
#define SOCM_ROLLOUT_TF32 16u         /* default net: the 3xTF32 tcgen05 engine even where the fp16-split one is the default */

/* step_tab: [5][K] = dt_k, sqrt(lmbd*dt_k), dt_k/lmbd, sqrt(dt_k/lmbd), t_k  computed by the
 *           caller in fp32 exactly like utils.py:38,47,95-98 (dt from the fp32 linspace).
 * noise_in: [K][B][d] injected N(0,1) draws, or NULL -> in-kernel Philox4x32-10 keyed by
 *           (seed, path_offset + path index, step) so results do not depend on the sharding.
 * outputs:  states [K+1][B][d], noises [K][B][d], controls [K][B][d], stop [K+1][B] (0/1 as
 *           fp32, utils.py:28,75), eff_dt [K][B] (utils.py:70-78), logw_det/sto/term [B].
 * warm:     rows = K (rank-2 branch time shift, models.py:170) or NULL.
 * workspace: socm_rollout_workspace_bytes() bytes (packed weight tape). */
int64_t socm_rollout_workspace_bytes(const socm_unet* net);
int socm_rollout_f32(const socm_setting* st, const socm_unet* net, const socm_warm_table* warm,
                     const float* x0, const float* step_tab, const float* noise_in,
                     uint64_t seed, uint64_t path_offset, int32_t B, int32_t K,
                     float* states, float* noises, float* controls, float* stop, float* eff_dt,
                     float* logw_det, float* logw_sto, float* logw_term,
                     void* workspace, uint32_t flags, void* stream);

/* The same rollout under a TABULATED control instead of the network: the `else` branch of NeuralSDE.control
 * (method.py:103-107) with the ground-truth controls of models.py:10-150, as used by control_objective /
 * normalization_constant on the optimal SDE (utils.py:131-231, main.py:117-153).
 *   SOCM_CONTROL_AFFINE  u_k(x) = A_k x + c_k, rows k = 0..K-1 tabulated by the caller on the grid times:
 *                        LinearControl (models.py:10-39)          A_k = u[floor((n-1) t_k / T)],  c = NULL
 *                        ConstantControlLinear (models.py:61-81)  A = NULL,  c_k = ut[floor(n t_k / T)]
 *   SOCM_CONTROL_LOOKUP  LowDimControl (models.py:84-150): u_j = ut[idx_t[k]][clamp(floor((x_j + xb) / dx), 0, nx-1)][j],
 *                        idx_t[k] = ceil(t_k / delta_t) tabulated by the caller, ut[nt][nx][d]. */
#define SOCM_CONTROL_AFFINE 0
#define SOCM_CONTROL_LOOKUP 1
typedef struct {
  int32_t kind;
  const float* A;       /* [K][d][d] or NULL */
  const float* c;       /* [K][d]    or NULL */
  const float* ut;      /* [nt][nx][d] */
  const int32_t* idx_t; /* [K] */
  int32_t nx;
  float xb, dx;
} socm_tab_control;
int socm_rollout_tabulated_f32(const socm_setting* st, const socm_tab_control* ctrl, const float* x0,
                               const float* step_tab, const float* noise_in, uint64_t seed, uint64_t path_offset,
                               int32_t B, int32_t K, float* states, float* noises, float* controls, float* stop,
                               float* eff_dt, float* logw_det, float* logw_sto, float* logw_term, uint32_t flags,
                               void* stream);

/* Philox4x32-10 + Box-Muller exactly as the rollout draws it: out[k][m][j]. */
int socm_philox_normal_f32(uint64_t seed, uint64_t path_offset, int32_t B, int32_t K, int32_t d,
                           float* out, void* stream);

/* UNet at n points: out[n][d] = nabla_V(tx[n][d+1])  (models.py:233-242). */
int socm_unet_forward_f32(const socm_unet* net, const float* tx, int32_t n, float* out, void* stream);

/* --- SOCM target  (replaces method.py:584-690 after the re-association of SURVEY.md A.3) - */
/* Right-hand side R[B][ldr] of the block-triangular contraction, per path m:
 *   cols (2j)d..(2j+1)d   a_jm = eff_dt*grad_f(x_j) - grad_b(x_j) c_jm
 *   cols (2j+1)d..(2j+2)d c_jm = sqrt(lmbd) sqrt(eff_dt) sigma^{-T} eps_jm + eff_dt sigma^{-T} u_jm
 *   cols 2Kd..(2K+1)d     grad_g(x_K),            j = 0..K-1,  ldr >= (2K+1)d
 * and the importance weight w[m] = exp(logw_det+logw_sto+logw_term) (method.py:258-262). */
int socm_target_prep_f32(const socm_setting* st, const float* states, const float* noises,
                         const float* controls, const float* eff_dt, const float* logw_det,
                         const float* logw_sto, const float* logw_term, int32_t B, int32_t K,
                         float* R, int32_t ldr, float* w, void* stream);

/* target[B][ldt] = R[B][ldr] * L^T,  L[(K+1)d][ldr] = rows (i,k): [M_i0 dM_i0 M_i1 dM_i1 ... M_iK]
 * (zero for j < i: only K-blocks with j >= i are read). */
int socm_target_gemm_f32(const float* L, const float* R, int32_t B, int32_t K, int32_t d,
                         int32_t ldr, float* target, int32_t ldt, void* stream);
/* The same contraction on the tcgen05 tensor cores, fp32 accumulation: fp16 hi/lo split on kind::f16 (csrc/target_h.cu,
 * default) or 3xTF32 (csrc/target_tc.cu), see socm_set_default_engine.
 * workspace: socm_target_gemm_tc_workspace_bytes(K, d) bytes (hi/lo split tape of L, rebuilt every call). */
int64_t socm_target_gemm_tc_workspace_bytes(int32_t K, int32_t d);
int socm_target_gemm_tc_f32(const float* L, const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                            float* target, int32_t ldt, void* workspace, void* stream);
/* dL[(K+1)d][ldr] (+)= G^T R  (contraction over paths), only the j >= i blocks are written. */
int socm_target_gemm_bwd_f32(const float* G, const float* R, int32_t B, int32_t K, int32_t d,
                             int32_t ldr, int32_t ldt, float* dL, int32_t accumulate, void* stream);
/* The same on the tcgen05 tensor cores: fp16 hi/lo planes on kind::f16 (csrc/target_bwd_h.cu, default) or 3xTF32
 * (csrc/target_bwd_tc.cu).  workspace: socm_target_gemm_bwd_tc_workspace_bytes(B, K, d) bytes (operand scratch).
 * `accumulate` is a bit field here: bit 0 = add into dL, SOCM_TARGET_BWD_TF32 / _F16 force an engine. */
#define SOCM_TARGET_BWD_TF32 2
#define SOCM_TARGET_BWD_F16 4
int64_t socm_target_gemm_bwd_tc_workspace_bytes(int32_t B, int32_t K, int32_t d);
int socm_target_gemm_bwd_tc_f32(const float* G, const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                                int32_t ldt, float* dL, int32_t accumulate, void* workspace, void* stream);
/* SOCM_const_M (method.py:289-369): target_i = sum_{j>=i} a_j + grad_g, i.e. M = I, dM = 0. */
int socm_target_const_m_f32(const float* R, int32_t B, int32_t K, int32_t d, int32_t ldr,
                            float* target, int32_t ldt, void* stream);
/* SOCM with stopping times (method.py:484-507, 524-564, 584-690): the per-sample table M(t, s, tau_m) depends on the
 * path only through its stopping index group[m] = #{k : Phi(x_km) > 0} - 1 (method.py:524-531), so the caller builds
 * one table per index, transposed: LT[n_groups][(2K+1)d][nrp] (columns of R x target rows, zero left of the block
 * diagonal), and   target[m][:] = R[m][:] . LT[group[m]].   perm[B] lists the paths sorted by group (the kernels
 * visit them in that order).  The reference's (K+1, K+1, B, d, d) intermediate never exists. */
int socm_target_grouped_f32(const float* LT, const float* R, const int32_t* group, const int32_t* perm,
                            int32_t n_groups, int32_t B, int32_t K, int32_t d, int32_t ldr, int32_t nrp,
                            float* target, int32_t ldt, void* stream);
/* dLT[g][c][r] += sum over the paths of group g of R[m][c] G[m][r]  (only the j >= i blocks are written). */
int socm_target_grouped_bwd_f32(const float* G, const float* R, const int32_t* group, const int32_t* perm,
                                int32_t n_groups, int32_t B, int32_t K, int32_t d, int32_t ldr, int32_t ldt,
                                int32_t nrp, float* dLT, void* stream);
/* SOCM_adjoint (method.py:722-749): target[m][i] = a_i, the adjoint state of path m from the backward recursion
 *   a_K = grad_g(x_K),  a_j = a_{j+1} + dt ((grad_f(x_j) + grad_f(x_{j+1})) / 2 + ((grad_b(x_j) + grad_b(x_{j+1})) / 2) a_{j+1})
 * with the constant dt = T / num_steps of method.py:169.  states: [K+1][B][d].  The loss that follows is the
 * same importance-weighted K3 as for SOCM (method.py:736-749 has the form of 692-720 with target = a). */
int socm_target_adjoint_f32(const socm_setting* st, const float* states, int32_t B, int32_t K, float dt,
                            float* target, int32_t ldt, void* stream);

/* --- optimiser step around the path -------------------------------------------
 * One launch for the Adam update of every parameter tensor (UNet, M-network, gamma, y0), replacing
 * torch.optim.Adam.step() [+ zero_grad()] of main.py:174-230, 350-352; torch's single-tensor Adam arithmetic
 * (no amsgrad, no weight decay), fp32.  `step` is the 1-based step count used for the bias corrections. */
typedef struct socm_adam_tensor {
  float* param;
  float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
  float lr;
} socm_adam_tensor;
int socm_adam_step_f32(const socm_adam_tensor* tensors /* host array */, int32_t n_tensors, double beta1, double beta2,
                       double eps, int32_t step, int32_t zero_grad, void* stream);

/* One launch for the per-iteration statistics of the reference's training loop (main.py:325-393, compute_EMA of
 * utils.py:389-396): squared norm of the control network's gradient, the EMA of every gradient tensor (updated in
 * place) and its squared norm, and the EMAs of loss / mean(w) / std(w) plus the running normalisation constant
 * (EMA of mean(w) with its own coefficient, main.py:354-359).
 *   scalars (device, fp32[3]): loss, mean(w), std(w) of this iteration
 *   stats   (device, fp32[8], in/out): 0 grad_norm_sqd  1 EMA_grad_norm_sqd  2 sqd_norm_EMA_grad  3 EMA_loss
 *                                      4 EMA_weight_mean  5 EMA_weight_std  6 normalization_const  7 unused
 *   scratch (device, 32 bytes, zeroed once by the caller; the kernel leaves it zeroed)
 *   itr = 0-based iteration index (selects the warm-up rule of compute_EMA). */
typedef struct socm_ema_tensor {
  const float* grad;
  float* ema_grad;
  int64_t n;
} socm_ema_tensor;
int socm_ema_stats_f32(const socm_ema_tensor* tensors /* host array */, int32_t n_tensors, const float* scalars,
                       float* stats, void* scratch, int32_t itr, double ema_coeff, double ema_weight_mean_coeff,
                       void* stream);

/* --- K3: UNet forward at all (K+1)B points + weighted loss + backward
 *     (replaces method.py:272-287, 692-720 and loss.backward(), main.py:323) -------------
 * loss_sums[0] (fp64) += sum_{i,m} s_im w_m |sigma^T (nabla_V(t_i,x_im) [- sigma^{-T} u_ws] - target_im)|^2 * scale
 * G[B][ldt]     = d loss / d target  (same scale);  grad[...] += d loss / d UNet parameters,
 * flat in the layer order of socm_unet (w then b per layer).
 * stop may be NULL (all ones).  warm: rows = K+1 (rank-3 branch, models.py:184-186) or NULL.
 * workspace: socm_loss_workspace_bytes() bytes.  flags: SOCM_LOSS_FORCE_GENERIC. */
#define SOCM_LOSS_FORCE_GENERIC 1u
#define SOCM_LOSS_FORCE_FFMA 2u /* default net: fp32 FFMA tile kernel instead of the tcgen05 kernels */
#define SOCM_LOSS_FORCE_TC 4u   /* default net: tcgen05 kernels even below SOCM_LOSS_TC_MIN_POINTS */
#define SOCM_LOSS_F16 8u        /* default net, d <= 15: fp16-split tcgen05 engine, two CTAs per SM (csrc/loss_h.cu) */
#define SOCM_LOSS_TF32 16u      /* default net: the 3xTF32 tcgen05 engine even where the fp16-split one is the default */
#define SOCM_LOSS_TC_MIN_POINTS 65536 /* (K+1)*B from which the tcgen05 kernels are the default */
int64_t socm_loss_workspace_bytes(const socm_unet* net, int32_t B, int32_t K);
int64_t socm_unet_param_count(const socm_unet* net);
int socm_unet_loss_fwdbwd_f32(const socm_setting* st, const socm_unet* net, const socm_warm_table* warm,
                              const float* ts, const float* states, const float* target, int32_t ldt,
                              const float* w, const float* stop, float scale, int32_t B, int32_t K,
                              float* G, float* grad, double* loss_sums, void* workspace,
                              uint32_t flags, void* stream);

/* sums[0] += sum w, sums[1] += sum w^2, sums[2] += sum stop  (method.py:715, 903-904);
 * accumulated in fp64 (warp-shuffle block reduction, one atomic per block). */
int socm_weight_stats_f32(const float* w, const float* stop, int32_t B, int32_t K, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SOCM_B200_H_ */
