"""Drop-in for the reference's ``SOC_Solver`` (method.py:146-906) restricted to the hot path:
``loss(..., algorithm in {"SOCM", "SOCM_const_M"})`` returns the same 8-tuple, and
``objective.backward()`` leaves ``.grad`` on ``neural_sde.nabla_V.parameters()``,
``neural_sde.M.sigmoid_layers.parameters()`` and ``neural_sde.gamma`` (``gamma2``).

One iteration is: K1 rollout -> R, w (prep) -> L from the M-network (torch, B-independent)
-> target = R L^T (K2) -> fused UNet forward + weighted loss + backward (K3, produces the
UNet gradients and G = d loss / d target) -> dL = G^T R (K2 backward) -> torch autograd through
L into the M-network.  Trajectories are processed in chunks of ``chunk_paths`` so that the
working set stays bounded for B up to 2^20 and beyond; every kernel accumulates.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, mtable, networks, simulate
from .sde import describe_setting


class _FusedObjective(torch.autograd.Function):
    """The kernels have already produced the loss value and all first-order gradients; this node
    only hands them to autograd (scaled by the incoming gradient, e.g. 1/normalization_const,
    main.py:320)."""

    @staticmethod
    def forward(ctx, value, lead, lead_grad, unet_grad_flat, *unet_params):
        ctx.save_for_backward(lead_grad if lead_grad is not None else value.new_zeros(()), unet_grad_flat)
        ctx.has_lead = lead is not None and lead_grad is not None
        ctx.shapes = [p.shape for p in unet_params]
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        lead_grad, flat = ctx.saved_tensors
        grads, off = [], 0
        for shp in ctx.shapes:
            n = math.prod(shp)
            grads.append((flat[off:off + n] * gout).reshape(shp))
            off += n
        return (None, lead_grad * gout if ctx.has_lead else None, None, None, *grads)


# losses that are functionals of per-path sums of a running term quadratic in the learned control (method.py:751-856)
PATH_FUNCTIONAL_LOSSES = ("cross_entropy", "log-variance", "variance", "moment")


class SOC_Solver(nn.Module):
    noise_type = "diagonal"
    sde_type = "ito"

    def __init__(self, neural_sde, x0, ut, T=1.0, num_steps=100, lmbd=1.0, d=2, sigma=None):
        super().__init__()
        self.dim = neural_sde.dim
        self.neural_sde = neural_sde
        self.x0, self.ut, self.T = x0, ut, T
        self.ts = torch.linspace(0, T, num_steps + 1).to(x0.device)   # method.py:167
        self.num_steps = num_steps
        self.dt = T / num_steps
        self.lmbd, self.d = lmbd, d
        self.y0 = nn.Parameter(torch.randn(1, device=x0.device))       # method.py:172 (unused by SOCM)
        self.sigma = neural_sde.sigma if sigma is None else sigma
        self.chunk_paths = None             # paths per kernel launch; None -> 4 full waves of 128-path tiles
        self.batch_reduce = None            # data-parallel shards: sums a small fp64 tensor over the ranks (dist.py)
        self.global_batch = None            # ... and the number of paths of all ranks together
        self.force_generic = False          # tests: run the shape-generic kernels
        self.force_ffma = False             # tests: fp32 FFMA tile kernels instead of the tcgen05 kernels
        self.force_tc = False               # tests: tcgen05 K3 even for small batches
        self.force_simt_target = False      # tests: fp32 SIMT target GEMMs (K2) next to the tcgen05 K1 / K3
        self._injected_noise = None         # tests: (K, B, d) noise replayed by the next loss() call
        self._pair_grid = None
        self.path_offset = 0                # first global path index of this rank (Philox counter)
        self.kernel_events = None           # bench: dict name -> [(start, end) CUDA events] when not None
        self.launch_count = 0               # kernels of libsocm_b200 launched by the last loss() call
        self.last_stats = None              # fp64 [sum w, sum w^2, sum stop] of the last loss() call

    # ------------------------------------------------------------------ helpers
    def inject_noise(self, noises: Optional[torch.Tensor]):
        """Parity hook: the next ``loss`` call uses these Brownian increments instead of Philox."""
        self._injected_noise = noises

    def control(self, t0, x0):
        """method.py:175-183."""
        x0 = x0.reshape(-1, self.dim)
        tx = torch.cat([t0.reshape(-1, 1).expand(x0.shape[0], 1), x0], dim=-1)
        return -torch.einsum("ij,bj->bi", self.sigma.t(), self.neural_sde.nabla_V(tx))

    def control_objective(self, batch_size, total_n_samples=65536):
        """method.py:185-221 (returns the trajectories of the first batch as the third item)."""
        mean, err = simulate.control_objective(self.neural_sde, self.x0, self.ts, self.lmbd, batch_size,
                                               total_n_samples)
        first = simulate.stochastic_trajectories(self.neural_sde, self.x0.repeat(batch_size, 1), self.ts,
                                                 self.lmbd)[0]
        return mean, err, first

    def _timed(self, name, n_launches, fn, *args):
        """Run one libsocm_b200 call; optionally bracket it with CUDA events on the current stream."""
        self.launch_count += n_launches
        if self.kernel_events is None:
            return _lib.check(fn(*args))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = fn(*args)
        b.record()
        self.kernel_events.setdefault(name, []).append((a, b))
        return _lib.check(rc)

    def _k3_launches(self, udesc, nb, K):
        """Kernels one socm_unet_loss_fwdbwd_f32 call launches (csrc/loss.cu, loss_tc.cu): the tcgen05 path
        packs the weight tapes once and runs K3a + K3b per sub-launch of the scratch (16384 / 8192 tiles); the FFMA tile path is
        pack + kernel + gradient reduction; the generic path one kernel."""
        default_arch = (udesc.h0, udesc.h1, udesc.h2) == (256, 128, 64)
        if self.force_generic or not default_arch:
            return 1
        tc = (not self.force_ffma) and udesc.d <= 23 and (self.force_tc or (K + 1) * nb >= 65536)
        if not tc:
            return 3
        n_tiles = (K + 1) * ((nb + 127) // 128)
        import ctypes
        sms = ctypes.c_int(0)
        _lib.check(_lib.load().socm_device_info(ctypes.byref(sms), None))
        n_sm = max(int(sms.value), 1)
        f16 = udesc.d <= 15 and simulate.ENGINE != "tf32" and os.environ.get("SOCM_F16", "1") != "0"
        if f16:   # csrc/loss_h.cu: fold + calibration + pack, (K3a + K3b) per sub-launch of 16384 // (2 SMs) * (2 SMs) tiles, fold_finish
            env = os.environ.get("SOCM_SUB_TILES")
            cap = int(env) if env is not None and env.isdigit() and int(env) >= 1024 else 16384   # loss_h.cu sub_tiles_cap()
            sub = (cap // (2 * n_sm)) * (2 * n_sm)
            return 4 + 2 * ((n_tiles + sub - 1) // sub)
        # csrc/loss_tc.cu: fold + pack, (K3a + K3b) per sub-launch, fold_finish
        sub = (8192 // n_sm) * n_sm
        return 3 + 2 * ((n_tiles + sub - 1) // sub)

    def _grid(self):
        if self._pair_grid is None or self._pair_grid.t.device != self.ts.device:
            self._pair_grid = mtable.make_pair_grid(self.ts, self.T)
        return self._pair_grid

    # ------------------------------------------------------------------ the hot path
    def loss(self, batch_size, compute_L2_error=False, optimal_control=None, compute_control_objective=False,
             algorithm="SOCM_const_M", add_weights=False, total_n_samples=65536, verbose=False,
             u_warm_start=None, use_warm_start=True, use_stopping_time=False):
        if algorithm in PATH_FUNCTIONAL_LOSSES:
            return self._path_functional_loss(batch_size, algorithm, add_weights, bool(use_stopping_time), u_warm_start,
                                              use_warm_start, compute_L2_error, optimal_control,
                                              compute_control_objective, total_n_samples)
        if algorithm not in ("SOCM", "SOCM_const_M", "SOCM_exp", "SOCM_adjoint"):
            raise NotImplementedError(
                f"algorithm {algorithm!r}: SOCM, SOCM_const_M, SOCM_exp, SOCM_adjoint and {PATH_FUNCTIONAL_LOSSES} run on "
                "the B200 path; rel_entropy (back-propagation through the rollout) does not (SURVEY.md section 8f)")
        if compute_L2_error and optimal_control is None:
            raise ValueError("compute_L2_error=True needs optimal_control (a callable (ts, states, t_is_tensor=True))")
        lib = _lib.load()
        simulate.sync_engine(lib)
        sde = self.neural_sde
        dev = self.x0.device
        _lib.require_cuda(self.x0, "x0")
        desc = describe_setting(sde, dev)
        desc.c_struct.lmbd = float(self.lmbd)
        d, K, B = desc.d, self.num_steps, int(batch_size)
        ts = self.ts.to(dev)
        # Only the SOCM branch of the reference looks at use_stopping_time (method.py:484-720): SOCM_const_M
        # (289-369), SOCM_exp (371-478) and SOCM_adjoint (722-749) never read the flag -- no mask, plain dts,
        # normaliser (K+1) B -- even when the rollout itself stopped paths (utils.py:33 keys on Phi alone).
        stopping = bool(use_stopping_time) and algorithm == "SOCM"
        if stopping and not desc.has_stopping:
            raise _lib.SocmError("use_stopping_time=True needs a setting with a stopping function Phi")
        if B <= 0:
            raise ValueError(f"batch_size must be positive, got {B}")
        # warm start: the reference only applies it in the loss if u_warm_start is passed AND
        # use_warm_start (method.py:280); the rollout uses the sde's own flags (method.py:77).
        warm_loss = None
        if u_warm_start and use_warm_start:
            warm_loss = simulate.resolve_warm_start(_WarmView(u_warm_start), ts)

        unet = sde.nabla_V
        udesc, keep = networks.unet_desc(unet)
        uparams = networks.unet_parameters(unet)
        n_par = int(lib.socm_unet_param_count(udesc))
        f32 = dict(device=dev, dtype=torch.float32)
        grad_flat = torch.zeros(n_par, **f32)
        loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
        stats = torch.zeros(3, device=dev, dtype=torch.float64)
        ldt = ((K + 1) * d + 3) // 4 * 4
        ldr = ((2 * K + 1) * d + 3) // 4 * 4
        nrows = (K + 1) * d
        stream = _lib.stream_ptr()

        # ---- B-independent part: the M table (SOCM without stopping times)
        L = dL = None
        m_all = dm_all = None
        if algorithm == "SOCM" and not stopping:
            grid = self._grid()
            m_all, dm_all = sde.M.value_and_ds(grid.t, grid.s)
            L = mtable.build_L(m_all, dm_all, grid, ldr)
            dL = torch.zeros(nrows, ldr, **f32)
        elif algorithm == "SOCM_exp":
            # method.py:371-478 is SOCM with M_t(s) = exp(-gamma (s - t)) I and the decay rate a Parameter of the
            # SOLVER (main.py:166-169): the same table, GEMMs and backward, with an analytic M and d/ds M
            gamma_p = getattr(self, "gamma", None)
            if not torch.is_tensor(gamma_p):
                raise ValueError("SOCM_exp: set solver.gamma = torch.nn.Parameter(torch.tensor([gamma])) (main.py:166-169)")
            grid = self._grid()
            decay = torch.exp(-gamma_p.to(dev) * (grid.s - grid.t)).reshape(-1, 1, 1)
            eye = torch.eye(d, **f32)
            L = mtable.build_L(decay * eye, (-gamma_p.to(dev)).reshape(1, 1, 1) * decay * eye, grid, ldr)
            dL = torch.zeros(nrows, ldr, **f32)
        LT = dLT = None
        if stopping:
            # method.py:484-507, 524-564: M(t, s, tau_m) depends on the path only through its stopping index
            # q_m = #{k : Phi(x_km) > 0} - 1 (tau_m = q_m / K), so there are K+1 tables, built once per iteration
            # from the M-network (B-independent, differentiable) -- csrc/target_grouped.cu contracts them per path
            n_groups = K + 1
            table_bytes = 4 * n_groups * (2 * K + 1) * d * ldt
            if table_bytes > (8 << 30):
                raise NotImplementedError(
                    f"stopping-time SOCM keeps one (2K+1)d x (K+1)d table per stopping index: {table_bytes / 2**30:.1f} "
                    "GiB at this (K, d); the reference's setting is d = 1 (molecular_dynamics.py)")
            grid = self._grid()
            tau_vals = torch.arange(n_groups, **f32) / K                          # (cnt - 1) / K, method.py:524-530
            m_all, dm_all = sde.M.value_and_ds(grid.t, grid.s, tau_vals.unsqueeze(0).expand(grid.P, n_groups))
            LT = mtable.build_LT_grouped(m_all, dm_all, grid, ldt)
            dLT = torch.zeros_like(LT)

        if self.chunk_paths is None:
            # whole waves of persistent CTAs (one 128-path tile per SM and wave): no ragged last wave
            import ctypes
            sms = ctypes.c_int(0)
            _lib.check(lib.socm_device_info(ctypes.byref(sms), None))
            self.chunk_paths = 4 * 128 * max(int(sms.value), 1)
        chunk = min(B, int(self.chunk_paths))
        wsp = simulate.RolloutWorkspace(desc, unet, chunk, K, dev, True)
        R = torch.empty(chunk, ldr, **f32)
        wbuf = torch.empty(chunk, **f32)
        target = torch.empty(chunk, ldt, **f32)
        G = torch.zeros(chunk, ldt, **f32)
        lws = int(lib.socm_loss_workspace_bytes(udesc, chunk, K))
        loss_ws = torch.empty((lws + 3) // 4, **f32)
        stop_all = []
        seed = simulate.next_seed()
        scale = 1.0 if stopping else 1.0 / ((K + 1) * B)                    # method.py:715 / 720
        warm_struct = simulate._warm_struct(warm_loss.A_loss, warm_loss.c_loss) if warm_loss is not None else None
        k2_ws = None          # workspace of the tcgen05 target GEMM
        # the tcgen05 target GEMM gives one CTA 128 paths and streams the whole table through it: below ~1k paths
        # most SMs would idle (0.72 ms at B = 128 vs 0.03 ms for the SIMT GEMM, whose grid tiles rows x paths)
        simt_target = self.force_ffma or self.force_generic or self.force_simt_target or B < 1024
        k2b_ws, k2b_nb = None, -1

        x0_rep = self.x0.detach().float().reshape(1, d)
        ts_f32 = ts.float().contiguous()
        # method.py:648-673: the fractional time steps of the rollout enter the target only under use_stopping_time;
        # otherwise the plain dts do (they coincide unless the setting has a stopping function)
        plain_dt = None
        if desc.has_stopping and not stopping:
            plain_dt = (ts_f32[1:] - ts_f32[:-1]).reshape(K, 1).expand(K, chunk).contiguous()
        self.launch_count = 0
        l2_sum = torch.zeros((), device=dev, dtype=torch.float64) if compute_L2_error else None
        for start in range(0, B, chunk):
            nb = min(chunk, B - start)
            if nb != wsp.B:   # ragged last chunk
                wsp = simulate.RolloutWorkspace(desc, unet, nb, K, dev, True)
                R, wbuf = R[:nb], wbuf[:nb]
                target, G = target[:nb], G[:nb]
            noises = None
            if self._injected_noise is not None:
                noises = self._injected_noise[:, start:start + nb].contiguous()
            simulate.rollout(sde, x0_rep.expand(nb, d).contiguous(), ts, self.lmbd, noises=noises, seed=seed,
                             path_offset=start + self.path_offset, desc=desc, workspace=wsp,
                             force_generic=self.force_generic, force_ffma=self.force_ffma, timer=self._timed)
            self._timed("prep", 2, lib.socm_target_prep_f32,
                        desc.c_struct, _lib.ptr(wsp.states), _lib.ptr(wsp.noises), _lib.ptr(wsp.controls),
                        _lib.ptr(wsp.eff_dt if plain_dt is None else plain_dt[:, :nb].contiguous()),
                        wsp.lw[0].data_ptr(), wsp.lw[1].data_ptr(), wsp.lw[2].data_ptr(), nb, K,
                        _lib.ptr(R), ldr, _lib.ptr(wbuf), stream)
            self._timed("stats", 1, lib.socm_weight_stats_f32, _lib.ptr(wbuf),
                        _lib.ptr(wsp.stop) if stopping else None, nb, K, _lib.ptr(stats), stream)
            if algorithm == "SOCM_const_M":
                self._timed("target", 1, lib.socm_target_const_m_f32, _lib.ptr(R), nb, K, d, ldr, _lib.ptr(target),
                            ldt, stream)
            elif algorithm == "SOCM_adjoint":                  # method.py:722-735: adjoint recursion per path
                self._timed("target", 1, lib.socm_target_adjoint_f32, desc.c_struct, _lib.ptr(wsp.states), nb, K,
                            float(self.dt), _lib.ptr(target), ldt, stream)
            elif not stopping:
                if k2_ws is None and not simt_target:
                    nbytes = int(lib.socm_target_gemm_tc_workspace_bytes(K, d))
                    if nbytes < 0:        # (K+1) d beyond the tcgen05 kernel's plan tables: fp32 SIMT GEMMs instead
                        simt_target = True
                    else:
                        k2_ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                if simt_target:                                # fp32 SIMT GEMM
                    self._timed("target", 1, lib.socm_target_gemm_f32, _lib.ptr(L.detach()), _lib.ptr(R), nb, K, d,
                                ldr, _lib.ptr(target), ldt, stream)
                else:   # tcgen05: fp16 split (2 x absmax + tape pack + GEMM) or 3xTF32 (tape pack + GEMM), simulate.sync_engine
                    k2_f16 = simulate.ENGINE != "tf32" and (simulate.ENGINE == "f16" or os.environ.get("SOCM_F16") != "0")
                    self._timed("target", 4 if k2_f16 else 2, lib.socm_target_gemm_tc_f32, _lib.ptr(L.detach()), _lib.ptr(R),
                                nb, K, d, ldr, _lib.ptr(target), ldt, k2_ws.data_ptr(), stream)
            else:
                # stopping index of every path (Phi(x) = -x_0 > 0, method.py:524-530) and the order that groups them
                q_idx = ((wsp.states[..., 0] < 0).sum(dim=0) - 1).to(torch.int32).contiguous()
                perm = torch.sort(q_idx, stable=True)[1].to(torch.int32).contiguous()
                self._timed("target", 1, lib.socm_target_grouped_f32, _lib.ptr(LT.detach()), _lib.ptr(R),
                            q_idx.data_ptr(), perm.data_ptr(), K + 1, nb, K, d, ldr, ldt, _lib.ptr(target), ldt, stream)
            self._timed("loss_fwdbwd", self._k3_launches(udesc, nb, K), lib.socm_unet_loss_fwdbwd_f32,
                        desc.c_struct, udesc, warm_struct, _lib.ptr(ts_f32), _lib.ptr(wsp.states),
                        _lib.ptr(target), ldt, _lib.ptr(wbuf), _lib.ptr(wsp.stop) if stopping else None, scale, nb, K,
                        _lib.ptr(G), _lib.ptr(grad_flat), _lib.ptr(loss_sum), _lib.ptr(loss_ws),
                        (_lib.LOSS_FORCE_GENERIC if self.force_generic else 0)
                        | (_lib.LOSS_FORCE_FFMA if self.force_ffma else 0)
                        | (_lib.LOSS_FORCE_TC if self.force_tc else 0)
                        | {None: 0, "f16": _lib.LOSS_F16, "tf32": _lib.LOSS_TF32}[simulate.ENGINE], stream)
            if L is not None:
                if simt_target:                                # fp32 SIMT GEMM
                    self._timed("target_bwd", 1, lib.socm_target_gemm_bwd_f32, _lib.ptr(G), _lib.ptr(R), nb, K, d, ldr,
                                ldt, _lib.ptr(dL), 1, stream)
                else:                                          # tcgen05: fp16 planes on kind::f16, or 3xTF32
                    if k2b_ws is None or k2b_nb != nb:
                        k2b_ws = torch.empty(int(lib.socm_target_gemm_bwd_tc_workspace_bytes(nb, K, d)), device=dev,
                                             dtype=torch.uint8)
                        k2b_nb = nb
                    k2b_f16 = simulate.ENGINE != "tf32" and (simulate.ENGINE == "f16" or os.environ.get("SOCM_F16") != "0")
                    self._timed("target_bwd", 4 if k2b_f16 else 3,   # absmax + 2 packs + GEMM | 2 transposes + GEMM
                                lib.socm_target_gemm_bwd_tc_f32, _lib.ptr(G), _lib.ptr(R), nb, K, d, ldr, ldt, _lib.ptr(dL),
                                1 | (_lib.TARGET_BWD_F16 if k2b_f16 else _lib.TARGET_BWD_TF32), k2b_ws.data_ptr(), stream)
            if stopping:
                self._timed("target_bwd", 1, lib.socm_target_grouped_bwd_f32, _lib.ptr(G), _lib.ptr(R), q_idx.data_ptr(),
                            perm.data_ptr(), K + 1, nb, K, d, ldr, ldt, ldt, _lib.ptr(dLT), stream)
            if compute_L2_error:
                l2_sum += self._l2_error_sum(sde, unet, optimal_control, warm_loss, ts_f32, wsp.states, wbuf)
            if getattr(self, "_debug_keep", False):   # scripts/k3_truth.py: inputs of the last K3 call
                self._debug_last = dict(states=wsp.states.clone(), target=target.clone(), w=wbuf.clone(), ldt=ldt)
            stop_all.append(wsp.stop if B <= chunk else wsp.stop.clone())
        self._injected_noise = None
        del keep

        value = loss_sum[0]
        lead, lead_grad = L, dL
        if stopping:
            z = stats[2]                                                   # sum(stop_indicators), method.py:715
            value = value / z
            grad_flat = (grad_flat.double() / z).float()
            lead, lead_grad = LT, (dLT.double() / z).float()
        objective = _FusedObjective.apply(value.float(), lead, lead_grad, grad_flat, *uparams)

        self.last_stats = stats
        mean_w = (stats[0] / B).float()
        var_w = (stats[1] - stats[0] * stats[0] / B) / max(B - 1, 1)       # unbiased, method.py:904
        std_w = torch.sqrt(torch.clamp(var_w, min=0.0)).float()
        stop_indicators = stop_all[0] if len(stop_all) == 1 else torch.cat(stop_all, dim=1)

        ctrl_mean = ctrl_err = trajectory = None
        if compute_control_objective:
            ctrl_mean, ctrl_err, trajectory = self.control_objective(batch_size, total_n_samples=total_n_samples)
        norm_sqd_diff = (l2_sum / ((K + 1) * B)).float() if compute_L2_error else None
        return (objective, norm_sqd_diff, ctrl_mean, ctrl_err, trajectory, mean_w, std_w, stop_indicators)

    # ------------------------------------------------------------------ cross_entropy / (log-)variance / moment
    def _path_functional_loss(self, batch_size, algorithm, add_weights, stopping, u_warm_start, use_warm_start,
                              compute_L2_error, optimal_control, compute_control_objective, total_n_samples):
        """method.py:751-856.  Every one of these objectives is F(S_1..S_B) with the per-path sum
            S_m = sum_k [ delta_km (-(u.c)/lmbd + |u|^2/(2 lmbd) [- f/lmbd]) - sqrt(delta_km/lmbd) u.eps ]  [- g(x_K)/lmbd],
        u = u_theta(t_k, x_km) = -sigma^T nabla_V, c / eps = the rollout's controls / noises, delta = dt_k (or the
        fractional step x stop indicator).  Completing the square,
            S_m = sum_k (delta/(2 lmbd)) |u - (c + sqrt(lmbd/delta) eps)|^2 + (terms without theta),
        so dF/dtheta is the gradient of the *importance-weighted squared error K3 already computes*, with path
        weights a_m = dF/dS_m, point weights delta/(2 lmbd) and target -sigma^{-T}(c + sqrt(lmbd/delta) eps).
        The value F and a_m come from one extra UNet forward (socm_unet_forward_f32) and reductions over the
        rollout outputs (torch ops on the GPU); the UNet gradient from the same fused K3 kernels as SOCM."""
        lib = _lib.load()
        simulate.sync_engine(lib)
        sde, dev = self.neural_sde, self.x0.device
        _lib.require_cuda(self.x0, "x0")
        desc = describe_setting(sde, dev)
        desc.c_struct.lmbd = float(self.lmbd)
        d, K, B = desc.d, self.num_steps, int(batch_size)
        if stopping and not desc.has_stopping:
            raise _lib.SocmError("use_stopping_time=True needs a setting with a stopping function Phi")
        if B <= 0:
            raise ValueError(f"batch_size must be positive, got {B}")
        # Memory bound: these objectives are functionals of ALL per-path sums S_m (a variance / second moment over
        # the batch), so the value side keeps about ten (K+1) x B x d fp32 tensors alive at once; unlike the SOCM
        # path it is not chunked.  Refuse batches beyond one chunk instead of running the box out of memory.
        limit = int(self.chunk_paths) if self.chunk_paths else 4 * 128 * 148
        if B > limit:
            raise NotImplementedError(
                f"{algorithm!r} keeps ~10 (K+1) x B x d tensors resident and is not chunked: batch_size {B} exceeds "
                f"chunk_paths = {limit}")
        ts = self.ts.to(dev).float().contiguous()
        lam = float(self.lmbd)
        unet = sde.nabla_V
        udesc, keep = networks.unet_desc(unet)
        uparams = networks.unet_parameters(unet)
        f32 = dict(device=dev, dtype=torch.float32)
        warm_loss = simulate.resolve_warm_start(_WarmView(u_warm_start), ts) if (u_warm_start and use_warm_start) else None
        self.launch_count = 0
        noises_in, self._injected_noise = self._injected_noise, None
        wsp = simulate.rollout(sde, self.x0.detach().float().reshape(1, d).expand(B, d).contiguous(), ts, self.lmbd,
                               noises=noises_in, seed=simulate.next_seed(), path_offset=self.path_offset, desc=desc,
                               force_generic=self.force_generic, force_ffma=self.force_ffma, timer=self._timed)
        states, noises, controls = wsp.states, wsp.noises, wsp.controls
        w = torch.exp(wsp.lw[0] + wsp.lw[1] + wsp.lw[2])                       # method.py:258-262
        sigma = self.sigma.to(dev).float()
        sig_inv = torch.inverse(sigma)
        # ---- value: one UNet forward at all (K+1) B points, then reductions (method.py:752-856)
        gv = unet(torch.cat([ts.reshape(-1, 1, 1).expand(K + 1, B, 1), states], dim=-1))
        if warm_loss is not None:                                              # method.py:280-287
            aff = warm_loss.c_loss.unsqueeze(1) + torch.einsum("kij,kbj->kbi", warm_loss.A_loss, states)
            gv = gv - torch.einsum("ij,kbj->kbi", sig_inv.t(), torch.einsum("ij,kbj->kbi", sig_inv, aff - sde.b(ts, states)))
        u = -torch.einsum("ij,abj->abi", sigma.t(), gv)[:-1]
        variance_family = algorithm != "cross_entropy"
        if stopping:
            delta = wsp.eff_dt * wsp.stop[:-1] if variance_family else wsp.eff_dt
        else:
            delta = (ts[1:] - ts[:-1]).unsqueeze(1).expand(K, B)
        det = -(1 / lam) * torch.sum(u * controls, dim=2) + (1 / (2 * lam)) * torch.sum(u**2, dim=2)
        if variance_family:
            det = det - (1 / lam) * sde.f(ts[0], states)[:-1]
        sto = -math.sqrt(1 / lam) * torch.sum(u * noises, dim=2)
        S = torch.sum(det * delta, dim=0) + torch.sum(sto * torch.sqrt(delta), dim=0)
        if variance_family:
            S = S - (1 / lam) * sde.g(states[-1])
        S_leaf = S.detach().requires_grad_(True)
        w2 = w if add_weights else torch.ones_like(w)
        if algorithm == "cross_entropy":
            obj = torch.mean(S_leaf * w)
        elif algorithm == "moment":
            obj = torch.mean((S_leaf + self.y0) ** 2 * w2)
        else:
            sums = S_leaf if algorithm == "log-variance" else torch.exp(S_leaf)
            if self.batch_reduce is None:
                obj = B / (B - 1) * (torch.mean(sums**2 * w2) - torch.mean(sums * w2) ** 2)
            else:
                # data-parallel shard (dist.sharded_loss_backward): the functional is a variance over ALL paths, so the
                # first and second moments are summed over the ranks; autograd sees this shard's terms, the other
                # ranks' enter as constants, and the ranks' gradients of the same global value add up to its gradient
                Bg = int(self.global_batch)
                m1, m2 = torch.sum(sums * w2), torch.sum(sums**2 * w2)
                tot = self.batch_reduce(torch.stack([m1.detach(), m2.detach()]).double())
                m1g = m1 + (tot[0] - m1.detach().double()).to(m1.dtype)
                m2g = m2 + (tot[1] - m2.detach().double()).to(m2.dtype)
                obj = Bg / (Bg - 1) * (m2g / Bg - (m1g / Bg) ** 2)
        a_m, = torch.autograd.grad(obj, S_leaf, retain_graph=algorithm == "moment")
        # ---- gradient: K3 with path weights a_m, point weights delta / (2 lmbd), target -sigma^{-T}(c + sqrt(lmbd/delta) eps)
        ldt = ((K + 1) * d + 3) // 4 * 4
        pos = delta > 0
        tgt = controls + torch.where(pos, torch.sqrt(lam / torch.where(pos, delta, torch.ones_like(delta))),
                                     torch.zeros_like(delta)).unsqueeze(2) * noises
        tgt = -torch.einsum("ij,abj->abi", sig_inv.t(), tgt)
        target = torch.zeros(B, ldt, **f32)
        target[:, :K * d] = tgt.permute(1, 0, 2).reshape(B, K * d)
        point_w = torch.zeros(K + 1, B, **f32)
        point_w[:-1] = delta / (2 * lam)
        G = torch.zeros(B, ldt, **f32)
        grad_flat = torch.zeros(int(lib.socm_unet_param_count(udesc)), **f32)
        loss_sum = torch.zeros(1, device=dev, dtype=torch.float64)
        loss_ws = torch.empty((int(lib.socm_loss_workspace_bytes(udesc, B, K)) + 3) // 4, **f32)
        warm_struct = simulate._warm_struct(warm_loss.A_loss, warm_loss.c_loss) if warm_loss is not None else None
        self._timed("loss_fwdbwd", self._k3_launches(udesc, B, K), lib.socm_unet_loss_fwdbwd_f32,
                    desc.c_struct, udesc, warm_struct, _lib.ptr(ts), _lib.ptr(states), _lib.ptr(target), ldt,
                    _lib.ptr(a_m.float().contiguous()), _lib.ptr(point_w), 1.0, B, K, _lib.ptr(G), _lib.ptr(grad_flat),
                    _lib.ptr(loss_sum), _lib.ptr(loss_ws),
                    (_lib.LOSS_FORCE_GENERIC if self.force_generic else 0)
                    | (_lib.LOSS_FORCE_FFMA if self.force_ffma else 0)
                    | (_lib.LOSS_FORCE_TC if self.force_tc else 0)
                        | {None: 0, "f16": _lib.LOSS_F16, "tf32": _lib.LOSS_TF32}[simulate.ENGINE], _lib.stream_ptr())
        del keep
        objective = _FusedObjective.apply(obj.detach().float(), None, None, grad_flat, *uparams)
        if algorithm == "moment":                      # d/d y0 through the torch-side functional
            objective = objective + (obj - obj.detach())
        norm_sqd_diff = None
        if compute_L2_error:
            if optimal_control is None:
                raise ValueError("compute_L2_error=True needs optimal_control")
            norm_sqd_diff = (self._l2_error_sum(sde, unet, optimal_control, warm_loss, ts, states, w) / ((K + 1) * B)).float()
        ctrl_mean = ctrl_err = trajectory = None
        if compute_control_objective:
            ctrl_mean, ctrl_err, trajectory = self.control_objective(batch_size, total_n_samples=total_n_samples)
        self.last_stats = torch.stack([w.double().sum(), (w.double() ** 2).sum(), wsp.stop.double().sum()])
        return (objective, norm_sqd_diff, ctrl_mean, ctrl_err, trajectory, torch.mean(w), torch.std(w), wsp.stop)

    # ------------------------------------------------------------------ evaluation metric (not on the training path)
    def _l2_error_sum(self, sde, unet, optimal_control, warm_loss, ts, states, w):
        """sum_{i,m} w_m |u*(t_i, x_im) - u_theta(t_i, x_im)|^2 (method.py:858-875; the caller divides by
        (K+1) B).  u_theta = -sigma^T nabla_V [+ u_ws]: the UNet runs through socm_unet_forward_f32, the
        ground-truth control is the caller's torch callable, the reduction is a torch op."""
        K1, nb, d = states.shape
        tx = torch.cat([ts.reshape(-1, 1, 1).expand(K1, nb, 1), states], dim=-1)
        learned = -torch.einsum("ij,abj->abi", self.sigma.t().float(), unet(tx))
        if warm_loss is not None:   # method.py:280-287: nabla_V - sigma^{-T} u_ws  =>  u_theta + u_ws
            aff = warm_loss.c_loss.unsqueeze(1) + torch.einsum("kij,kbj->kbi", warm_loss.A_loss, states)
            learned = learned + torch.einsum("ij,kbj->kbi", torch.inverse(self.sigma.float()), aff - sde.b(ts, states))
        with torch.no_grad():
            target_control = optimal_control(ts, states, t_is_tensor=True).detach()
        return torch.sum(((target_control - learned) ** 2).double() * w.double().reshape(1, -1, 1))


class _WarmView:
    """Lets ``resolve_warm_start`` treat an explicitly passed u_warm_start like an sde attribute."""

    def __init__(self, ws):
        self.u_warm_start, self.use_warm_start = ws, True
