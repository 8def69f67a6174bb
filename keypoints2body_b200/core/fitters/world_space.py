"""World-space fitter backed by the fused CUDA kernel.

Drop-in for the reference's ``WorldSpaceFitter``
(/root/reference/keypoints2body/core/fitters/world_space.py:53-323): same
constructor arguments, same ``fit_frame`` signature / result contract.  What
differs is the execution: the body model's weights are uploaded once
(``NativeModel``), and a whole batch of frames is fitted by ONE persistent kernel
launch (``k2b_fit_batch``) that runs forward, analytic backward, all priors and
the Adam / L-BFGS update for every iteration on chip, followed by one mesh pass
(``k2b_mesh_batch``) for the returned vertices / joints.

``fit_batch`` is the batched entry the sequence driver and the benchmark use;
``fit_frame`` is the reference-shaped call.  It accepts B > 1 like the reference
does (with the reference's quirk that a 2-D confidence collapses to its first
row, world_space.py:163-164); for Adam that is identical to the reference (Adam is
batch-separable), for L-BFGS the B frames are B INDEPENDENT fits here, whereas the
reference runs one joint L-BFGS over the summed loss (one line search for all
frames) -- the API path is B = 1, where the two coincide.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from ... import _native as nat
from ...body_model import BodyModelWeights, extract_weights, full_pose_from_params
from ...models.smpl_data import BodyModelFitResult, SMPLData, SMPLHData, SMPLXData
from ..constants import AMASS_JOINT_MAP, JOINT_MAP
from ..prior import GMMConstants, load_gmm, prepare_gmm

_EXTRA_BLOCKS = ("left_hand_pose", "right_hand_pose", "expression", "jaw_pose", "leye_pose", "reye_pose")


def _f32(t, device):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)       # numpy prev_params are accepted (the reference crashes on them)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _check_widths(B, **blocks):
    """The kernels read fixed strides (72 / 69 / 10 / 3 floats per row): anything else would be read misaligned and
    out of bounds, so widths are validated before raw pointers cross the C ABI."""
    for name, (t, width) in blocks.items():
        if t is None:
            continue
        if t.dim() != 2 or t.shape[1] != width or t.shape[0] not in (1, B):
            raise ValueError(f"{name} must be (B={B}, {width}), got {tuple(t.shape)}"
                             + (" -- body_pose is the reference's 69-D block for every SMPL-family model "
                                "(SURVEY.md section 8c)" if name == "body_pose" else ""))


def guess_init_transl_from_root(fitter_or_model, pose_aa, betas, j3d_world_frame, joints_category="SMPL24"):
    """Root-joint alignment (world_space.py:13-50): ``target_root - model_root``.

    The model root at the given pose / shape comes from the mesh kernel's skeleton pass.
    """
    if joints_category not in ("SMPL24", "AMASS"):
        raise ValueError(f"Unknown joints category: {joints_category}")
    root = (JOINT_MAP if joints_category == "SMPL24" else AMASS_JOINT_MAP)["MidHip"]
    fitter = fitter_or_model
    pose_aa = _f32(pose_aa, fitter.device)
    out = fitter.forward_batch(
        {"global_orient": pose_aa[:, :3], "body_pose": pose_aa[:, 3:], "betas": _f32(betas, fitter.device)},
        with_vertices=False,
    )
    return (_f32(j3d_world_frame, fitter.device)[:, root, :] - out["joints"][:, root, :]).detach()


class _OverlapTuner:
    """Chooses ``mesh_capped_fraction`` of ``fit_chain`` by measurement.  While windows are being fitted the mesh pass
    of the finished ones can either be held to the SMs the fit leaves free (slower mesh, undisturbed fit) or run
    everywhere (its long-lived CTAs delay the next window's fit); which share of the windows to hold back depends on
    how long a window's fit takes relative to its mesh pass (optimiser, iteration budgets, model).  Each candidate share
    is tried on whole calls (after a first, cold call that is not measured), timed by two events on the caller's stream;
    the timing is read at the start of the next call (by then the work has long finished), so the policy adds no
    synchronisation to the call it measures."""

    CANDIDATES = (0.55, 0.65, 0.75, 0.85)
    REPEATS = 2           # each candidate is timed twice, the faster run counts (one disturbed call must not decide)

    def __init__(self):
        self.times = {}
        self.counts = {}
        self.best = None
        self.pending = None
        self.calls = 0          # the first call of a problem shape is cold (allocations, first launches): not measured

    def _next_candidate(self):
        for r in range(1, self.REPEATS + 1):
            for c in self.CANDIDATES:
                if self.counts.get(c, 0) < r:
                    return c
        return None

    def collect(self):
        if self.pending is not None:
            frac, e0, e1 = self.pending
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            self.times[frac] = min(ms, self.times.get(frac, ms))
            self.counts[frac] = self.counts.get(frac, 0) + 1
            self.pending = None
            if self.best is None and self._next_candidate() is None:
                self.best = min(self.CANDIDATES, key=lambda c: self.times[c])

    def next_fraction(self):
        self.collect()
        self.calls += 1
        if self.best is not None:
            return self.best
        if self.calls == 1:
            return self.CANDIDATES[-2]
        return self._next_candidate()

    def begin(self, stream):
        self._frac_e0 = torch.cuda.Event(enable_timing=True)
        self._frac_e0.record(stream)

    def end(self, stream):
        if self.best is None and self.calls > 1:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record(stream)
            self.pending = (self._next_candidate(), self._frac_e0, e1)


class WorldSpaceFitter:
    """Per-frame optimiser in world coordinates, B frames per kernel launch."""

    def __init__(
        self,
        smpl_model,
        step_size=1e-2,
        num_iters_first=30,
        num_iters_followup=10,
        use_lbfgs=True,
        joints_category="SMPL24",
        device=None,
        pose_prior_num_gaussians=8,
        *,
        model_type: Optional[str] = None,
        prior_folder: str = "./data/models/",
        gmm: Optional[dict] = None,
    ):
        if not torch.cuda.is_available():
            raise RuntimeError("keypoints2body_b200 needs a CUDA device (there is no CPU fallback)")
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if device.type != "cuda":
            raise ValueError(f"keypoints2body_b200 runs on CUDA devices only, got device={device}")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if joints_category not in ("SMPL24", "AMASS", "GENERIC"):
            raise ValueError("No such joints category!")
        self.smpl = smpl_model
        self.step_size = float(step_size)
        self.num_iters_first = int(num_iters_first)
        self.num_iters_followup = int(num_iters_followup)
        self.use_lbfgs = bool(use_lbfgs)
        self.device = device
        self.joints_category = joints_category
        self.num_obs = 24 if joints_category == "SMPL24" else 22
        # fit_batch(kernel="auto"): up to this many frames the warp-per-frame kernel is faster than the thread-per-frame
        # one, whose time is flat up to a full grid (measured on B200, 30 / 10 iterations, ms: L-BFGS B=4096 2.1 / 0.7
        # vs 8.8 / 2.9, B=16384 6.9 / 2.3 vs 9.7 / 2.8, B=32768 13.0 / 4.4 vs 10.8 / 3.1; Adam B=8192 2.3 / 0.8 vs
        # 4.2 / 1.6, B=16384 4.5 / 1.5 vs 4.3 / 1.6)
        self.warp_kernel_max_frames = int(os.environ.get("K2B_WARP_MAX_FRAMES", "16384" if use_lbfgs else "12288"))

        if isinstance(smpl_model, nat.NativeModel):
            self.native = smpl_model
        else:
            cached = getattr(smpl_model, "_k2b_native", None)
            key = (str(device), int(pose_prior_num_gaussians), prior_folder if gmm is None else id(gmm))
            if cached is not None and cached[0] == key:
                self.native = cached[1]
            else:
                weights = smpl_model if isinstance(smpl_model, BodyModelWeights) else extract_weights(smpl_model, model_type)
                consts: GMMConstants = prepare_gmm(gmm if gmm is not None else load_gmm(prior_folder, pose_prior_num_gaussians))
                self.native = nat.NativeModel(weights, consts, device)
                try:   # upload once per model object (the reference re-unpickles the prior per call)
                    smpl_model._k2b_native = (key, self.native)
                except Exception:
                    pass
        self.model_type = self.native.weights.model_type
        self.has_expr = self.native.num_shape == 20
        if joints_category == "GENERIC":
            # dict-block / explicit-index observations (world_space.py:198-201): supported where every index addresses
            # a body joint the kernels fit (see scatter_observations); the observation slots are the model's own
            self.num_obs = 24 if self.model_type == "smpl" else 22

    # ------------------------------------------------------------------ kernels
    def _run_fit(self, B, targets, conf, conf_per_frame, pose, betas, transl, expr, preserve, frame_iters,
                 frame_preserve, preserve_all, num_iters, optimizer, joint_loss_weight, pose_preserve_weight,
                 freeze_betas, want_joints=True, loss_kind=0, final_loss_mode=0, depth_ref=None, depth_weight=100.0):
        dev = self.device
        out_pose = torch.empty(B, 72, device=dev)
        out_betas = torch.empty(B, 10, device=dev)
        out_transl = torch.empty(B, 3, device=dev)
        out_expr = torch.empty(B, 10, device=dev) if self.has_expr else None
        out_loss = torch.empty(B, device=dev)
        out_joints = torch.empty(B, self.num_obs, 3, device=dev) if want_joints else None
        out_evals = torch.empty(B, dtype=torch.int32, device=dev)
        lib = self.native.lib
        ws_bytes = lib.k2b_fit_workspace_bytes(self.native.handle, B, optimizer, int(num_iters))
        ws = self.native.workspace("fit", ws_bytes)
        a = nat.FitArgs(
            num_frames=B, num_obs=self.num_obs, optimizer=optimizer, num_iters=int(num_iters),
            freeze_betas=int(freeze_betas), conf_per_frame=int(conf_per_frame), lr=self.step_size,
            joint_loss_weight=float(joint_loss_weight), pose_preserve_weight=float(pose_preserve_weight),
            targets=nat.ptr(targets), conf=nat.ptr(conf), init_pose=nat.ptr(pose), init_betas=nat.ptr(betas),
            init_transl=nat.ptr(transl), init_expr=nat.ptr(expr), preserve_pose=nat.ptr(preserve),
            frame_iters=nat.ptr(frame_iters), frame_preserve=nat.ptr(frame_preserve),
            preserve_all=int(preserve_all),
            out_pose=nat.ptr(out_pose), out_betas=nat.ptr(out_betas), out_transl=nat.ptr(out_transl),
            out_expr=nat.ptr(out_expr), out_loss=nat.ptr(out_loss), out_joints=nat.ptr(out_joints),
            out_evals=nat.ptr(out_evals), workspace=nat.ptr(ws), workspace_bytes=ws.numel(),
            loss_kind=int(loss_kind), final_loss_mode=int(final_loss_mode), depth_weight=float(depth_weight),
            depth_ref=nat.ptr(depth_ref),
        )
        with torch.cuda.device(dev):
            nat.check(lib.k2b_fit_batch(self.native.handle, C.byref(a), nat.current_stream()))
        return dict(pose=out_pose, betas=out_betas, transl=out_transl, expression=out_expr, loss=out_loss,
                    fit_joints=out_joints, evals=out_evals)

    def _run_chain(self, S, T, targets, conf, conf_mode, pose, betas, transl, expr, preserve, first_seq_ind, chain,
                   iters_first, iters_follow, optimizer, joint_loss_weight, pose_preserve_weight, freeze_betas,
                   want_joints=True, outs=None, window=None, time_major=False, seq_first=None, loss_kind=0,
                   final_loss_mode=0, depth_ref=None, depth_weight=100.0, camera_sequence=False):
        """One launch of the warp-per-sequence kernel (``k2b_fit_chain``): S sequences x T frames, serial in t.

        ``outs``: preallocated output dict (rows of this launch), else allocated here.  ``window = (a, b, T_total)``:
        the launch covers frames [a, b) of sequences that are T_total frames long in ``targets`` / ``conf``.
        """
        dev = self.device
        a0, b0, t_total = window if window is not None else (0, T, T)
        Tw = b0 - a0
        F = S * Tw
        if outs is None:
            outs = dict(pose=torch.empty(F, 72, device=dev), betas=torch.empty(F, 10, device=dev),
                        transl=torch.empty(F, 3, device=dev),
                        expression=torch.empty(F, 10, device=dev) if self.has_expr else None,
                        loss=torch.empty(F, device=dev),
                        fit_joints=torch.empty(F, self.num_obs, 3, device=dev) if want_joints else None,
                        evals=torch.empty(F, dtype=torch.int32, device=dev))
        lib = self.native.lib
        ws_bytes = lib.k2b_chain_workspace_bytes(self.native.handle, S, optimizer, int(max(iters_first, iters_follow)))
        ws = self.native.workspace("chain", ws_bytes)
        K = self.num_obs
        tgt_ptr = targets.data_ptr() + 4 * a0 * K * 3
        conf_ptr = None if conf is None else conf.data_ptr() + (4 * a0 * K if conf_mode == 2 else 0)
        keep_ptr = None if preserve is None else preserve.data_ptr() + 4 * a0 * 69
        a = nat.ChainArgs(
            num_sequences=S, frames_per_sequence=Tw, num_obs=K, optimizer=optimizer,
            num_iters_first=int(iters_first), num_iters_followup=int(iters_follow),
            first_seq_ind=int(first_seq_ind) + a0, chain_init=int(bool(chain)), freeze_betas=int(freeze_betas),
            conf_mode=int(conf_mode), out_time_major=int(bool(time_major)), in_sequence_stride=int(t_total),
            lr=self.step_size, joint_loss_weight=float(joint_loss_weight),
            pose_preserve_weight=float(pose_preserve_weight),
            targets=tgt_ptr, conf=conf_ptr, init_pose=nat.ptr(pose), init_betas=nat.ptr(betas),
            init_transl=nat.ptr(transl), init_expr=nat.ptr(expr), preserve_pose=keep_ptr,
            seq_first_ind=nat.ptr(seq_first), out_pose=nat.ptr(outs["pose"]), out_betas=nat.ptr(outs["betas"]), out_transl=nat.ptr(outs["transl"]),
            out_expr=nat.ptr(outs["expression"]), out_loss=nat.ptr(outs["loss"]),
            out_joints=nat.ptr(outs["fit_joints"]), out_evals=nat.ptr(outs["evals"]), workspace=nat.ptr(ws),
            workspace_bytes=ws.numel(), loss_kind=int(loss_kind), final_loss_mode=int(final_loss_mode),
            depth_weight=float(depth_weight), depth_ref=nat.ptr(depth_ref), camera_sequence=int(bool(camera_sequence)),
        )
        with torch.cuda.device(dev):
            nat.check(lib.k2b_fit_chain(self.native.handle, C.byref(a), nat.current_stream()))
        return outs

    def evaluate_batch(self, params: dict, j3d, conf=None, preserve_pose=None, preserve_on=False,
                       joint_loss_weight=600.0, pose_preserve_weight=5.0, kernel="frame"):
        """One evaluation of the loss and its gradient (parity / debugging entry).  ``kernel``: "frame" = the
        one-thread-per-frame evaluator (what ``k2b_fit_batch`` runs), "warp" = the warp-per-frame evaluator (what
        ``k2b_fit_chain`` runs)."""
        dev = self.device
        go, bp = _f32(params["global_orient"], dev), _f32(params["body_pose"], dev)
        B = go.shape[0]
        _check_widths(B, global_orient=(go, 3), body_pose=(bp, 69), betas=(_f32(params["betas"], dev), 10),
                      transl=(_f32(params["transl"], dev), 3), preserve_pose=(_f32(preserve_pose, dev), 69))
        pose = torch.cat([go, bp], dim=1).contiguous()
        targets = _f32(j3d, dev)[:, : self.num_obs].contiguous()
        conf = _f32(conf, dev)
        conf_pf = conf is not None and conf.dim() == 2
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
        expr = _f32(params.get("expression"), dev)
        if self.has_expr and expr is None:
            expr = torch.zeros(B, 10, device=dev)
        # keep every converted input alive until the launch has been enqueued
        betas, transl, keep = _f32(params["betas"], dev), _f32(params["transl"], dev), _f32(preserve_pose, dev)
        outs = dict(loss=torch.empty(B, device=dev), grad_pose=torch.empty(B, 72, device=dev),
                    grad_betas=torch.empty(B, 10, device=dev), grad_transl=torch.empty(B, 3, device=dev),
                    grad_expression=torch.empty(B, 10, device=dev) if self.has_expr else None,
                    joints=torch.empty(B, self.num_obs, 3, device=dev),
                    gmm_component=torch.empty(B, dtype=torch.int32, device=dev))
        lib = self.native.lib
        ws = self.native.workspace("fit", lib.k2b_fit_workspace_bytes(self.native.handle, B, nat.OPT_ADAM, 1))
        a = nat.EvalArgs(
            num_frames=B, num_obs=self.num_obs, conf_per_frame=int(conf_pf), preserve_all=int(bool(preserve_on)),
            joint_loss_weight=float(joint_loss_weight), pose_preserve_weight=float(pose_preserve_weight),
            targets=nat.ptr(targets), conf=nat.ptr(conf), pose=nat.ptr(pose),
            betas=nat.ptr(betas), transl=nat.ptr(transl), expr=nat.ptr(expr), preserve_pose=nat.ptr(keep),
            out_loss=nat.ptr(outs["loss"]), out_grad_pose=nat.ptr(outs["grad_pose"]),
            out_grad_betas=nat.ptr(outs["grad_betas"]), out_grad_transl=nat.ptr(outs["grad_transl"]),
            out_grad_expr=nat.ptr(outs["grad_expression"]), out_joints=nat.ptr(outs["joints"]),
            out_gmm_component=nat.ptr(outs["gmm_component"]), workspace=nat.ptr(ws), workspace_bytes=ws.numel(),
            warp_evaluator=int(kernel == "warp"),
        )
        with torch.cuda.device(dev):
            nat.check(lib.k2b_evaluate_batch(self.native.handle, C.byref(a), nat.current_stream()))
        return outs

    def forward_batch(self, params: dict, with_vertices=True, out_vertices=None, out_joints=None, max_ctas=0):
        """Body-model forward (mesh kernels): joints ``(B, n_j + extras, 3)`` and vertices ``(B,V,3)``."""
        dev = self.device
        go = _f32(params["global_orient"], dev)
        B = go.shape[0]
        extras = {k: _f32(params.get(k), dev) for k in _EXTRA_BLOCKS}
        full_pose = full_pose_from_params(self.model_type, go, _f32(params["body_pose"], dev), extras).contiguous()
        betas = _f32(params["betas"], dev)
        if betas.shape[0] != B:
            betas = betas.expand(B, -1)
        shape = betas
        if self.has_expr:
            expr = extras.get("expression")
            shape = torch.cat([betas, expr if expr is not None else torch.zeros(B, 10, device=dev)], dim=1)
        shape = shape.contiguous()
        transl = _f32(params.get("transl"), dev)
        n_out = self.native.num_joints + self.native.num_extra
        joints = out_joints if out_joints is not None else torch.empty(B, n_out, 3, device=dev)
        verts = None
        if with_vertices:
            verts = out_vertices if out_vertices is not None else torch.empty(B, self.native.num_vertices, 3, device=dev)
        lib = self.native.lib
        ws = self.native.workspace("mesh", lib.k2b_mesh_workspace_bytes(self.native.handle, B))
        a = nat.MeshArgs(num_frames=B, full_pose=nat.ptr(full_pose), shape=nat.ptr(shape), transl=nat.ptr(transl),
                         out_vertices=nat.ptr(verts), out_joints=nat.ptr(joints), workspace=nat.ptr(ws),
                         workspace_bytes=ws.numel(), max_ctas=int(max_ctas))
        with torch.cuda.device(dev):
            nat.check(lib.k2b_mesh_batch(self.native.handle, C.byref(a), nat.current_stream()))
        return {"joints": joints, "vertices": verts}

    def shape_pass(self, init_betas, pose_init, j3d_world, *, frame_indices=None, conf=None, num_iters=40,
                   step_size=1e-1, shape_prior_weight=5.0, num_sequences=1):
        """Shared-betas L-BFGS pre-pass (core/shape.py:10-115) for ``num_sequences`` sequences.

        ``j3d_world`` is (S*T, K, 3) with sequences back to back, ``pose_init`` (S*T, 72) the fixed
        poses; ``frame_indices`` must be a prefix ``range(n)`` (what the reference's driver passes,
        engine.py:244-248).  Returns betas (S,10) -- (1,10) for a single sequence.
        """
        dev = self.device
        targets = _f32(j3d_world, dev)[:, : self.num_obs].contiguous()
        S = int(num_sequences)
        T = targets.shape[0] // S
        n_use = T if frame_indices is None else len(frame_indices)
        if frame_indices is not None and list(frame_indices) != list(range(n_use)):
            raise NotImplementedError("shape pass supports a leading range of frames only")
        poses = _f32(pose_init, dev).contiguous()
        betas0 = _f32(init_betas, dev)
        if betas0.shape[0] != S:
            betas0 = betas0.expand(S, -1)
        betas0 = betas0.contiguous()
        conf = _f32(conf, dev)
        conf_ps = conf is not None and conf.dim() == 2
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
        out_betas = torch.empty(S, 10, device=dev)
        out_loss = torch.empty(S, device=dev)
        out_evals = torch.empty(S, dtype=torch.int32, device=dev)
        lib = self.native.lib
        ws = self.native.workspace("shape", lib.k2b_shape_workspace_bytes(self.native.handle, S, int(num_iters)))
        a = nat.ShapeArgs(num_sequences=S, frames_per_sequence=n_use, sequence_stride=T, num_obs=self.num_obs,
                          pose_per_frame=1, conf_per_sequence=int(conf_ps), num_iters=int(num_iters),
                          lr=float(step_size), shape_prior_weight=float(shape_prior_weight),
                          targets=nat.ptr(targets), poses=nat.ptr(poses), conf=nat.ptr(conf),
                          init_betas=nat.ptr(betas0), out_betas=nat.ptr(out_betas), out_loss=nat.ptr(out_loss),
                          out_evals=nat.ptr(out_evals), workspace=nat.ptr(ws), workspace_bytes=ws.numel())
        with torch.cuda.device(dev):
            nat.check(lib.k2b_shape_pass(self.native.handle, C.byref(a), nat.current_stream()))
        self.last_shape_evals = out_evals
        return out_betas

    def body_joints_only(self, model_indices) -> bool:
        """True when every index addresses a fitted body joint exactly once: the body-keypoint kernels apply
        (observations scattered into their slots); otherwise the general articulated fit takes the call."""
        idx = [int(i) for i in torch.as_tensor(model_indices).reshape(-1).tolist()]
        return all(0 <= i < self.num_obs for i in idx) and len(set(idx)) == len(idx)

    def scatter_observations(self, j3d, conf, model_indices):
        """Observations given against explicit model-joint indices (``target_model_indices``, the reference's
        GENERIC path, world_space.py:198-201) -> this fitter's fixed observation slots.

        ``j3d`` (B,K,3), ``conf`` (K,) | (B,K) | None, ``model_indices`` (K,).  Slot i of the result holds the
        observation of model joint i; joints nobody observed get confidence 0, i.e. no loss and no gradient,
        which is exactly the sum the reference forms over the observed joints only.  Indices beyond the fitted
        body joints (SMPL-H / SMPL-X hand joints, vertex-picked fingertips and face landmarks) are served by
        ``fit_frame`` through the general articulated fit (csrc/artic_core.cuh), not by the batched kernels.
        """
        dev = self.device
        idx = [int(i) for i in torch.as_tensor(model_indices).reshape(-1).tolist()]
        if any(i < 0 or i >= self.num_obs for i in idx):
            raise NotImplementedError(
                f"observations of model joints >= {self.num_obs} (hand joints, vertex-picked fingertips / face "
                "landmarks) go through fit_frame / optimize_params_frame / optimize_params_sequence (the general "
                "articulated fit); the batched body-keypoint entry points take body joints only")
        if len(set(idx)) != len(idx):
            raise NotImplementedError("a model joint observed more than once: use fit_frame (general articulated fit)")
        j3d = _f32(j3d, dev)
        if j3d.dim() != 3 or j3d.shape[1] != len(idx):
            raise ValueError(f"j3d must be (B, {len(idx)}, 3) for {len(idx)} model indices, got {tuple(j3d.shape)}")
        B = j3d.shape[0]
        sel = torch.tensor(idx, device=dev, dtype=torch.long)
        full = torch.zeros(B, self.num_obs, 3, device=dev)
        full[:, sel] = j3d
        conf = _f32(conf, dev)
        if conf is None:
            conf = torch.ones(len(idx), device=dev)
        cfull = torch.zeros(conf.shape[:-1] + (self.num_obs,), device=dev)
        cfull[..., sel] = conf
        return full, cfull

    def _fit_frame_articulated(self, init_params, j3d, conf_3d, seq_ind, idx, joint_loss_weight, pose_preserve_weight,
                               freeze_betas) -> BodyModelFitResult:
        """``fit_frame`` for observations beyond the body joints (world_space.py:198-201): every block the caller
        supplied is optimised, in the reference's order; blocks left at None stay at the model's zero default."""
        from .articulated import articulated_fit, get_articulated

        am = get_articulated(self.native, with_body_priors=True)
        blocks = {k: getattr(init_params, k, None) for k, _ in am.blocks}
        if conf_3d is not None:
            conf_3d = _f32(conf_3d, self.device)
            if conf_3d.dim() == 2:
                conf_3d = conf_3d[0]       # reference quirk, world_space.py:163-164
            conf_3d = conf_3d[: len(idx)].contiguous()
        iters = self.num_iters_first if seq_ind == 0 else self.num_iters_followup
        p, loss, _evals, _pts = articulated_fit(am, blocks, j3d, conf_3d, idx, seq_ind=int(seq_ind), num_iters=iters,
                                                use_lbfgs=self.use_lbfgs, lr=self.step_size,
                                                joint_loss_weight=joint_loss_weight,
                                                pose_preserve_weight=pose_preserve_weight, freeze_betas=freeze_betas)
        given = {k: v for k, v in p.items() if blocks.get(k) is not None}
        mesh = self.forward_batch(dict(p))
        base = dict(betas=p["betas"], global_orient=p["global_orient"], body_pose=p["body_pose"], transl=p["transl"])
        if isinstance(init_params, SMPLXData):
            fitted = SMPLXData(**base, **{k: given.get(k) for k in ("left_hand_pose", "right_hand_pose", "expression",
                                                                    "jaw_pose", "leye_pose", "reye_pose")})
        elif isinstance(init_params, SMPLHData):
            fitted = SMPLHData(**base, left_hand_pose=given.get("left_hand_pose"), right_hand_pose=given.get("right_hand_pose"))
        else:
            fitted = SMPLData(**base)
        return BodyModelFitResult(params=fitted, vertices=mesh["vertices"], joints=mesh["joints"], loss=loss.sum())

    # ------------------------------------------------------------------ public
    def fit_batch(self, init: dict, j3d, conf=None, *, seq_ind=0, preserve_pose=None, num_iters=None,
                  joint_loss_weight=600.0, pose_preserve_weight=5.0, freeze_betas=False, use_lbfgs=None,
                  with_mesh=True, out_vertices=None, kernel="auto", fit_joints=True):
        """Fit B independent frames in one launch.

        ``fit_joints=False``: the fit kernel does not return the posed kinematic joints (the mesh pass returns all
        joints anyway), which spares an L-BFGS fit its extra forward pass at the returned parameters.

        ``kernel``: "frame" = one thread per frame (``k2b_fit_batch``, throughput), "warp" = one warp per frame
        (``k2b_fit_chain``, latency), "auto" = warp for small batches.

        ``init``: dict of (B,dim) arrays with keys global_orient, body_pose, betas, transl
        (+ SMPL-H / SMPL-X blocks).  ``seq_ind``: int or (B,) integer tensor; 0 selects the
        first-frame budget without the temporal term, > 0 the follow-up budget with it
        (world_space.py:211,214).  ``preserve_pose``: (B,69) temporal anchor, default the
        initial body pose (world_space.py:159).  Returns a dict of device tensors.
        """
        dev = self.device
        go, bp = _f32(init["global_orient"], dev), _f32(init["body_pose"], dev)
        transl = init.get("transl")
        if transl is None:
            raise ValueError("init_params.transl must be provided")
        B = go.shape[0]
        betas = _f32(init["betas"], dev)
        transl = _f32(transl, dev)
        keep_pose = _f32(preserve_pose, dev)
        _check_widths(B, global_orient=(go, 3), body_pose=(bp, 69), betas=(betas, 10), transl=(transl, 3),
                      preserve_pose=(keep_pose, 69), expression=(_f32(init.get("expression"), dev), 10))
        if bp.shape[0] != B or transl.shape[0] != B or (keep_pose is not None and keep_pose.shape[0] != B):
            raise ValueError("body_pose, transl and preserve_pose need one row per frame")
        pose = torch.cat([go, bp], dim=1).contiguous()
        if betas.shape[0] != B:
            betas = betas.expand(B, -1).contiguous()
        targets = _f32(j3d, dev)
        if targets.dim() != 3 or targets.shape[0] != B or targets.shape[1] < self.num_obs:
            raise ValueError(f"j3d must be (B={B}, K>={self.num_obs}, 3), got {tuple(targets.shape)}")
        targets = targets[:, : self.num_obs].contiguous()
        conf = _f32(conf, dev)
        conf_pf = conf is not None and conf.dim() == 2
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
        extras = {k: _f32(init.get(k), dev) for k in _EXTRA_BLOCKS}
        expr = extras["expression"]
        freeze = int(bool(freeze_betas))
        if self.has_expr and expr is None:
            # the reference optimises the expression only when the caller supplied one (world_space.py:222-223);
            # otherwise the model's zero default stays fixed and the result carries expression=None
            expr = torch.zeros(B, 10, device=dev)
            freeze |= nat.FREEZE_EXPR
        elif self.has_expr and expr.shape[0] != B:
            expr = expr.expand(B, -1).contiguous()

        frame_iters = frame_preserve = None
        if isinstance(seq_ind, torch.Tensor):
            seq = seq_ind.to(dev)
            first = seq == 0
            n_first = self.num_iters_first if num_iters is None else num_iters
            n_follow = self.num_iters_followup if num_iters is None else num_iters
            frame_iters = torch.where(first, n_first, n_follow).to(torch.int32).contiguous()
            frame_preserve = (~first).to(torch.uint8).contiguous()
            budget, preserve_all = max(n_first, n_follow), 0
        else:
            budget = num_iters if num_iters is not None else (
                self.num_iters_first if seq_ind == 0 else self.num_iters_followup)
            preserve_all = int(seq_ind > 0)
        lbfgs = self.use_lbfgs if use_lbfgs is None else use_lbfgs
        if kernel == "auto":
            # few frames: one thread per frame leaves the GPU idle and a frame takes ~1.3 ms; a warp per frame
            # (k2b_fit_chain with one-frame sequences) finishes in a fraction of that
            kernel = "warp" if B <= self.warp_kernel_max_frames else "frame"
        if kernel == "warp":
            if conf is not None and conf_pf:
                conf = conf.reshape(B, 1, self.num_obs)
            if frame_iters is not None:      # per-frame seq_ind: every frame is a one-frame sequence with its own index
                seq_first, i_first, i_follow, s0 = seq.to(torch.int32).contiguous(), n_first, n_follow, 0
            else:
                seq_first, i_first, i_follow, s0 = None, budget, budget, (1 if preserve_all else 0)
            res = self._run_chain(B, 1, targets.reshape(B, 1, self.num_obs, 3), conf,
                                  0 if conf is None else (2 if conf_pf else 1), pose, betas, transl,
                                  expr if self.has_expr else None, keep_pose, s0, True, i_first, i_follow,
                                  nat.OPT_LBFGS if lbfgs else nat.OPT_ADAM, joint_loss_weight,
                                  pose_preserve_weight, freeze, seq_first=seq_first, want_joints=fit_joints)
        else:
            res = self._run_fit(B, targets, conf, conf_pf, pose, betas, transl, expr if self.has_expr else None,
                                keep_pose, frame_iters, frame_preserve, preserve_all, budget,
                                nat.OPT_LBFGS if lbfgs else nat.OPT_ADAM, joint_loss_weight, pose_preserve_weight,
                                freeze, want_joints=fit_joints)
        params = {"global_orient": res["pose"][:, :3], "body_pose": res["pose"][:, 3:], "betas": res["betas"],
                  "transl": res["transl"]}
        for k in _EXTRA_BLOCKS:
            if extras[k] is not None:
                params[k] = extras[k]      # receive no gradient from body keypoints: passed through
        if self.has_expr and not (freeze & nat.FREEZE_EXPR):
            params["expression"] = res["expression"]
        out = {"params": params, "loss": res["loss"], "evals": res["evals"], "fit_joints": res["fit_joints"]}
        if with_mesh:
            out.update(self.forward_batch(params, with_vertices=True, out_vertices=out_vertices))
        return out

    def fit_chain(self, init: dict, j3d, conf=None, *, first_seq_ind=0, chain=True, joint_loss_weight=600.0,
                  pose_preserve_weight=5.0, freeze_betas=False, use_lbfgs=None, with_mesh=True, out_vertices=None,
                  time_major=False, chunks=1, params_ready=None, mesh_capped_fraction=None, fit_joints=True,
                  window_done=None, window_ready=None):
        """Fit S sequences of T frames each the way the reference's sequence loop does (api/sequence.py:214-281):
        serially in t, frame t starting from frame t-1's result (``chain=True``) or from the sequence's
        initialisation (``chain=False``) -- one warp per sequence, all frames inside one launch.

        ``init``: dict of (S,dim) arrays = initialisation of every sequence's frame 0.  ``j3d``: (S,T,K,3).
        ``conf``: None, (K,) or (S,T,K).  Returns the same dict as ``fit_batch`` with S*T leading rows, sequence-major
        (row s*T + t) or, with ``time_major``, time-major (row t*S + s).

        ``chunks`` > 1 (needs ``time_major`` or S == 1, and ``chain``): the time axis is cut into that many windows;
        each window is one launch on a high-priority stream that continues from the previous window's last frame
        (same arithmetic, same results), and the mesh pass of a finished window runs on the caller's stream
        while the next window is being fitted -- the fit leaves most of every SM idle when there are few
        sequences.  ``params_ready``: optional event recorded when the fitted parameters are final.
        ``mesh_capped_fraction``: share of the windows whose mesh pass is held to the SMs the fit leaves free
        (the rest, at the end, run on all SMs).  Default (None): chosen by measurement -- the first calls of a
        problem shape (S, T, chunks, optimiser) each run with one candidate share and are timed with CUDA events
        on the caller's stream (nine calls: a cold one, then four candidates twice), later calls use the fastest (``overlap_policy()`` reports what was measured);
        a number, or the environment variable K2B_MESH_CAPPED_FRACTION, pins it.  ``fit_joints=False``: the fit kernel does not return the posed kinematic joints
        (``out["fit_joints"]`` is None; the mesh pass returns all joints anyway), which also spares the L-BFGS fit its
        extra forward pass at the returned parameters -- the returned loss is the accepted trial's, bit for bit.
        ``window_done(c, rows, fit_done, outs, joints)``: called after window c's launches have been enqueued -- ``rows`` of
        the raw output buffers ``outs`` (pose / betas / transl / loss / evals) are final once the event ``fit_done`` has
        passed, ``joints[rows]`` (and the vertices) once everything enqueued so far on the current stream has run; a caller
        can stream a window's results to the host while later windows are fitted.  ``window_ready``: optional list of
        ``chunks`` events; the fit of window c waits for event c (its keypoints may still be on their way to the device).
        """
        dev = self.device
        targets = _f32(j3d, dev)
        if targets.dim() != 4 or targets.shape[2] < self.num_obs:
            raise ValueError(f"j3d must be (S, T, K>={self.num_obs}, 3), got {tuple(targets.shape)}")
        S, T = targets.shape[0], targets.shape[1]
        targets = targets[:, :, : self.num_obs].contiguous()
        go, bp = _f32(init["global_orient"], dev), _f32(init["body_pose"], dev)
        if init.get("transl") is None:
            raise ValueError("init_params.transl must be provided")
        _check_widths(S, global_orient=(go, 3), body_pose=(bp, 69), betas=(_f32(init["betas"], dev), 10),
                      transl=(_f32(init["transl"], dev), 3), expression=(_f32(init.get("expression"), dev), 10))
        pose = torch.cat([go, bp], dim=1).expand(S, -1).contiguous()
        betas = _f32(init["betas"], dev).expand(S, -1).contiguous()
        transl = _f32(init["transl"], dev).expand(S, -1).contiguous()
        conf = _f32(conf, dev)
        conf_mode = 0
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
            conf_mode = 1 if conf.dim() == 1 else 2
            if conf_mode == 2 and conf.numel() != S * T * self.num_obs:
                raise ValueError("per-frame confidences must be (S, T, K)")
        extras = {k: _f32(init.get(k), dev) for k in _EXTRA_BLOCKS}
        expr = extras["expression"]
        freeze = int(bool(freeze_betas))
        if self.has_expr:
            if expr is None:        # no expression supplied: it stays at the model's zero default (world_space.py:222-223)
                freeze |= nat.FREEZE_EXPR
            expr = (expr if expr is not None else torch.zeros(S, 10, device=dev)).expand(S, -1).contiguous()
        lbfgs = self.use_lbfgs if use_lbfgs is None else use_lbfgs
        optimizer = nat.OPT_LBFGS if lbfgs else nat.OPT_ADAM
        chunks = max(1, min(int(chunks), T))
        if chunks > 1 and not (chain and (time_major or S == 1)):
            raise ValueError("chunks > 1 needs chain=True and time_major=True (or a single sequence)")
        tm = bool(time_major) and S > 1
        F = S * T
        outs = dict(pose=torch.empty(F, 72, device=dev), betas=torch.empty(F, 10, device=dev),
                    transl=torch.empty(F, 3, device=dev),
                    expression=torch.empty(F, 10, device=dev) if self.has_expr else None,
                    loss=torch.empty(F, device=dev),
                    fit_joints=torch.empty(F, self.num_obs, 3, device=dev) if fit_joints else None,
                    evals=torch.empty(F, dtype=torch.int32, device=dev))
        common = (self.num_iters_first, self.num_iters_followup, optimizer, joint_loss_weight, pose_preserve_weight,
                  freeze)

        def params_of(rows):
            p = {"global_orient": outs["pose"][rows, :3], "body_pose": outs["pose"][rows, 3:],
                 "betas": outs["betas"][rows], "transl": outs["transl"][rows]}
            n = outs["pose"][rows].shape[0]
            for k in _EXTRA_BLOCKS:
                if extras[k] is not None:    # receive no gradient from body keypoints: passed through to every frame
                    e = extras[k].expand(S, -1)
                    p[k] = e.repeat(n // S, 1) if tm else e.repeat_interleave(n // S, dim=0)
            if self.has_expr and not (freeze & nat.FREEZE_EXPR):
                p["expression"] = outs["expression"][rows]
            return p

        out = {"loss": outs["loss"], "evals": outs["evals"], "fit_joints": outs["fit_joints"]}
        if chunks == 1:
            self._run_chain(S, T, targets, conf, conf_mode, pose, betas, transl, expr if self.has_expr else None, None,
                            first_seq_ind, chain, *common, outs=outs, time_major=tm)
            if params_ready is not None:
                params_ready.record()
            out["params"] = params_of(slice(0, F))
            if with_mesh:
                out.update(self.forward_batch(out["params"], with_vertices=True, out_vertices=out_vertices))
            return out

        # ---- windows of the time axis: fit on a high-priority stream, mesh of finished windows behind it --------
        n_out = self.native.num_joints + self.native.num_extra
        joints = torch.empty(F, n_out, 3, device=dev) if with_mesh else None
        verts = None
        if with_mesh:
            verts = out_vertices if out_vertices is not None else torch.empty(F, self.native.num_vertices, 3, device=dev)
        cur = torch.cuda.current_stream(dev)
        if getattr(self, "_chain_stream", None) is None:
            self._chain_stream = torch.cuda.Stream(device=dev, priority=-1)
        hp = self._chain_stream
        hp.wait_stream(cur)
        timing = getattr(self, "chain_events", None)
        if timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(hp)
        # while windows are still being fitted, the mesh kernel is held to the SMs the fit leaves free (its CTAs are
        # long-lived and would otherwise keep the next window's CTAs waiting); the last windows' mesh runs everywhere
        ctas, warps = C.c_int32(), C.c_int32()
        n_sms = self.native.lib.k2b_chain_geometry(self.native.handle, S, C.byref(ctas), C.byref(warps))
        free_sms = n_sms - ctas.value
        frac = mesh_capped_fraction
        if frac is None and os.environ.get("K2B_MESH_CAPPED_FRACTION"):
            frac = float(os.environ["K2B_MESH_CAPPED_FRACTION"])
        tuner = None
        if frac is None:
            if with_mesh and free_sms >= 8:
                tuner = self._overlap_tuner((S, T, chunks, bool(lbfgs)))
                frac = tuner.next_fraction()
                tuner.begin(cur)
            else:
                frac = 0.0
        capped = min(chunks - 1, int(round(chunks * frac))) if free_sms >= 8 else 0
        bounds = [(T * c) // chunks for c in range(chunks + 1)]
        init_c = (pose, betas, transl, expr if self.has_expr else None)
        for c in range(chunks):
            a0, b0 = bounds[c], bounds[c + 1]
            rows = slice(a0 * S, b0 * S)
            o = {k: (v[rows] if v is not None else None) for k, v in outs.items()}
            with torch.cuda.stream(hp):
                if window_ready is not None:
                    hp.wait_event(window_ready[c])
                self._run_chain(S, T, targets, conf, conf_mode, *init_c, None, first_seq_ind, True, *common, outs=o,
                                window=(a0, b0, T), time_major=tm)
                done = torch.cuda.Event()
                done.record(hp)
            last = slice((b0 - 1) * S, b0 * S)      # the next window starts from this window's last frame
            init_c = (outs["pose"][last], outs["betas"][last], outs["transl"][last],
                      outs["expression"][last] if self.has_expr else None)
            if c == chunks - 1:
                if timing is not None:
                    t1.record(hp)
                    timing.append((t0, t1))
                if params_ready is not None:
                    params_ready.record(hp)
            cur.wait_event(done)
            if with_mesh:
                self.forward_batch(params_of(rows), with_vertices=True, out_vertices=verts[rows], out_joints=joints[rows],
                                   max_ctas=free_sms if c < capped else 0)
            if window_done is not None:
                window_done(c, rows, done, outs, joints if with_mesh else None)
        if tuner is not None:
            tuner.end(cur)
        out["params"] = params_of(slice(0, F))
        if with_mesh:
            out["joints"], out["vertices"] = joints, verts
        return out

    def _overlap_tuner(self, key):
        tuners = self.__dict__.setdefault("_overlap_tuners", {})
        if key not in tuners:
            tuners[key] = _OverlapTuner()
        return tuners[key]

    def overlap_policy(self) -> dict:
        """What the mesh-overlap policy of ``fit_chain(chunks > 1)`` has measured so far:
        ``{(S, T, chunks, lbfgs): {"fraction": chosen or None, "ms": {candidate: milliseconds}}}``."""
        out = {}
        for key, t in self.__dict__.get("_overlap_tuners", {}).items():
            t.collect()
            out[key] = {"fraction": t.best, "ms": dict(t.times)}
        return out

    def fit_frame(
        self,
        init_params: SMPLData,
        j3d: torch.Tensor,
        conf_3d: Optional[torch.Tensor] = None,
        seq_ind: int = 0,
        target_model_indices: Optional[torch.Tensor] = None,
        joint_loss_weight: float = 600.0,
        pose_preserve_weight: float = 5.0,
        freeze_betas: bool = False,
    ) -> BodyModelFitResult:
        """Reference-shaped single call (world_space.py:93-323)."""
        if init_params.transl is None:
            raise ValueError("init_params.transl must be provided")
        if target_model_indices is not None:
            idx = [int(i) for i in torch.as_tensor(target_model_indices).reshape(-1).tolist()]
            if not self.body_joints_only(idx):
                # hand joints, vertex-picked finger tips / face landmarks (or repeated indices): the general
                # articulated fit (csrc/artic_core.cuh) instead of the body-keypoint kernels
                return self._fit_frame_articulated(init_params, j3d, conf_3d, seq_ind, idx, joint_loss_weight,
                                                   pose_preserve_weight, freeze_betas)
            j3d, conf_3d = self.scatter_observations(j3d, conf_3d, target_model_indices)
        elif self.joints_category == "GENERIC":
            raise ValueError("joints_category='GENERIC' needs target_model_indices")
        init = {k: getattr(init_params, k) for k in ("global_orient", "body_pose", "betas", "transl")}
        if isinstance(init_params, SMPLHData):
            init["left_hand_pose"] = init_params.left_hand_pose
            init["right_hand_pose"] = init_params.right_hand_pose
        if isinstance(init_params, SMPLXData):
            for k in ("expression", "jaw_pose", "leye_pose", "reye_pose"):
                init[k] = getattr(init_params, k)
        if self.has_expr and not isinstance(init_params, SMPLXData):
            raise ValueError("an SMPL-X model needs SMPLXData init_params")
        if conf_3d is not None:
            conf_3d = _f32(conf_3d, self.device)
            if conf_3d.dim() == 2:
                conf_3d = conf_3d[0]       # reference quirk, world_space.py:163-164
        out = self.fit_batch(init, j3d, conf_3d, seq_ind=int(seq_ind), joint_loss_weight=joint_loss_weight,
                             pose_preserve_weight=pose_preserve_weight, freeze_betas=freeze_betas)
        p = out["params"]
        base = dict(betas=p["betas"], global_orient=p["global_orient"], body_pose=p["body_pose"], transl=p["transl"])
        if isinstance(init_params, SMPLXData):
            fitted = SMPLXData(**base, left_hand_pose=p.get("left_hand_pose"), right_hand_pose=p.get("right_hand_pose"),
                               expression=p.get("expression") if init_params.expression is not None else None,
                               jaw_pose=p.get("jaw_pose"), leye_pose=p.get("leye_pose"), reye_pose=p.get("reye_pose"))
        elif isinstance(init_params, SMPLHData):
            fitted = SMPLHData(**base, left_hand_pose=p.get("left_hand_pose"), right_hand_pose=p.get("right_hand_pose"))
        else:
            fitted = SMPLData(**base)
        loss = out["loss"].sum()   # losses.py:67 sums over the batch
        return BodyModelFitResult(params=fitted, vertices=out["vertices"], joints=out["joints"], loss=loss)
