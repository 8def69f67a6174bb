"""Batched sequence fitting on device-resident (or pinned-host) frame arrays.

This is the public bulk entry point the benchmark measures: it fits F frames that belong to any
number of back-to-back sequences with the frame-parallel two-sweep schedule S2 (see
``api/sequence.py``) in two launches of the fused kernel plus the mesh pass, with the sweep-0
parameters handed to sweep 1 as a zero-copy view shifted by one frame.
"""

from __future__ import annotations

from typing import Optional

import torch

from .. import _native as nat
from ..core.config import FrameOptimizeConfig
from ..core.fitters.world_space import WorldSpaceFitter
from ..distributed import exchange_halo, plan_two_sweep


class SequenceBatchFitter:
    """Preallocated buffers + schedule S2 for ``num_frames`` frames on one GPU."""

    def __init__(self, fitter: WorldSpaceFitter, num_frames: int, frame_cfg: Optional[FrameOptimizeConfig] = None,
                 with_vertices: bool = True):
        self.f = fitter
        self.F = int(num_frames)
        self.cfg = frame_cfg or FrameOptimizeConfig()
        dev, F = fitter.device, self.F
        ne = 10 if fitter.has_expr else 0
        # sweep-0 outputs carry one extra leading row: the halo (left neighbour's last frame)
        self.s0 = dict(pose=torch.zeros(F + 1, 72, device=dev), betas=torch.zeros(F + 1, 10, device=dev),
                       transl=torch.zeros(F + 1, 3, device=dev),
                       expr=torch.zeros(F + 1, 10, device=dev) if ne else None)
        self.s1 = dict(pose=torch.empty(F, 72, device=dev), betas=torch.empty(F, 10, device=dev),
                       transl=torch.empty(F, 3, device=dev), expr=torch.empty(F, 10, device=dev) if ne else None)
        self.loss0 = torch.empty(F, device=dev)
        self.loss1 = torch.empty(F, device=dev)
        self.evals0 = torch.empty(F, dtype=torch.int32, device=dev)
        self.evals1 = torch.empty(F, dtype=torch.int32, device=dev)
        self.init_pose = torch.zeros(F, 72, device=dev)
        self.init_betas = torch.zeros(F, 10, device=dev)
        self.init_transl = torch.empty(F, 3, device=dev)
        self.init_expr = torch.zeros(F, 10, device=dev) if ne else None
        n_out = fitter.native.num_joints + fitter.native.num_extra
        self.joints = torch.empty(F, n_out, 3, device=dev)
        self.vertices = torch.empty(F, fitter.native.num_vertices, 3, device=dev) if with_vertices else None
        self.full_pose = torch.empty(F, 3 * fitter.native.num_joints, device=dev)
        self.model_root0 = None
        self.kernel_events = None

    def _fit(self, targets, conf, init, preserve, iters, preserve_flags, budget, out, loss, evals, optimizer):
        f, F = self.f, self.F
        lib = f.native.lib
        ws_bytes = lib.k2b_fit_workspace_bytes(f.native.handle, F, optimizer, int(budget))
        ws = f.native.workspace("fit", ws_bytes)
        import ctypes as C

        conf_pf = conf is not None and conf.dim() == 2
        a = nat.FitArgs(
            num_frames=F, num_obs=f.num_obs, optimizer=optimizer, num_iters=int(budget),
            freeze_betas=int(self.cfg.freeze_betas), conf_per_frame=int(conf_pf), lr=f.step_size,
            joint_loss_weight=float(self.cfg.joint_loss_weight),
            pose_preserve_weight=float(self.cfg.pose_preserve_weight),
            targets=nat.ptr(targets), conf=nat.ptr(conf), init_pose=nat.ptr(init["pose"]),
            init_betas=nat.ptr(init["betas"]), init_transl=nat.ptr(init["transl"]),
            init_expr=nat.ptr(init["expr"]), preserve_pose=nat.ptr(preserve), frame_iters=nat.ptr(iters),
            frame_preserve=nat.ptr(preserve_flags), preserve_all=0,
            out_pose=nat.ptr(out["pose"]), out_betas=nat.ptr(out["betas"]), out_transl=nat.ptr(out["transl"]),
            out_expr=nat.ptr(out["expr"]), out_loss=nat.ptr(loss), out_joints=None, out_evals=nat.ptr(evals),
            workspace=nat.ptr(ws), workspace_bytes=ws.numel())
        events = getattr(self, "kernel_events", None)
        with torch.cuda.device(f.device):
            if events is not None:       # bench.py: per-launch device time of the fit kernel
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            nat.check(lib.k2b_fit_batch(f.native.handle, C.byref(a), nat.current_stream()))
            if events is not None:
                e1.record()
                events.append((e0, e1))

    def run(self, targets: torch.Tensor, seq_ind: torch.Tensor, conf: Optional[torch.Tensor] = None,
            init_pose: Optional[torch.Tensor] = None, init_betas: Optional[torch.Tensor] = None,
            use_lbfgs: Optional[bool] = None, with_mesh: bool = True, group=None,
            params_ready: Optional[torch.cuda.Event] = None) -> dict:
        """Fit ``targets`` (F,K,3) whose frame f is frame ``seq_ind[f]`` of its sequence.

        Sweep-0 initialisation = ``init_pose`` (1,72)|(F,72) (default zeros = the synthetic mean pose),
        ``init_betas`` likewise, and a per-frame root-aligned translation (what
        ``optimize_params_frame`` does for a lone frame, engine.py:89-128).
        ``params_ready`` (optional event) is recorded once the fitted parameters and losses are final, i.e.
        before the mesh pass, so a caller can copy them to the host on another stream while the mesh runs.
        """
        f, F, cfg = self.f, self.F, self.cfg
        dev = f.device
        targets = targets.to(dev, torch.float32)[:, : f.num_obs].contiguous()
        seq_ind = seq_ind.to(dev)
        optimizer = nat.OPT_LBFGS if (f.use_lbfgs if use_lbfgs is None else use_lbfgs) else nat.OPT_ADAM
        if init_pose is not None:
            self.init_pose.copy_(init_pose.to(dev).expand(F, 72))
        if init_betas is not None:
            self.init_betas.copy_(init_betas.to(dev).expand(F, 10))
        # model root at the initial pose/shape: one skeleton pass over the distinct inits
        if self.model_root0 is None or init_pose is not None or init_betas is not None:
            n_distinct = F if (init_pose is not None and init_pose.shape[0] == F) or (
                init_betas is not None and init_betas.shape[0] == F) else 1
            p = {"global_orient": self.init_pose[:n_distinct, :3], "body_pose": self.init_pose[:n_distinct, 3:],
                 "betas": self.init_betas[:n_distinct]}
            self.model_root0 = f.forward_batch(p, with_vertices=False)["joints"][:, 0, :]
        torch.sub(targets[:, 0, :], self.model_root0, out=self.init_transl)

        iters0, iters1, preserve1, starts = plan_two_sweep(seq_ind, f.num_iters_first, f.num_iters_followup)
        s0_rows = {k: (v[1:] if v is not None else None) for k, v in self.s0.items()}
        init0 = dict(pose=self.init_pose, betas=self.init_betas, transl=self.init_transl, expr=self.init_expr)
        self._fit(targets, conf, init0, None, iters0, None, f.num_iters_first, s0_rows, self.loss0, self.evals0,
                  optimizer)
        # halo: left neighbour's last sweep-0 frame -> row 0 (one packed row over NCCL / NVLink)
        last = {"global_orient": self.s0["pose"][F:F + 1, :3], "body_pose": self.s0["pose"][F:F + 1, 3:],
                "betas": self.s0["betas"][F:F + 1], "transl": self.s0["transl"][F:F + 1]}
        if self.s0["expr"] is not None:
            last["expression"] = self.s0["expr"][F:F + 1]
        hev = getattr(self, "halo_events", None)      # bench.py: device time of the exchange
        if hev is not None:
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record()
        halo = exchange_halo(last, group)
        if hev is not None:
            h1.record()
            hev.append((h0, h1))
        if halo is None:
            # no left neighbour: a shard that starts mid-sequence falls back to its own first frame
            for k in ("pose", "betas", "transl", "expr"):
                if self.s0[k] is not None:
                    self.s0[k][0] = self.s0[k][1]
        else:
            self.s0["pose"][0, :3] = halo["global_orient"][0]
            self.s0["pose"][0, 3:] = halo["body_pose"][0]
            self.s0["betas"][0] = halo["betas"][0]
            self.s0["transl"][0] = halo["transl"][0]
            if self.s0["expr"] is not None:
                self.s0["expr"][0] = halo["expression"][0]
        # sweep 1: frame f starts from sweep0[f-1] = row f of the padded buffer (zero-copy view)
        init1 = {k: (v[:F] if v is not None else None) for k, v in self.s0.items()}
        self._fit(targets, conf, init1, None, iters1, preserve1, max(f.num_iters_followup, 1), self.s1, self.loss1,
                  self.evals1, optimizer)
        # sequence starts keep their sweep-0 (seq_ind = 0) fit
        if starts.numel():
            for k in ("pose", "betas", "transl", "expr"):
                if self.s1[k] is not None:
                    self.s1[k].index_copy_(0, starts, s0_rows[k].index_select(0, starts))
            self.loss1.index_copy_(0, starts, self.loss0.index_select(0, starts))
        params = {"global_orient": self.s1["pose"][:, :3], "body_pose": self.s1["pose"][:, 3:],
                  "betas": self.s1["betas"], "transl": self.s1["transl"]}
        if self.s1["expr"] is not None:
            params["expression"] = self.s1["expr"]
        out = {"params": params, "pose": self.s1["pose"], "loss": self.loss1,
               "evals": self.evals0 + self.evals1}
        if params_ready is not None:
            params_ready.record()
        if with_mesh:
            mesh = f.forward_batch(params, with_vertices=self.vertices is not None, out_vertices=self.vertices)
            out["joints"], out["vertices"] = mesh["joints"], mesh["vertices"]
        return out
