"""Camera-space two-stage fitter on the fused CUDA kernel.

Drop-in for the reference's ``CameraSpaceFitter``
(/root/reference/keypoints2body/core/fitters/camera_space.py:44-339):

stage 0  forward at the initial parameters; camera translation initialised as the mean offset of
         RHip / LHip / RShoulder / LShoulder (``guess_init_3d``, camera_space.py:16-41);
stage 1  ``num_iters`` iterations over ``[global_orient, camera_translation]`` on
         ``camera_fitting_loss_3d`` (core/losses.py:70-93) -- kernel ``loss_kind = 1``;
stage 2  ``num_iters`` iterations over ``[body_pose, (betas), global_orient, camera_translation]`` on
         ``body_fitting_loss_3d(camera_translation=...)`` -- the world-space kernel with the translation
         playing the camera translation; betas move only if ``seq_ind == 0 or not freeze_betas``
         (camera_space.py:219-224); the returned loss is re-evaluated without the preserve term
         (camera_space.py:316-326) -- kernel ``final_loss_mode = 1``.

Returned vertices / joints do NOT include the camera translation (camera_space.py:301-306).
"""

from __future__ import annotations

from typing import Optional

import torch

from ... import _native as nat
from ...models.smpl_data import BodyModelFitResult, SMPLData
from .world_space import WorldSpaceFitter, _f32

_TORSO = (2, 1, 17, 16)   # RHip, LHip, RShoulder, LShoulder (same indices in the SMPL24 and AMASS maps)


def guess_init_3d(model_joints: torch.Tensor, j3d: torch.Tensor, joints_category: str = "SMPL24") -> torch.Tensor:
    """Initial camera translation from the four torso joints (camera_space.py:16-41)."""
    if joints_category not in ("SMPL24", "AMASS"):
        raise ValueError(f"Unknown joints category: {joints_category}")
    idx = list(_TORSO)
    return (j3d[:, idx] - model_joints[:, idx]).sum(dim=1) / 4.0


class CameraSpaceFitter(WorldSpaceFitter):
    """Per-frame optimiser in camera coordinates (two kernel launches + one mesh pass per batch)."""

    def __init__(self, smpl_model, step_size=1e-2, num_iters=100, use_lbfgs=True, joints_category="SMPL24",
                 device=None, pose_prior_num_gaussians=8, **kw):
        super().__init__(smpl_model, step_size=step_size, num_iters_first=num_iters, num_iters_followup=num_iters,
                         use_lbfgs=use_lbfgs, joints_category=joints_category, device=device,
                         pose_prior_num_gaussians=pose_prior_num_gaussians, **kw)
        self.num_iters = int(num_iters)
        if self.has_expr:
            raise NotImplementedError("the camera-space fitter takes SMPLData (camera_space.py:83); SMPL-X is world-space only")

    def fit_batch(self, init: dict, j3d, conf=None, *, seq_ind: int = 0, joint_loss_weight=600.0,
                  pose_preserve_weight=5.0, freeze_betas=True, init_cam_t=None, with_mesh=True, **_):
        dev = self.device
        go, bp = _f32(init["global_orient"], dev), _f32(init["body_pose"], dev)
        betas = _f32(init["betas"], dev)
        B = go.shape[0]
        if betas.shape[0] != B:
            betas = betas.expand(B, -1).contiguous()
        targets = _f32(j3d, dev)[:, : self.num_obs].contiguous()
        conf = _f32(conf, dev)
        conf_pf = conf is not None and conf.dim() == 2
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
        if init_cam_t is None:
            joints0 = self.forward_batch({"global_orient": go, "body_pose": bp, "betas": betas},
                                         with_vertices=False)["joints"]
            init_cam_t = guess_init_3d(joints0, targets, self.joints_category)
        cam_t0 = _f32(init_cam_t, dev).contiguous()
        opt = nat.OPT_LBFGS if self.use_lbfgs else nat.OPT_ADAM
        pose = torch.cat([go, bp], dim=1).contiguous()
        move_betas = seq_ind == 0 or not freeze_betas
        if B <= self.warp_kernel_max_frames:
            # few frames (the reference's camera fits are B = 1): one warp per frame (k2b_fit_chain with one-frame
            # sequences) instead of one thread per frame
            tg = targets.reshape(B, 1, self.num_obs, 3)
            cm = 0 if conf is None else (2 if conf_pf else 1)
            cf = conf.reshape(B, 1, self.num_obs) if conf_pf else conf
            s1 = self._run_chain(B, 1, tg, cf, cm, pose, betas, cam_t0, None, None, 0, True, self.num_iters,
                                 self.num_iters, opt, joint_loss_weight, 0.0, True, want_joints=False, loss_kind=1,
                                 depth_ref=cam_t0)
            s2 = self._run_chain(B, 1, tg, cf, cm, s1["pose"], betas, s1["transl"], None, bp.contiguous(),
                                 int(seq_ind > 0), True, self.num_iters, self.num_iters, opt, joint_loss_weight,
                                 pose_preserve_weight, not move_betas, final_loss_mode=1)
        else:
            s1, s2 = self._fit_two_stage_frames(B, targets, conf, conf_pf, pose, betas, cam_t0, bp, seq_ind, opt,
                                                joint_loss_weight, pose_preserve_weight, move_betas)
        params = {"global_orient": s2["pose"][:, :3], "body_pose": s2["pose"][:, 3:], "betas": s2["betas"],
                  "transl": s2["transl"]}
        out = {"params": params, "loss": s2["loss"], "evals": s1["evals"] + s2["evals"], "fit_joints": s2["fit_joints"]}
        if with_mesh:
            out.update(self.forward_batch({k: v for k, v in params.items() if k != "transl"}))
        return out

    def fit_sequences(self, init: dict, j3d, conf=None, *, first_seq_ind: int = 0, chain: bool = True,
                      joint_loss_weight=600.0, pose_preserve_weight=5.0, freeze_betas=True, with_mesh=True):
        """S camera-space sequences of T frames in ONE launch (``k2b_fit_chain`` with ``camera_sequence = 1``): every frame
        is a whole ``fit_frame`` -- camera translation from the torso joints at the frame's initial parameters, stage 1,
        stage 2 -- and with ``chain`` frame t starts from frame t-1's result, which is the reference's loop
        (api/sequence.py:214-281 with ``coordinate_mode="camera"``).  ``init``: (S,dim) blocks; ``j3d`` (S,T,K,3);
        ``conf`` None | (K,) | (S,T,K).  Returns the batched result dict with rows ordered [sequence][frame]."""
        dev = self.device
        targets = _f32(j3d, dev)
        if targets.dim() != 4:
            raise ValueError(f"j3d must be (S, T, K, 3), got {tuple(targets.shape)}")
        S, T = targets.shape[0], targets.shape[1]
        targets = targets[:, :, : self.num_obs].contiguous()
        go, bp = _f32(init["global_orient"], dev), _f32(init["body_pose"], dev)
        betas = _f32(init["betas"], dev)
        if betas.shape[0] != S:
            betas = betas.expand(S, -1)
        conf = _f32(conf, dev)
        cm = 0 if conf is None else (2 if conf.dim() == 3 else 1)
        if conf is not None:
            conf = conf[..., : self.num_obs].contiguous()
        pose = torch.cat([go, bp], dim=1).expand(S, -1).contiguous()
        opt = nat.OPT_LBFGS if self.use_lbfgs else nat.OPT_ADAM
        s2 = self._run_chain(S, T, targets, conf, cm, pose, betas.contiguous(), torch.zeros(S, 3, device=dev), None, None,
                             int(first_seq_ind), chain, self.num_iters, self.num_iters, opt, joint_loss_weight,
                             pose_preserve_weight, bool(freeze_betas), camera_sequence=True)
        params = {"global_orient": s2["pose"][:, :3], "body_pose": s2["pose"][:, 3:], "betas": s2["betas"],
                  "transl": s2["transl"]}
        out = {"params": params, "loss": s2["loss"], "evals": s2["evals"], "fit_joints": s2["fit_joints"]}
        if with_mesh:
            out.update(self.forward_batch({k: v for k, v in params.items() if k != "transl"}))
        return out

    def _fit_two_stage_frames(self, B, targets, conf, conf_pf, pose, betas, cam_t0, bp, seq_ind, opt, joint_loss_weight,
                              pose_preserve_weight, move_betas):
        """Both stages on the one-thread-per-frame kernel (large batches)."""
        # stage 1: orientation + camera translation against the torso joints
        s1 = self._run_fit(B, targets, conf, conf_pf, pose, betas, cam_t0, None, None, None, None, 0, self.num_iters,
                           opt, joint_loss_weight, 0.0, True, want_joints=False, loss_kind=1, depth_ref=cam_t0)
        # stage 2: full body fit; the temporal anchor is the INITIAL body pose (camera_space.py:136)
        s2 = self._run_fit(B, targets, conf, conf_pf, s1["pose"], betas, s1["transl"], None, bp.contiguous(), None, None,
                           int(seq_ind > 0), self.num_iters, opt, joint_loss_weight, pose_preserve_weight,
                           not move_betas, want_joints=True, final_loss_mode=1)
        return s1, s2

    def fit_frame(self, init_params: SMPLData, j3d: torch.Tensor, conf_3d: Optional[torch.Tensor] = None,
                  seq_ind: int = 0, target_model_indices: Optional[torch.Tensor] = None,
                  joint_loss_weight: float = 600.0, pose_preserve_weight: float = 5.0, freeze_betas: bool = True,
                  init_cam_t: Optional[torch.Tensor] = None) -> BodyModelFitResult:
        if target_model_indices is not None:
            raise NotImplementedError("explicit target_model_indices (GENERIC observations) are not built yet")
        init = {k: getattr(init_params, k) for k in ("global_orient", "body_pose", "betas")}
        if conf_3d is not None:
            conf_3d = _f32(conf_3d, self.device)
            if conf_3d.dim() == 2:
                conf_3d = conf_3d[0]
        out = self.fit_batch(init, j3d, conf_3d, seq_ind=int(seq_ind), joint_loss_weight=joint_loss_weight,
                             pose_preserve_weight=pose_preserve_weight, freeze_betas=freeze_betas,
                             init_cam_t=init_cam_t)
        p = out["params"]
        fitted = SMPLData(betas=p["betas"], global_orient=p["global_orient"], body_pose=p["body_pose"],
                          transl=p["transl"])
        return BodyModelFitResult(params=fitted, vertices=out["vertices"], joints=out["joints"], loss=out["loss"].sum())
