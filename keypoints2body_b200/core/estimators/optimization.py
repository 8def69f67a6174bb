"""Optimisation estimator: the CUDA fitter behind the ``BodyEstimator`` protocol.

Mirrors /root/reference/keypoints2body/core/estimators/optimization.py:14-85: it forwards
``joint_loss_weight`` / ``pose_preserve_weight`` / ``freeze_betas`` from the frame config to
the fitter.  An instance can be dropped into the REFERENCE's own ``OptimizeEngine`` (assign it
to ``engine.estimator``) for A/B runs -- see INTEGRATION.md.
"""

from __future__ import annotations

from typing import Optional

import torch

from ...models.smpl_data import BodyModelFitResult, BodyModelParams
from ..config import FrameOptimizeConfig
from ..fitters.camera_space import CameraSpaceFitter
from ..fitters.world_space import WorldSpaceFitter


class OptimizationEstimator:
    def __init__(self, model, frame_config: FrameOptimizeConfig, device, model_type: str = "smpl"):
        self.frame_config = frame_config
        if model_type in ("mano", "flame"):
            from ..fitters.misc_models import FLAMEFitter, MANOFitter

            self._fitter = (MANOFitter if model_type == "mano" else FLAMEFitter)(
                model=model, coordinate_mode=frame_config.coordinate_mode, step_size=frame_config.step_size,
                num_iters_first=frame_config.num_iters_first, num_iters_followup=frame_config.num_iters_followup,
                use_lbfgs=frame_config.use_lbfgs, device=device)
            return
        common = dict(model_type=model_type, prior_folder=getattr(frame_config, "prior_folder", "./data/models/"))
        if frame_config.coordinate_mode == "camera":
            if model_type != "smpl":
                raise NotImplementedError("the camera-space fitter handles SMPLData only (camera_space.py:83)")
            self._fitter = CameraSpaceFitter(
                smpl_model=model, step_size=frame_config.step_size, num_iters=frame_config.num_iters,
                use_lbfgs=frame_config.use_lbfgs, joints_category=frame_config.joints_category, device=device,
                pose_prior_num_gaussians=frame_config.pose_prior_num_gaussians, **common)
            return
        self._fitter = WorldSpaceFitter(
            smpl_model=model, step_size=frame_config.step_size, num_iters_first=frame_config.num_iters_first,
            num_iters_followup=frame_config.num_iters_followup, use_lbfgs=frame_config.use_lbfgs,
            joints_category=frame_config.joints_category, device=device,
            pose_prior_num_gaussians=frame_config.pose_prior_num_gaussians, model_type=model_type,
            prior_folder=getattr(frame_config, "prior_folder", "./data/models/"))

    @property
    def fitter(self) -> WorldSpaceFitter:
        return self._fitter

    def fit_frame(self, init_params: BodyModelParams, j3d: torch.Tensor, conf_3d: Optional[torch.Tensor],
                  seq_ind: int, target_model_indices: Optional[torch.Tensor] = None) -> BodyModelFitResult:
        return self._fitter.fit_frame(
            init_params=init_params, j3d=j3d, conf_3d=conf_3d, seq_ind=seq_ind,
            target_model_indices=target_model_indices, joint_loss_weight=self.frame_config.joint_loss_weight,
            pose_preserve_weight=self.frame_config.pose_preserve_weight,
            freeze_betas=self.frame_config.freeze_betas)
