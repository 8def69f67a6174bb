"""``optimize_params_frame`` -- single-frame entry point.

Signature, argument meaning, error behaviour and result contract follow
/root/reference/keypoints2body/api/frame.py:34-219.  Differences, all deliberate:
results live on the CUDA device; ``device=None`` means the current CUDA device; numpy
``prev_params`` are converted instead of crashing.  World and camera
coordinate modes and dict-block / explicit-index observations are served by the CUDA path; what
still raises ``NotImplementedError`` is listed in DESIGN.md section 7.
"""

from __future__ import annotations

from typing import Optional

import torch

from ..core.config import BodyModelConfig, FrameOptimizeConfig, ModelType
from ..core.engine import (OptimizeEngine, default_init_params, default_init_params_for_model, load_mean_pose_shape,
                           upgrade_smpl_family_init_params)
from ..core.joints.adapters import normalize_frame_observations
from ..models.smpl_data import BodyModelFitResult, BodyModelParams
from ._common import (DEFAULT_MEAN_FILE, MISC_PARAM_TYPES, PARAM_TYPES, SMPL_FAMILY, canonical_layout, check_request,
                      dict_to_params, params_to_dict, resolve_device)
from .model_factory import load_body_model


def optimize_params_frame(
    joints,
    *,
    prev_params: Optional[BodyModelParams] = None,
    body_model: ModelType = "smpl",
    joint_layout: Optional[str] = None,
    model=None,
    config: Optional[FrameOptimizeConfig | dict] = None,
    device=None,
) -> BodyModelFitResult:
    """Fit body-model parameters to one frame of 3D joints ((K,3) or (K,4) with confidences)."""
    if isinstance(config, dict):
        frame_cfg = FrameOptimizeConfig(**config)
    elif isinstance(config, FrameOptimizeConfig):
        frame_cfg = config
    else:
        frame_cfg = FrameOptimizeConfig()
    check_request(frame_cfg, body_model)
    device = resolve_device(device)

    j3d, conf_3d, model_indices, in_layout = normalize_frame_observations(joints, layout=joint_layout,
                                                                          body_model=body_model)
    if body_model in SMPL_FAMILY and in_layout != "GENERIC":
        j3d, conf2d, out_layout = canonical_layout(j3d, conf_3d[None, :], joint_layout, device)
        conf_3d = conf2d[0]
        frame_cfg.joints_category = out_layout      # the reference mutates the caller's config too
    else:
        if joint_layout is not None and in_layout != "GENERIC":
            raise ValueError(
                "joint_layout adapters are currently defined for SMPL-family body "
                "layouts only. Use raw MANO/FLAME joint order with joint_layout=None.")
        frame_cfg.joints_category = "GENERIC"

    if model is None:
        model = load_body_model(BodyModelConfig(model_type=body_model), device)
    engine = OptimizeEngine(model=model, frame_config=frame_cfg, device=device, model_type=body_model)
    fitter = engine.fitter

    if body_model not in SMPL_FAMILY:          # MANO / FLAME (frame.py:144-150, 209-211)
        j3d, conf_3d = j3d.to(device), conf_3d.to(device)
        if prev_params is None:
            init_params = default_init_params_for_model(body_model, model, j3d, device, frame_cfg.coordinate_mode)
        else:
            expected = MISC_PARAM_TYPES[body_model]
            if not isinstance(prev_params, expected):
                raise ValueError(f"prev_params must be {expected.__name__} for body_model={body_model}.")
            init_params = prev_params.to(device)
            if frame_cfg.coordinate_mode == "world" and init_params.transl is None:
                init_params.transl = j3d[:, 0, :].detach()
        return engine.fit_frame(init_params=init_params, j3d=j3d, conf_3d=conf_3d, seq_ind=0,
                                target_model_indices=model_indices)

    if prev_params is None:
        mean_pose, mean_shape = load_mean_pose_shape(DEFAULT_MEAN_FILE, device)
        base = default_init_params(mean_pose, mean_shape, j3d, fitter, joints_category=frame_cfg.joints_category,
                                   coordinate_mode=frame_cfg.coordinate_mode)
        init_params = upgrade_smpl_family_init_params(base, model_type=body_model, model=model, device=device)
    else:
        expected = PARAM_TYPES[body_model]
        if not isinstance(prev_params, expected):
            raise ValueError(f"prev_params must be {expected.__name__} for body_model={body_model}.")
        d = {k: (torch.as_tensor(v, dtype=torch.float32).to(device) if v is not None else None)
             for k, v in params_to_dict(prev_params).items()}
        if d["transl"] is None:
            pose = torch.cat([d["global_orient"], d["body_pose"]], dim=1)
            d["transl"] = default_init_params(pose, d["betas"], j3d, fitter, frame_cfg.joints_category,
                                              frame_cfg.coordinate_mode).transl
        init_params = dict_to_params(body_model, d)

    return engine.fit_frame(init_params=init_params, j3d=j3d, conf_3d=conf_3d, seq_ind=0,
                            target_model_indices=model_indices)
