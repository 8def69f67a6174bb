"""Engine wrapper and initialisation helpers (host side, thin).

API-compatible restatement of /root/reference/keypoints2body/core/engine.py
(OptimizeEngine :23-68, load_mean_pose_shape :71-86, default_init_params :89-128,
upgrade_smpl_family_init_params :170-214, optimize_shape_pass :217-262).  Every
numeric step runs on the GPU through the fitter's kernels.
"""

from __future__ import annotations

import importlib.util
import os
from typing import Optional

import numpy as np
import torch

from ..models.smpl_data import (BodyModelFitResult, BodyModelParams, FLAMEData, MANOData, SMPLData, SMPLHData,
                                SMPLXData)
from .config import FrameOptimizeConfig, SequenceOptimizeConfig
from .estimators.factory import create_estimator
from .fitters.world_space import guess_init_transl_from_root


class OptimizeEngine:
    """Routes ``fit_frame`` to the estimator selected by the frame config."""

    def __init__(self, model, frame_config: FrameOptimizeConfig, device, model_type: str = "smpl"):
        self.model = model
        self.frame_config = frame_config
        self.device = device
        self.estimator = create_estimator(model=model, frame_config=frame_config, device=device,
                                          model_type=model_type)

    @property
    def fitter(self):
        return self.estimator.fitter

    def fit_frame(self, init_params: BodyModelParams, j3d: torch.Tensor, conf_3d: Optional[torch.Tensor],
                  seq_ind: int, target_model_indices: Optional[torch.Tensor] = None) -> BodyModelFitResult:
        return self.estimator.fit_frame(init_params=init_params, j3d=j3d, conf_3d=conf_3d, seq_ind=seq_ind,
                                        target_model_indices=target_model_indices)


_MEAN_CACHE: dict = {}


def load_mean_pose_shape(mean_file: str, device) -> tuple[torch.Tensor, torch.Tensor]:
    """Mean pose (1,72) and shape (1,10) from ``neutral_smpl_mean_params.h5``.

    The reference needs h5py (engine.py:83-85).  When h5py is unavailable, or the ``.h5`` is
    absent, an ``.npz`` with the same stem and the same ``pose`` / ``shape`` keys is read.
    """
    stem = os.path.splitext(mean_file)[0]
    use_h5 = os.path.exists(mean_file) and importlib.util.find_spec("h5py") is not None
    src = mean_file if use_h5 else stem + ".npz"
    if not os.path.exists(src):
        raise FileNotFoundError(f"mean-parameter file not found: {mean_file} (or {stem}.npz)")
    # the reference re-reads the file on every call (sequence.py:139-141, frame.py:113-115); here the parsed arrays are
    # kept per (file, modification time, size, device): the zip / HDF5 reader was a fifth of a single-frame call
    st = os.stat(src)
    key = (os.path.abspath(src), st.st_mtime_ns, st.st_size, str(device))
    hit = _MEAN_CACHE.get(key)
    if hit is None:
        if use_h5:
            import h5py  # noqa: WPS433

            with h5py.File(src, "r") as f:
                pose, shape = np.asarray(f["pose"][:]), np.asarray(f["shape"][:])
        else:
            with np.load(src) as f:
                pose, shape = f["pose"], f["shape"]
        to = dict(device=device, dtype=torch.float32)
        hit = (torch.as_tensor(pose).reshape(1, -1).to(**to), torch.as_tensor(shape).reshape(1, -1).to(**to))
        if len(_MEAN_CACHE) > 16:
            _MEAN_CACHE.clear()
        _MEAN_CACHE[key] = hit
    return hit[0].clone(), hit[1].clone()


def default_init_params(mean_pose, mean_shape, joints_frame, fitter, joints_category: str,
                        coordinate_mode: str) -> SMPLData:
    """Mean pose / shape plus a root-aligned translation (engine.py:89-128).

    ``fitter`` is the CUDA-backed fitter (the reference passes the smplx module here; the
    forward it needs is the fitter's mesh kernel).
    """
    pose = mean_pose.clone().detach()
    betas = mean_shape.clone().detach()
    transl = None
    if coordinate_mode == "world":
        transl = guess_init_transl_from_root(fitter, pose, betas, joints_frame,
                                             joints_category="AMASS" if joints_category == "GENERIC" else joints_category)
    return SMPLData(betas=betas, global_orient=pose[:, :3], body_pose=pose[:, 3:], transl=transl)


def default_init_params_for_model(model_type: str, model, joints_frame: torch.Tensor, device, coordinate_mode: str):
    """Zero initialisation for the MANO / FLAME fitters, translation = the first observed joint (engine.py:131-167)."""
    batch = joints_frame.shape[0]
    num_betas = int(getattr(model, "num_betas", 10))
    transl = joints_frame[:, 0, :].clone().detach().to(device) if coordinate_mode == "world" else None
    z = lambda d: torch.zeros((batch, d), device=device)  # noqa: E731
    if model_type == "mano":
        return MANOData(betas=z(num_betas), global_orient=z(3), body_pose=z(0), transl=transl,
                        hand_pose=z(int(getattr(model, "NUM_HAND_JOINTS", 15)) * 3))
    if model_type == "flame":
        return FLAMEData(betas=z(num_betas), global_orient=z(3), body_pose=z(0), transl=transl,
                         expression=z(int(getattr(model, "num_expression_coeffs", 10)) or 10), jaw_pose=z(3),
                         neck_pose=z(3), leye_pose=z(3), reye_pose=z(3))
    raise ValueError(f"default_init_params_for_model is unsupported for {model_type}")


def upgrade_smpl_family_init_params(base_params: SMPLData, model_type: str, model, device) -> BodyModelParams:
    """Zero-initialised hand / face blocks for SMPL-H / SMPL-X (engine.py:170-214)."""
    if model_type == "smpl":
        return base_params
    B = base_params.body_pose.shape[0]
    core = dict(betas=base_params.betas, global_orient=base_params.global_orient,
                body_pose=base_params.body_pose, transl=base_params.transl)
    hand_dim = int(getattr(model, "NUM_HAND_JOINTS", 15)) * 3
    hands = dict(left_hand_pose=torch.zeros((B, hand_dim), device=device),
                 right_hand_pose=torch.zeros((B, hand_dim), device=device))
    if model_type == "smplh":
        return SMPLHData(**core, **hands)
    if model_type == "smplx":
        expr_dim = int(getattr(model, "num_expression_coeffs", 10)) or 10
        zeros3 = lambda: torch.zeros((B, 3), device=device)  # noqa: E731
        return SMPLXData(**core, **hands, expression=torch.zeros((B, expr_dim), device=device),
                         jaw_pose=zeros3(), leye_pose=zeros3(), reye_pose=zeros3())
    raise ValueError(f"Unsupported SMPL-family model_type: {model_type}")


def optimize_shape_pass(fitter, seq_config: SequenceOptimizeConfig, init_mean_shape, init_mean_pose,
                        data_tensor, confidence_input, device):
    """Shared-betas pre-pass over the first frames of a sequence (engine.py:217-262)."""
    if not seq_config.use_shape_optimization:
        return init_mean_shape
    t_size = data_tensor.shape[0]
    n = seq_config.num_shape_frames
    frame_indices = list(range(t_size)) if (n < 0 or n >= t_size) else list(range(n))
    from .shape import optimize_shape_multi_frame

    return optimize_shape_multi_frame(
        fitter, init_betas=init_mean_shape, pose_init=init_mean_pose.repeat(t_size, 1), j3d_world=data_tensor,
        joints_category=seq_config.frame.joints_category, num_iters=seq_config.num_shape_iters, step_size=1e-1,
        use_lbfgs=seq_config.frame.use_lbfgs, device=device, frame_indices=frame_indices,
        joints3d_conf=confidence_input, shape_prior_weight=seq_config.frame.shape_prior_weight)
