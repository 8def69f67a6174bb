"""MPJAE evaluation on the GPU (reference: keypoints2body/cli/eval.py:60-157).

``evaluate_pose_pair`` keeps the reference's signature and return value ``(mean_deg, sum_deg, count)``;
the per-(frame, joint) geodesic angles and their float64 sum come from ``k2b_mpjae`` (eval_kernel.cuh).
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Tuple

import numpy as np
import torch

from . import _native as nat


def load_amass_sequence(npz_path) -> Tuple[np.ndarray, np.ndarray]:
    """AMASS-style npz -> (joints (T,22,3), gt_pose (T, 3 + D)) truncated to the common length (eval.py:60-85)."""
    with np.load(Path(npz_path)) as data:
        missing = [k for k in ("joints", "global_orient", "body_pose") if k not in data]
        if missing:
            raise KeyError(f"Missing keys {missing} in {npz_path}")
        joints = np.asarray(data["joints"], dtype=np.float32)[:, :22, :]
        go = np.atleast_2d(np.asarray(data["global_orient"], dtype=np.float32))
        bp = np.atleast_2d(np.asarray(data["body_pose"], dtype=np.float32))
    gt_pose = np.concatenate([go, bp], axis=1).astype(np.float32)
    frames = min(joints.shape[0], gt_pose.shape[0])
    if frames == 0:
        raise ValueError(f"Empty sequence in {npz_path}")
    return joints[:frames], gt_pose[:frames]


def _device_f32(x, device) -> torch.Tensor:
    return torch.as_tensor(x, dtype=torch.float32).to(device).contiguous()


def angular_error_deg(pred_pose, gt_pose, device=None):
    """Per-(frame, joint) angular error in degrees, on the device -> (angles (n, J) tensor, float64 sum tensor)."""
    lib = nat.load_library()
    device = torch.device(device or "cuda")
    n = min(len(pred_pose), len(gt_pose))
    pred = _device_f32(pred_pose[:n], device)
    gt = _device_f32(gt_pose[:n], device)
    joints = min(pred.shape[1], gt.shape[1]) // 3
    if n == 0 or joints == 0:
        raise ValueError("need at least one frame and one joint")
    angles = torch.empty(n, joints, dtype=torch.float32, device=device)
    total = torch.empty(1, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        nat.check(lib.k2b_mpjae(C.c_void_p(pred.data_ptr()), pred.shape[1], C.c_void_p(gt.data_ptr()), gt.shape[1],
                                n, C.c_void_p(angles.data_ptr()), C.c_void_p(total.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return angles, total


def evaluate_pose_pair(pred_pose, gt_pose, device=None) -> Tuple[float, float, int]:
    """MPJAE of predicted vs ground-truth axis-angle poses -> ``(mean_deg, sum_deg, count)`` (eval.py:142-157)."""
    angles, total = angular_error_deg(pred_pose, gt_pose, device)
    s = float(total.item())
    return s / angles.numel(), s, int(angles.numel())
