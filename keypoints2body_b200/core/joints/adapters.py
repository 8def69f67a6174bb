"""Host-side normalisation of keypoint inputs.

Semantics restate /root/reference/keypoints2body/core/joints/adapters.py
(:35-99 layout registry, :102-117 resolution, :138-199 array normalisation,
:202-221 layout adaptation, :224-380 dict-block observations).  Everything here
is cheap index / rotate work that stays on the host; the kernels only ever see
the canonical AMASS-22 / SMPL-24 layout (or explicit model indices).
"""

from __future__ import annotations

from typing import NamedTuple, Optional, Sequence

import numpy as np
import torch

from ..constants import (
    AMASS_SMPL_IDX,
    SMPL_IDX,
    SMPLX_FACE_IDX_START,
    SMPLX_LEFT_HAND_IDX,
    SMPLX_RIGHT_HAND_IDX,
)


class JointLayoutAdapter(NamedTuple):
    """One supported input layout: size, canonical target, optional gather/rotation."""

    name: str
    expected_joints: int
    out_layout: str
    mapping: Optional[Sequence[int]] = None
    rotate_osim_to_smpl: bool = False


# Source-joint gather tables into AMASS-22 order (reference adapters.py:38-98).
_MANNY25_TO_AMASS = (0, 21, 17, 1, 22, 18, 3, 23, 19, 5, 24, 20, 6, 13, 9, 8, 14, 10,
                     15, 11, 16, 12)
_SPINETRACK37_TO_AMASS = (0, 31, 25, 1, 32, 26, 3, 33, 27, 5, 34, 28, 6, 21, 17, 8,
                          22, 18, 23, 19, 24, 20)

ADAPTERS = {
    ad.name: ad
    for ad in (
        JointLayoutAdapter("SMPL24", 24, "SMPL24"),
        JointLayoutAdapter("AMASS", 22, "AMASS"),
        JointLayoutAdapter("Manny25", 25, "AMASS", _MANNY25_TO_AMASS, True),
        # placeholder truncation, as in the reference (adapters.py:68)
        JointLayoutAdapter("Halpe26", 26, "AMASS", tuple(range(22)), False),
        JointLayoutAdapter("SpineTrack37", 37, "AMASS", _SPINETRACK37_TO_AMASS, True),
    )
}

# OpenSim (x fwd, y up, z right) -> SMPL axes: p_smpl = ROT @ p_osim.
_OSIM_TO_SMPL = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]])


def resolve_adapter(joint_count: int, layout: Optional[str]) -> JointLayoutAdapter:
    """Pick the adapter named ``layout`` or, if None, the one with ``joint_count`` joints."""
    if layout is None:
        for ad in ADAPTERS.values():
            if ad.expected_joints == joint_count:
                return ad
        raise ValueError(f"Unsupported number of joints: {joint_count}")
    ad = ADAPTERS.get(layout)
    if ad is None:
        raise ValueError(f"Unsupported layout: {layout}")
    if ad.expected_joints != joint_count:
        raise ValueError(
            f"Layout {layout} expects {ad.expected_joints} joints, got {joint_count}"
        )
    return ad


def _gather_points(seq: np.ndarray, ad: JointLayoutAdapter) -> np.ndarray:
    if ad.mapping is not None:
        seq = seq[:, list(ad.mapping), :]
    if ad.rotate_osim_to_smpl:
        seq = seq @ _OSIM_TO_SMPL.astype(seq.dtype).T
    return seq


def adapt_layout(joints_seq: np.ndarray, layout: Optional[str]):
    """(T,K,3) in any supported layout -> (T,22|24,3) canonical + its layout label."""
    ad = resolve_adapter(joints_seq.shape[1], layout)
    return _gather_points(joints_seq, ad), ad.out_layout


def adapt_layout_and_conf(joints_seq: np.ndarray, conf_seq: np.ndarray, layout: Optional[str]):
    """Like :func:`adapt_layout` but also gathers the (T,K) confidences."""
    ad = resolve_adapter(joints_seq.shape[1], layout)
    conf = conf_seq if ad.mapping is None else conf_seq[:, list(ad.mapping)]
    return _gather_points(joints_seq, ad), conf, ad.out_layout


def _as_float_tensor(x, what: str) -> torch.Tensor:
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x, dtype=torch.float32)
    if isinstance(x, torch.Tensor):
        return x.float()
    raise ValueError(f"{what} must be numpy array or torch tensor")


def normalize_joints_frame(joints):
    """(K,3|4) -> xyz (1,K,3), conf (K,); a 4th column is the confidence."""
    jt = _as_float_tensor(joints, "joints")
    if jt.ndim != 2 or jt.shape[1] not in (3, 4):
        raise ValueError(f"Expected joints shape (K,3) or (K,4), got {tuple(jt.shape)}")
    if jt.shape[1] == 4:
        return jt[:, :3].unsqueeze(0), jt[:, 3].clone()
    return jt.unsqueeze(0), torch.ones(jt.shape[0], dtype=jt.dtype, device=jt.device)


def normalize_joints_sequence(joints_seq):
    """(T,K,3|4) -> xyz (T,K,3), conf (T,K)."""
    jt = _as_float_tensor(joints_seq, "joints_seq")
    if jt.ndim != 3 or jt.shape[2] not in (3, 4):
        raise ValueError(
            f"Expected joints_seq shape (T,K,3) or (T,K,4), got {tuple(jt.shape)}"
        )
    if jt.shape[2] == 4:
        return jt[..., :3], jt[..., 3]
    return jt, torch.ones(jt.shape[:2], dtype=jt.dtype, device=jt.device)


_BLOCK_ORDER = ("body", "left_hand", "right_hand", "face")
_NEED_DICT_MSG = "dict input must provide at least one of: body, left_hand, right_hand, face"


def _block_model_indices(key: str, k: int, body_model: str) -> torch.Tensor:
    """Model-joint indices a dict block is fitted against (reference adapters.py:251-294)."""
    if key == "body":
        if k == 24:
            return torch.tensor(list(SMPL_IDX), dtype=torch.long)
        if k == 22:
            return torch.tensor(list(AMASS_SMPL_IDX), dtype=torch.long)
        raise ValueError("body block must have 22 or 24 joints")
    if key in ("left_hand", "right_hand"):
        if body_model not in {"smplh", "smplx"}:
            raise ValueError(f"{key} block requires body_model='smplh' or 'smplx'")
        if k != 21:
            raise ValueError(f"{key} block must have 21 joints")
        rng = SMPLX_LEFT_HAND_IDX if key == "left_hand" else SMPLX_RIGHT_HAND_IDX
        return torch.tensor(list(rng), dtype=torch.long)
    if body_model != "smplx":
        raise ValueError("face block requires body_model='smplx'")
    return torch.arange(SMPLX_FACE_IDX_START, SMPLX_FACE_IDX_START + k, dtype=torch.long)


def normalize_frame_observations(joints, *, layout: Optional[str], body_model: str):
    """One frame -> (j3d (1,K,3), conf (K,), model_indices | None, "AUTO" | "GENERIC")."""
    if not isinstance(joints, dict):
        j3d, conf = normalize_joints_frame(joints)
        return j3d, conf, None, "AUTO"
    pts, confs, idxs = [], [], []
    for key in _BLOCK_ORDER:
        if key not in joints:
            continue
        p, c = normalize_joints_frame(joints[key])
        idxs.append(_block_model_indices(key, p.shape[1], body_model))
        pts.append(p)
        confs.append(c)
    if not pts:
        raise ValueError(_NEED_DICT_MSG)
    return torch.cat(pts, dim=1), torch.cat(confs, dim=0), torch.cat(idxs, dim=0), "GENERIC"


def normalize_sequence_observations(joints_seq, *, layout: Optional[str], body_model: str):
    """A sequence -> (xyz (T,K,3), conf (T,K), model_indices | None, "AUTO" | "GENERIC")."""
    if not isinstance(joints_seq, dict):
        xyz, conf = normalize_joints_sequence(joints_seq)
        return xyz, conf, None, "AUTO"
    pts, confs, idxs = [], [], []
    for key in _BLOCK_ORDER:
        if key not in joints_seq:
            continue
        p, c = normalize_joints_sequence(joints_seq[key])
        if pts and p.shape[0] != pts[0].shape[0]:
            raise ValueError("all dict sequence blocks must share same T")
        idxs.append(_block_model_indices(key, p.shape[1], body_model))
        pts.append(p)
        confs.append(c)
    if not pts:
        raise ValueError(_NEED_DICT_MSG)
    return torch.cat(pts, dim=1), torch.cat(confs, dim=1), torch.cat(idxs, dim=0), "GENERIC"
