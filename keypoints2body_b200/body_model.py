"""Body-model weights as the kernels need them.

The reference treats the body model as an opaque ``smplx`` module and calls it
for every loss evaluation (/root/reference/keypoints2body/core/fitters/world_space.py:174-193).
Here the weight buffers are read ONCE from whatever ``model`` object the caller
passes (by their smplx attribute names [smplx-from-memory]) and uploaded into a
device-resident ``k2b_model`` together with the pose prior.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

_SMPLX_BUFFERS = ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "parents")


def _np(x, dtype=np.float32):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


@dataclass
class BodyModelWeights:
    """Contiguous host arrays in the layout ``k2b_model_desc`` expects."""

    model_type: str
    parents: np.ndarray          # (n_j,) int32
    v_template: np.ndarray       # (V,3)
    shapedirs: np.ndarray        # (V,3,S) with S = 10 (betas) or 20 (betas + expression)
    posedirs: np.ndarray         # (9(n_j-1), 3V)
    J_regressor: np.ndarray      # (n_j,V)
    lbs_weights: np.ndarray      # (V,n_j)
    extra_vertex_ids: np.ndarray # (E,) int32

    @property
    def num_joints(self) -> int:
        return int(self.parents.shape[0])

    @property
    def num_vertices(self) -> int:
        return int(self.v_template.shape[0])

    @property
    def num_shape(self) -> int:
        return int(self.shapedirs.shape[2])

    @property
    def num_extra(self) -> int:
        return int(self.extra_vertex_ids.shape[0])


def infer_model_type(model) -> str:
    mt = getattr(model, "model_type", None)
    if isinstance(mt, str):
        return mt
    n_j = int(_np(model.parents, np.int64).shape[0])
    return {24: "smpl", 52: "smplh", 55: "smplx", 16: "mano", 5: "flame"}.get(n_j, "smpl")


def extract_weights(model, model_type: str | None = None, num_betas: int = 10,
                    num_expression: int = 10) -> BodyModelWeights:
    """Read the smplx-named buffers of ``model`` (an smplx module or any look-alike)."""
    missing = [b for b in _SMPLX_BUFFERS if not hasattr(model, b)]
    if missing:
        raise ValueError(f"model lacks body-model buffers {missing}; pass an smplx-style module")
    model_type = model_type or infer_model_type(model)
    parents = _np(model.parents, np.int64).astype(np.int32)
    parents[0] = -1
    shapedirs = _np(model.shapedirs)
    if shapedirs.ndim != 3:
        raise ValueError("shapedirs must be (V,3,S)")
    if model_type in ("smplx", "flame"):
        expr_dirs = getattr(model, "expr_dirs", None)
        if expr_dirs is not None and shapedirs.shape[2] <= num_betas:
            shapedirs = np.concatenate([shapedirs[..., :num_betas], _np(expr_dirs)[..., :num_expression]], axis=2)
        shapedirs = shapedirs[..., : num_betas + num_expression]
        if shapedirs.shape[2] != num_betas + num_expression:
            raise ValueError("SMPL-X / FLAME need betas + expression shape directions")
    else:
        shapedirs = shapedirs[..., :num_betas]
    if shapedirs.shape[2] not in (10, 20):
        raise ValueError(f"unsupported number of shape coefficients: {shapedirs.shape[2]}")
    posedirs = _np(model.posedirs)
    n_j, n_v = parents.shape[0], _np(model.v_template).shape[0]
    if posedirs.shape != ((n_j - 1) * 9, n_v * 3):
        raise ValueError(f"posedirs must be ({(n_j - 1) * 9},{n_v * 3}), got {posedirs.shape}")
    extra = getattr(model, "extra_vertex_ids", None)
    if extra is None:
        vjs = getattr(model, "vertex_joint_selector", None)
        extra = getattr(vjs, "extra_joints_idxs", None) if vjs is not None else None
    extra = np.zeros((0,), np.int32) if extra is None else _np(extra, np.int64).astype(np.int32)
    J_regressor = model.J_regressor
    if isinstance(J_regressor, torch.Tensor) and J_regressor.is_sparse:
        J_regressor = J_regressor.to_dense()
    return BodyModelWeights(
        model_type=model_type,
        parents=np.ascontiguousarray(parents),
        v_template=_np(model.v_template),
        shapedirs=np.ascontiguousarray(shapedirs),
        posedirs=posedirs,
        J_regressor=_np(J_regressor),
        lbs_weights=_np(model.lbs_weights),
        extra_vertex_ids=np.ascontiguousarray(extra),
    )


def full_pose_from_params(model_type: str, global_orient, body_pose, extras: dict) -> torch.Tensor:
    """Assemble the model's own full-pose vector from the API blocks.

    SMPL: ``[go | body69]``; SMPL-H: ``[go | body63 | lh45 | rh45]``; SMPL-X:
    ``[go | body63 | jaw | leye | reye | lh45 | rh45]`` (the
    reference feeds 69-D body poses to SMPL-H/X; dims 63..68 are prior-only).
    """
    B = global_orient.shape[0]

    def opt(name, dim):
        v = extras.get(name)
        if v is None:
            return torch.zeros(B, dim, dtype=global_orient.dtype, device=global_orient.device)
        return v

    if model_type == "smpl":
        return torch.cat([global_orient, body_pose], dim=1)
    parts = [global_orient, body_pose[:, :63]]
    if model_type == "smplx":
        parts += [opt("jaw_pose", 3), opt("leye_pose", 3), opt("reye_pose", 3)]
    parts += [opt("left_hand_pose", 45), opt("right_hand_pose", 45)]
    return torch.cat(parts, dim=1)
