"""Parameter / result containers of the fitting API.

Field names, inheritance and the ``pose`` / ``to`` / ``detach`` / ``validate``
behaviour follow the reference contract
(/root/reference/keypoints2body/models/smpl_data.py:33-120) because callers
construct and read these objects directly.  numpy members are passed through
untouched by ``to``/``detach`` exactly as the reference does (:19-30).
"""

from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Any, Optional, Union

import numpy as np
import torch

ArrayLike = Union[np.ndarray, torch.Tensor]

_CORE_FIELDS = ("betas", "global_orient", "body_pose", "transl")


def _map_core(obj, fn):
    """Rebuild ``obj`` with ``fn`` applied to its four core arrays (None kept)."""
    changes = {}
    for name in _CORE_FIELDS:
        value = getattr(obj, name)
        changes[name] = None if value is None else fn(value)
    return dataclasses.replace(obj, **changes)


@dataclass
class BodyModelParams:
    """betas / global_orient / body_pose (+ optional transl) of one or more frames."""

    betas: ArrayLike
    global_orient: ArrayLike
    body_pose: ArrayLike
    transl: Optional[ArrayLike] = None
    metadata: dict[str, Any] = field(default_factory=dict)

    @property
    def pose(self) -> ArrayLike:
        """``[global_orient | body_pose]`` along the last axis."""
        if isinstance(self.global_orient, torch.Tensor):
            return torch.cat((self.global_orient, self.body_pose), dim=-1)
        return np.concatenate((self.global_orient, self.body_pose), axis=-1)

    def validate(self) -> None:
        for name in _CORE_FIELDS[:3]:
            if getattr(self, name) is None:
                raise ValueError("betas, global_orient, and body_pose are required")

    def to(self, device) -> "BodyModelParams":
        def move(x):
            if isinstance(x, torch.Tensor) and device is not None:
                return x.to(device=device)
            return x

        return _map_core(self, move)

    def detach(self) -> "BodyModelParams":
        return _map_core(
            self, lambda x: x.detach() if isinstance(x, torch.Tensor) else x
        )


@dataclass
class SMPLData(BodyModelParams):
    """SMPL: nothing beyond the base fields."""


@dataclass
class SMPLHData(SMPLData):
    """SMPL-H: adds two 45-D axis-angle hand poses."""

    left_hand_pose: Optional[ArrayLike] = None
    right_hand_pose: Optional[ArrayLike] = None


@dataclass
class SMPLXData(SMPLHData):
    """SMPL-X: adds expression, jaw and eye poses."""

    expression: Optional[ArrayLike] = None
    jaw_pose: Optional[ArrayLike] = None
    leye_pose: Optional[ArrayLike] = None
    reye_pose: Optional[ArrayLike] = None


@dataclass
class MANOData(BodyModelParams):
    """MANO hand model parameters (container only; fitter is out of scope)."""

    hand_pose: Optional[ArrayLike] = None


@dataclass
class FLAMEData(BodyModelParams):
    """FLAME head model parameters (container only; fitter is out of scope)."""

    expression: Optional[ArrayLike] = None
    jaw_pose: Optional[ArrayLike] = None
    neck_pose: Optional[ArrayLike] = None
    leye_pose: Optional[ArrayLike] = None
    reye_pose: Optional[ArrayLike] = None


@dataclass
class BodyModelFitResult:
    """What every ``fit_frame`` returns: params, mesh vertices, joints, loss."""

    params: BodyModelParams
    vertices: torch.Tensor
    joints: torch.Tensor
    loss: Optional[torch.Tensor] = None
