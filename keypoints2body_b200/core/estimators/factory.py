"""Estimator factory (reference core/estimators/factory.py:20-44)."""

from __future__ import annotations

from ..config import FrameOptimizeConfig
from .optimization import OptimizationEstimator


def create_estimator(model, frame_config: FrameOptimizeConfig, device, model_type: str):
    kind = frame_config.estimator_type
    if kind == "optimization":
        return OptimizationEstimator(model=model, frame_config=frame_config, device=device, model_type=model_type)
    if kind == "learned":
        raise NotImplementedError(
            "estimator_type='learned' is not implemented yet. "
            "Implement under keypoints2body.core.estimators and wire model loading/inference.")
    if kind == "ikgat":
        raise NotImplementedError(
            "estimator_type='ikgat' (learned IK-GAT regressor) is outside the accelerated path; "
            "its torch_geometric dependency and weights are not available offline")
    raise ValueError(f"Unknown estimator_type: {kind}")
