"""Configuration objects accepted by the public API.

Same field names and defaults as the reference
(/root/reference/keypoints2body/core/config.py:10-59) so that existing config
dicts / objects keep working.  Two fields are additions of this build and are
ignored by the reference: ``FrameOptimizeConfig.prior_folder`` (where
``gmm_XX.pkl`` lives; the reference hard-codes ``./data/models/``,
world_space.py:88) and ``SequenceOptimizeConfig.schedule`` (see
``api/sequence.py``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Literal, Optional

ModelType = Literal["smpl", "smplh", "smplx", "mano", "flame"]


@dataclass
class BodyModelConfig:
    """How ``load_body_model`` finds a model file (reference config.py:10-19)."""

    model_type: ModelType = "smpl"
    model_family: str = "smpl_family"
    gender: str = "neutral"
    ext: Optional[str] = None
    batch_size: int = 1
    model_dir: Path = Path("./data/models/")


@dataclass
class FrameOptimizeConfig:
    """Per-frame fit options (reference config.py:22-46)."""

    estimator_type: Literal["optimization", "learned", "ikgat"] = "optimization"
    input_type: Literal["joints3d", "joints2d", "multiview_joints2d"] = "joints3d"
    coordinate_mode: Literal["camera", "world"] = "world"
    use_lbfgs: bool = True
    step_size: float = 1e-2
    num_iters: int = 100  # camera-space fitter only
    num_iters_first: int = 30
    num_iters_followup: int = 10
    joint_loss_weight: float = 600.0
    pose_preserve_weight: float = 5.0
    freeze_betas: bool = False
    shape_prior_weight: float = 5.0  # shape pass only; the frame loss uses 5.0 (losses.py:35)
    pose_prior_num_gaussians: int = 8
    joints_category: Literal["SMPL24", "AMASS", "GENERIC"] = "AMASS"
    # IK-GAT estimator knobs: accepted for config compatibility, unused here.
    ikgat_model_dir: Path = Path("./data/estimators")
    ikgat_model_format: str = "manny"
    ikgat_model_type: str = "pos_to_rot6"
    ikgat_parent_ids: Optional[list[int]] = None
    ikgat_hidden_dim: int = 128
    ikgat_num_layers: int = 3
    ikgat_num_heads: int = 4
    # --- additions of this build -------------------------------------------
    prior_folder: str = "./data/models/"


@dataclass
class SequenceOptimizeConfig:
    """Sequence-level options (reference config.py:49-59)."""

    frame: FrameOptimizeConfig = field(default_factory=FrameOptimizeConfig)
    num_shape_iters: int = 40
    num_shape_frames: int = 50
    use_shape_optimization: bool = True
    use_previous_frame_init: bool = True
    fix_foot: bool = False
    limit_frames: Optional[int] = None
    # --- addition of this build --------------------------------------------
    # "reference": honour use_previous_frame_init exactly (S1 chain when True,
    #              S0 independent frames when False);
    # "two_sweep": frame-parallel Jacobi schedule S2.
    schedule: Literal["reference", "two_sweep"] = "reference"
