"""B200-native drop-in for keypoints2body's per-frame SMPL-family fitting path.

Same public surface as the reference package (/root/reference/keypoints2body/__init__.py:3-32);
the fitting loop runs as hand-written sm_100a CUDA kernels behind a C ABI
(include/k2b_b200.h, keypoints2body_b200/libk2b_b200.so).  There is no CPU path.
"""

from .api.frame import optimize_params_frame
from .api.sequence import optimize_params_sequence, optimize_shape_sequence
from .models.smpl_data import (
    BodyModelFitResult,
    BodyModelParams,
    FLAMEData,
    MANOData,
    SMPLData,
    SMPLHData,
    SMPLXData,
)

__version__ = "0.1.0"

__all__ = [
    "__version__",
    "optimize_params_frame",
    "optimize_params_sequence",
    "optimize_shape_sequence",
    "BodyModelFitResult",
    "BodyModelParams",
    "MANOData",
    "FLAMEData",
    "SMPLData",
    "SMPLHData",
    "SMPLXData",
]
