"""Joint-sequence readers and the per-frame SMPL-X zip writer.

Host-side formats, same files and same semantics as the reference's
``keypoints2body/io/motion.py``:

* ``load_motion_data`` (motion.py:16-57): ``.npy`` array, ``.npz`` with a ``joints`` entry, or the
  mocap ``.csv`` export (5 header lines, 2 leading index columns, then x,y,z triples per joint);
  the result goes through ``adapt_layout`` and a warning reports a changed joint count;
* ``write_smplx_zip`` (motion.py:60-108): one ``frame_%06d/person_%02d.npz`` member per frame with
  the eight SMPL-X keys a renderer expects (21-joint body pose, zero hands / jaw / expression);
* ``write_smplx_zip_from_smpl_data`` (motion.py:111-138).
"""

from __future__ import annotations

import io
import warnings
import zipfile
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from ..core.joints.adapters import adapt_layout
from ..models.smpl_data import SMPLData

_CSV_HEADER_LINES = 5
_CSV_INDEX_COLUMNS = 2


def _read_csv_joints(path: Path) -> np.ndarray:
    frames = []
    with open(path, "r") as fh:
        for n, line in enumerate(fh):
            if n < _CSV_HEADER_LINES:
                continue
            line = line.strip()
            if not line:
                continue
            values = [float(tok) for tok in line.split(",")[_CSV_INDEX_COLUMNS:]]
            usable = len(values) // 3 * 3          # zip() in the reference drops an incomplete trailing triple
            frames.append(np.asarray(values[:usable], dtype=np.float64).reshape(-1, 3))
    return np.array(frames)


def load_motion_data(path, layout: Optional[str] = None) -> Tuple[np.ndarray, str, int]:
    """Load a (T,K,3) joint sequence and normalise its layout -> ``(joints, layout, num_joints)``."""
    path = Path(path)
    ext = path.suffix.lower()
    if ext == ".npy":
        joints = np.load(path)
    elif ext == ".csv":
        joints = _read_csv_joints(path)
    elif ext == ".npz":
        with np.load(path) as data:
            if "joints" not in data:
                raise ValueError(f"Unsupported .npz format: found keys {list(data.keys())}")
            joints = data["joints"]
    else:
        raise ValueError(f"Unsupported 3D joints file format: {ext}")
    k_in = joints.shape[1]
    joints, out_layout = adapt_layout(joints, layout)
    if joints.shape[1] != k_in:
        warnings.warn(f"Converted input joints from {k_in} to {joints.shape[1]} for layout {out_layout}.")
    return joints, out_layout, joints.shape[1]


def write_smplx_zip(output_dir, poses: np.ndarray, betas: np.ndarray, transl: np.ndarray,
                    zip_name: str = "smpl_params.zip", person_idx: int = 0) -> Path:
    """Write (T,72) SMPL poses + shape + translation as a renderer-compatible SMPL-X zip."""
    zip_path = Path(output_dir).expanduser() / zip_name
    if poses.ndim != 2 or poses.shape[1] != 72:
        raise ValueError(f"Expected poses shape (T,72); got {poses.shape}")
    frames = poses.shape[0]
    if betas.ndim == 1:
        betas = np.broadcast_to(betas[None, :], (frames, betas.shape[0]))
    if betas.shape[0] != frames:
        raise ValueError(f"Expected betas shape (T,10); got {betas.shape}")
    if transl.shape[0] != frames:
        raise ValueError(f"Expected transl shape (T,3); got {transl.shape}")
    f32 = np.float32
    constant = {"expression": np.zeros(10, f32), "left_hand_pose": np.zeros(45, f32),
                "right_hand_pose": np.zeros(45, f32), "jaw_pose": np.zeros(3, f32)}
    with zipfile.ZipFile(zip_path, "w", compression=zipfile.ZIP_DEFLATED) as zf:
        for t in range(frames):
            member = {"betas": betas[t].astype(f32), "expression": constant["expression"],
                      "global_orient": poses[t, :3].astype(f32), "body_pose": poses[t, 3:66].astype(f32),
                      "left_hand_pose": constant["left_hand_pose"], "right_hand_pose": constant["right_hand_pose"],
                      "jaw_pose": constant["jaw_pose"], "transl": transl[t].astype(f32)}
            blob = io.BytesIO()
            np.savez(blob, **member)
            zf.writestr(f"frame_{t:06d}/person_{person_idx:02d}.npz", blob.getvalue())
    return zip_path


def _to_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def write_smplx_zip_from_smpl_data(output_dir, smpl_data: SMPLData, zip_name: str = "smpl_params.zip",
                                   person_idx: int = 0) -> Path:
    """Export one ``SMPLData`` sequence (batch-leading tensors or arrays) to the SMPL-X zip format."""
    if smpl_data.transl is None:
        raise ValueError("smpl_data.transl is required to export SMPL-X zip")
    return write_smplx_zip(output_dir, _to_numpy(smpl_data.pose), _to_numpy(smpl_data.betas),
                           _to_numpy(smpl_data.transl), zip_name=zip_name, person_idx=person_idx)
