"""Input handling shared by the frame and sequence entry points."""

from __future__ import annotations

import torch

from ..core.joints.adapters import adapt_layout_and_conf
from ..models.smpl_data import FLAMEData, MANOData, SMPLData, SMPLHData, SMPLXData

OPTIMIZATION_BODY_MODELS = {"smpl", "smplh", "smplx", "mano", "flame"}
SMPL_FAMILY = {"smpl", "smplh", "smplx"}
PARAM_TYPES = {"smpl": SMPLData, "smplh": SMPLHData, "smplx": SMPLXData}
MISC_PARAM_TYPES = {"mano": MANOData, "flame": FLAMEData}
DEFAULT_MEAN_FILE = "./data/models/neutral_smpl_mean_params.h5"


def resolve_device(device):
    """None -> current CUDA device (the reference defaults to CPU, which this build has no path for)."""
    if not torch.cuda.is_available():
        raise RuntimeError("keypoints2body_b200 needs a CUDA device (there is no CPU fallback)")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    if dev.type != "cuda":
        raise ValueError(f"keypoints2body_b200 runs on CUDA devices only, got device={dev}")
    return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())


def check_request(frame_cfg, body_model):
    if frame_cfg.input_type != "joints3d":
        raise NotImplementedError(
            f"input_type='{frame_cfg.input_type}' is not implemented in this release. "
            "Current APIs support only joints3d.")
    if body_model not in OPTIMIZATION_BODY_MODELS:
        raise ValueError(f"Unsupported body_model: {body_model}")


def canonical_layout(xyz, conf, joint_layout, device):
    """(T,K,3),(T,K) tensors in any supported layout -> AMASS-22 / SMPL-24 tensors on ``device``.

    Same numpy round trip as the reference (api/frame.py:82-93, api/sequence.py:96-107).
    """
    xyz_np, conf_np, out_layout = adapt_layout_and_conf(xyz.cpu().numpy(), conf.cpu().numpy(), joint_layout)
    if out_layout not in ("SMPL24", "AMASS"):
        raise ValueError(f"Unsupported output layout after adaptation: {out_layout}")
    return (torch.as_tensor(xyz_np, dtype=torch.float32, device=device),
            torch.as_tensor(conf_np, dtype=torch.float32, device=device), out_layout)


def params_to_dict(p) -> dict:
    out = {k: getattr(p, k) for k in ("global_orient", "body_pose", "betas", "transl")}
    for k in ("left_hand_pose", "right_hand_pose", "expression", "jaw_pose", "leye_pose", "reye_pose"):
        if hasattr(p, k):
            out[k] = getattr(p, k)
    return out


def dict_to_params(body_model: str, d: dict, index=None):
    """Build SMPLData / SMPLHData / SMPLXData from a dict of (B,dim) tensors (optionally one row)."""
    def pick(k):
        v = d.get(k)
        if v is None:
            return None
        return v if index is None else v[index:index + 1]

    core = dict(betas=pick("betas"), global_orient=pick("global_orient"), body_pose=pick("body_pose"),
                transl=pick("transl"))
    if body_model == "smpl":
        return SMPLData(**core)
    hands = dict(left_hand_pose=pick("left_hand_pose"), right_hand_pose=pick("right_hand_pose"))
    if body_model == "smplh":
        return SMPLHData(**core, **hands)
    return SMPLXData(**core, **hands, expression=pick("expression"), jaw_pose=pick("jaw_pose"),
                     leye_pose=pick("leye_pose"), reye_pose=pick("reye_pose"))
