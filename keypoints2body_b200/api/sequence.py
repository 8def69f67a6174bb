"""``optimize_params_sequence`` / ``optimize_shape_sequence`` -- sequence entry points.

Signatures and result contract follow /root/reference/keypoints2body/api/sequence.py:40-319.
The reference walks the frames serially in Python (:214-281).  Here the per-frame work is
batched onto the GPU according to the schedule:

S1  ``schedule="reference"`` and ``use_previous_frame_init=True`` (the reference default): a
    Gauss-Seidel chain -- frame t starts from frame t-1's result, so it is serial in t by
    construction; ONE launch of the warp-per-sequence kernel (``k2b_fit_chain``) walks the whole
    sequence, the mesh of all frames is produced by one batched pass at the end.
S0  ``schedule="reference"`` and ``use_previous_frame_init=False``: every frame starts from the
    frame-0 initialisation; ONE launch fits all T frames (frame 0: first-frame budget without the
    temporal term; frames t>0: follow-up budget with it) -- exactly the reference's semantics.
S2  ``schedule="two_sweep"``: frame-parallel Jacobi variant of the chain: sweep 0 fits every
    frame independently with first-frame semantics, sweep 1 re-fits frame t>0 as
    ``fit_frame(init=sweep0[t-1], seq_ind=t)``.  Two launches for any T; with several ranks the
    left neighbour's last sweep-0 frame is received over NCCL (``keypoints2body_b200.distributed``).
    This is a documented deviation from S1; each call is still a reference ``fit_frame``.
"""

from __future__ import annotations

import os
from typing import Optional

import torch

from ..core.config import BodyModelConfig, FrameOptimizeConfig, ModelType, SequenceOptimizeConfig
from ..core.constants import FIX_FOOT_CONF, FIX_FOOT_IDX
from ..core.engine import (OptimizeEngine, default_init_params, default_init_params_for_model, load_mean_pose_shape,
                           optimize_shape_pass, upgrade_smpl_family_init_params)
from ..core.joints.adapters import normalize_sequence_observations
from ..models.smpl_data import BodyModelFitResult, BodyModelParams
from ._common import (DEFAULT_MEAN_FILE, MISC_PARAM_TYPES, PARAM_TYPES, SMPL_FAMILY, canonical_layout, check_request,
                      dict_to_params, params_to_dict, resolve_device)
from .model_factory import load_body_model


def _parse_config(config) -> SequenceOptimizeConfig:
    if isinstance(config, dict):
        known = {k: config[k] for k in ("num_shape_iters", "num_shape_frames", "use_shape_optimization",
                                        "use_previous_frame_init", "fix_foot", "limit_frames", "schedule")
                 if k in config}
        return SequenceOptimizeConfig(frame=FrameOptimizeConfig(**config.get("frame", {})), **known)
    if isinstance(config, SequenceOptimizeConfig):
        return config
    return SequenceOptimizeConfig()


def _expand(d: dict, T: int) -> dict:
    return {k: (v.expand(T, -1).contiguous() if v is not None else None) for k, v in d.items()}


def fit_sequence_batched(fitter, xyz, conf, init: dict, seq_cfg: SequenceOptimizeConfig, halo_prev: Optional[dict] = None,
                         first_seq_ind: int = 0):
    """Run schedule S0 / S1 / S2 for one (shard of a) sequence; returns the batched result dict.

    ``init``: dict of (1,dim) tensors = initialisation of (global) frame 0.  ``first_seq_ind``: global
    index of this shard's first frame.  ``halo_prev`` (S2, shards with ``first_seq_ind > 0``): sweep-0
    parameters of the frame just before the shard, as a dict of (1,dim) tensors.
    """
    fc = seq_cfg.frame
    T = xyz.shape[0]
    seq_ind = torch.arange(first_seq_ind, first_seq_ind + T, device=xyz.device)
    kw = dict(joint_loss_weight=fc.joint_loss_weight, pose_preserve_weight=fc.pose_preserve_weight,
              freeze_betas=fc.freeze_betas)
    camera = getattr(fitter, "num_iters", None) is not None and type(fitter).__name__ == "CameraSpaceFitter"
    if camera:
        # camera-space fits are B = 1 by construction in the reference (camera_space.py broadcasts (B,J,3)+(B,3));
        # frames run as a serial loop (chained or from the fixed initialisation) and the mesh, which excludes
        # the camera translation (camera_space.py:301-306), is produced by one batched pass at the end
        if seq_cfg.schedule == "two_sweep":
            raise NotImplementedError("schedule='two_sweep' is defined for the world-space fitter")
        if os.environ.get("K2B_CAMERA_LAUNCH_PER_FRAME", "0") != "1":
            # the whole loop in ONE launch: a warp walks the sequence, both camera stages per frame (camera_sequence)
            return fitter.fit_sequences({k: v for k, v in init.items() if v is not None and k != "transl"}, xyz[None], conf[None],
                                        first_seq_ind=first_seq_ind, chain=seq_cfg.use_previous_frame_init, **kw)
        # diagnostic path: two launches per frame
        prev, rows = init, []
        for t in range(T):
            r = fitter.fit_batch(prev, xyz[t:t + 1], conf[t], seq_ind=first_seq_ind + t, with_mesh=False, **kw)
            rows.append(r)
            if seq_cfg.use_previous_frame_init:
                prev = r["params"]
        params = {k: torch.cat([r["params"][k] for r in rows], dim=0) for k in rows[0]["params"]}
        out = {"params": params, "loss": torch.cat([r["loss"] for r in rows]),
               "evals": torch.cat([r["evals"] for r in rows])}
        out.update(fitter.forward_batch({k: v for k, v in params.items() if k != "transl"}))
        return out
    if seq_cfg.schedule == "two_sweep":
        s0 = fitter.fit_batch(_expand(init, T), xyz, conf, seq_ind=torch.zeros_like(seq_ind), with_mesh=False, **kw)
        p0 = s0["params"]
        if halo_prev is None and first_seq_ind > 0:
            raise ValueError("two_sweep on a shard that does not start the sequence needs halo_prev")
        prev = {}
        for k, v in p0.items():
            head = v[:1] if halo_prev is None else halo_prev[k].to(v.device)
            prev[k] = torch.cat([head, v[:-1]], dim=0).contiguous()
        out = fitter.fit_batch(prev, xyz, conf, seq_ind=seq_ind, **kw)
        if first_seq_ind == 0:      # frame 0 keeps its sweep-0 (seq_ind = 0) fit
            keep0 = fitter.forward_batch({k: v[:1] for k, v in p0.items()})
            for k in out["params"]:
                out["params"][k] = torch.cat([p0[k][:1], out["params"][k][1:]], dim=0)
            out["loss"] = torch.cat([s0["loss"][:1], out["loss"][1:]])
            out["evals"] = torch.cat([s0["evals"][:1], out["evals"][1:]])
            out["joints"] = torch.cat([keep0["joints"], out["joints"][1:]], dim=0)
            out["vertices"] = torch.cat([keep0["vertices"], out["vertices"][1:]], dim=0)
        out["sweep0_last"] = {k: v[-1:].clone() for k, v in p0.items()}
        return out
    if not seq_cfg.use_previous_frame_init:
        return fitter.fit_batch(_expand(init, T), xyz, conf, seq_ind=seq_ind, **kw)
    # S1: the serial chain runs inside ONE launch -- a warp walks the sequence frame by frame
    # (k2b_fit_chain); the mesh of all frames is one batched pass at the end
    if os.environ.get("K2B_S1_LAUNCH_PER_FRAME", "0") != "1":
        return fitter.fit_chain({k: v for k, v in init.items() if v is not None}, xyz[None], conf[None],
                                first_seq_ind=first_seq_ind, chain=True, **kw)
    # diagnostic path: one B=1 launch per frame
    prev, rows = init, []
    for t in range(T):
        r = fitter.fit_batch(prev, xyz[t:t + 1], conf[t], seq_ind=first_seq_ind + t, with_mesh=False, **kw)
        rows.append(r)
        prev = r["params"]
    params = {k: torch.cat([r["params"][k] for r in rows], dim=0) for k in rows[0]["params"]}
    out = {"params": params, "loss": torch.cat([r["loss"] for r in rows]),
           "evals": torch.cat([r["evals"] for r in rows])}
    out.update(fitter.forward_batch(params))
    return out


def _fit_sequence_serial(engine, xyz, conf, prev, model_indices, seq_cfg) -> list[BodyModelFitResult]:
    """The reference's frame loop as written (sequence.py:214-281): one ``fit_frame`` per frame, frame t starting from
    frame t-1's result.  Used where the observations need the general articulated fit (MANO, FLAME, hand / face blocks of
    SMPL-H / SMPL-X): one launch of ``k2b_artic_fit`` and one mesh pass per frame."""
    results = []
    for t in range(xyz.shape[0]):
        frame = xyz[t:t + 1]
        if seq_cfg.frame.coordinate_mode == "world" and prev.transl is None:
            prev.transl = frame[:, 0, :].detach()
        res = engine.fit_frame(init_params=prev, j3d=frame, conf_3d=conf[t], seq_ind=t, target_model_indices=model_indices)
        results.append(res)
        if seq_cfg.use_previous_frame_init:
            prev = res.params
    return results


def optimize_params_sequence(
    joints_seq,
    *,
    init_params: Optional[BodyModelParams] = None,
    body_model: ModelType = "smpl",
    joint_layout: Optional[str] = None,
    model=None,
    config: Optional[SequenceOptimizeConfig | dict] = None,
    device=None,
) -> list[BodyModelFitResult]:
    """Fit a (T,K,3|4) motion sequence; returns per-frame results in temporal order."""
    seq_cfg = _parse_config(config)
    check_request(seq_cfg.frame, body_model)
    device = resolve_device(device)

    xyz, conf, model_indices, in_layout = normalize_sequence_observations(joints_seq, layout=joint_layout,
                                                                          body_model=body_model)
    if body_model in SMPL_FAMILY and in_layout != "GENERIC":
        xyz, conf, out_layout = canonical_layout(xyz, conf, joint_layout, device)
        seq_cfg.frame.joints_category = out_layout
    else:
        if joint_layout is not None and in_layout != "GENERIC":
            raise ValueError(
                "joint_layout adapters are currently defined for SMPL-family body "
                "layouts only. Use raw MANO/FLAME joint order with joint_layout=None.")
        seq_cfg.frame.joints_category = "GENERIC"

    if seq_cfg.limit_frames is not None and seq_cfg.limit_frames > 0:
        xyz, conf = xyz[: seq_cfg.limit_frames], conf[: seq_cfg.limit_frames]
    if seq_cfg.fix_foot and xyz.shape[1] > 11:
        conf = conf.clone()
        conf[:, list(FIX_FOOT_IDX)] = FIX_FOOT_CONF

    if model is None:
        model = load_body_model(BodyModelConfig(model_type=body_model), device)
    engine = OptimizeEngine(model=model, frame_config=seq_cfg.frame, device=device, model_type=body_model)
    fitter = engine.fitter

    if body_model not in SMPL_FAMILY:          # MANO / FLAME (sequence.py:155-158, 192-212)
        xyz, conf = xyz.to(device), conf.to(device)
        if init_params is None:
            prev = default_init_params_for_model(body_model, model, xyz[0:1], device, seq_cfg.frame.coordinate_mode)
        else:
            expected = MISC_PARAM_TYPES[body_model]
            if not isinstance(init_params, expected):
                raise ValueError(f"init_params must be {expected.__name__} for body_model={body_model}.")
            prev = init_params.to(device)
        return _fit_sequence_serial(engine, xyz, conf, prev, model_indices, seq_cfg)

    mean_pose, mean_shape = load_mean_pose_shape(DEFAULT_MEAN_FILE, device)   # always, like sequence.py:139-141
    if seq_cfg.frame.joints_category != "GENERIC":
        betas_opt = optimize_shape_pass(fitter=fitter, seq_config=seq_cfg, init_mean_shape=mean_shape,
                                        init_mean_pose=mean_pose, data_tensor=xyz, confidence_input=conf[0],
                                        device=device)
    else:       # dict-block / explicit-index observations skip the shape pre-pass (sequence.py:142-153)
        betas_opt = mean_shape
    if init_params is None:
        base = default_init_params(mean_pose, betas_opt, xyz[0:1], fitter,
                                   joints_category=seq_cfg.frame.joints_category,
                                   coordinate_mode=seq_cfg.frame.coordinate_mode)
        prev = upgrade_smpl_family_init_params(base, model_type=body_model, model=model, device=device)
    else:
        expected = PARAM_TYPES[body_model]
        if not isinstance(init_params, expected):
            raise ValueError(f"init_params must be {expected.__name__} for body_model={body_model}.")
        prev = init_params
    init = {k: (torch.as_tensor(v, dtype=torch.float32).to(device) if v is not None else None)
            for k, v in params_to_dict(prev).items()}
    if init["transl"] is None:
        pose = torch.cat([init["global_orient"], init["body_pose"]], dim=1)
        init["transl"] = default_init_params(pose, init["betas"], xyz[0:1], fitter, seq_cfg.frame.joints_category,
                                             seq_cfg.frame.coordinate_mode).transl

    if model_indices is not None and not fitter.body_joints_only(model_indices):
        # hand joints / vertex-picked landmarks among the observations: the general articulated fit, frame by frame
        if seq_cfg.frame.coordinate_mode != "world":
            raise NotImplementedError("hand / face observations are fitted in world coordinates (like the reference's "
                                      "CameraSpaceFitter, which takes SMPL body joints only)")
        start = dict_to_params(body_model, init)
        return _fit_sequence_serial(engine, xyz.to(device), conf.to(device), start, model_indices, seq_cfg)
    if model_indices is not None:
        # dict-block observations: into the fitter's observation slots (unobserved joints get confidence 0)
        xyz, conf = fitter.scatter_observations(xyz, conf, model_indices)
    out = fit_sequence_batched(fitter, xyz, conf, init, seq_cfg)
    # per-frame views in one C++ pass per field (split / unbind), not 8 Python slicing calls per frame
    rows = {k: v.split(1) for k, v in out["params"].items() if v is not None}
    verts, joints, losses = out["vertices"].split(1), out["joints"].split(1), out["loss"].unbind(0)
    results = []
    for t in range(xyz.shape[0]):
        results.append(BodyModelFitResult(params=dict_to_params(body_model, {k: r[t] for k, r in rows.items()}),
                                          vertices=verts[t], joints=joints[t], loss=losses[t]))
    return results


def optimize_shape_sequence(joints_seq, *, body_model: ModelType = "smpl", joint_layout: Optional[str] = None,
                            model=None, config: Optional[SequenceOptimizeConfig | dict] = None,
                            device=None) -> BodyModelParams:
    """Run the sequence fit and return the last frame's parameters (sequence.py:286-319)."""
    results = optimize_params_sequence(joints_seq, init_params=None, body_model=body_model,
                                       joint_layout=joint_layout, model=model, config=config, device=device)
    if not results:
        raise ValueError("No frames were optimized")
    return results[-1].params
