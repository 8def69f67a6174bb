"""MANO / FLAME fitters backed by the general articulated CUDA fit (``k2b_artic_fit``).

Drop-ins for the reference's ``MANOFitter`` / ``FLAMEFitter``
(/root/reference/keypoints2body/core/fitters/misc_models.py:18-359): same constructor arguments, ``fit_frame``
signature and result contract.  Loss = ``generic_keypoint_loss_3d`` (core/losses.py:96-112) + the fitter's L2
regularisers + the temporal term for ``seq_ind > 0``; the model's weights are read once from the module passed as
``model`` (smplx buffer names).  ``coordinate_mode="camera"`` fits the same objective (the translation is added to the
joints instead of being passed to the model) and returns vertices / joints without the translation, like the reference.
"""

from __future__ import annotations

from typing import Optional

import torch

from ... import _native as nat
from ...body_model import BodyModelWeights, extract_weights
from ...models.smpl_data import BodyModelFitResult, FLAMEData, MANOData
from .articulated import ArticulatedModel, _f32, articulated_fit, mesh_forward


class _ArticulatedFitter:
    model_type = ""

    def __init__(self, model, coordinate_mode: str, step_size=1e-2, num_iters_first=30, num_iters_followup=10,
                 use_lbfgs=True, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("keypoints2body_b200 needs a CUDA device (there is no CPU fallback)")
        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if device.type != "cuda":
            raise ValueError(f"keypoints2body_b200 runs on CUDA devices only, got device={device}")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.model, self.coordinate_mode, self.device = model, coordinate_mode, device
        self.step_size = float(step_size)
        self.num_iters_first, self.num_iters_followup = int(num_iters_first), int(num_iters_followup)
        self.use_lbfgs = bool(use_lbfgs)
        cached = getattr(model, "_k2b_artic", None)
        if cached is not None and cached[0] == str(device):
            self.native, self.artic = cached[1], cached[2]
        else:
            weights = model if isinstance(model, BodyModelWeights) else extract_weights(model, self.model_type)
            self.native = nat.NativeModel(weights, None, device)           # mesh-only: no SMPL pose prior
            self.artic = ArticulatedModel(weights, self.native, with_body_priors=False)
            try:
                model._k2b_artic = (str(device), self.native, self.artic)
            except Exception:
                pass

    def _fit(self, blocks: dict, j3d, conf_3d, seq_ind, target_model_indices, joint_loss_weight, pose_preserve_weight,
             freeze_betas):
        dev = self.device
        j3d = _f32(j3d, dev)
        if conf_3d is not None:
            conf_3d = _f32(conf_3d, dev)
            if conf_3d.dim() == 2:
                conf_3d = conf_3d[0]                    # reference quirk (misc_models.py:59-60, 207-208)
        idx = (torch.arange(j3d.shape[1]) if target_model_indices is None
               else torch.as_tensor(target_model_indices).reshape(-1))
        if conf_3d is not None:
            conf_3d = conf_3d[: idx.numel()].contiguous()
        if blocks.get("transl") is None:
            blocks["transl"] = j3d[:, 0, :].clone()
        iters = self.num_iters_first if seq_ind == 0 else self.num_iters_followup
        return articulated_fit(self.artic, blocks, j3d, conf_3d, idx, seq_ind=int(seq_ind), num_iters=iters,
                               use_lbfgs=self.use_lbfgs, lr=self.step_size, joint_loss_weight=joint_loss_weight,
                               pose_preserve_weight=pose_preserve_weight, freeze_betas=freeze_betas)

    def _mesh(self, full_pose, shape, transl):
        return mesh_forward(self.native, full_pose, shape, transl if self.coordinate_mode == "world" else None)


class MANOFitter(_ArticulatedFitter):
    """Optimisation fitter for the MANO hand model (misc_models.py:18-167)."""

    model_type = "mano"

    def fit_frame(self, init_params: MANOData, j3d: torch.Tensor, conf_3d: Optional[torch.Tensor] = None, seq_ind: int = 0,
                  target_model_indices: Optional[torch.Tensor] = None, joint_loss_weight: float = 600.0,
                  pose_preserve_weight: float = 5.0, freeze_betas: bool = False) -> BodyModelFitResult:
        if init_params.hand_pose is None:
            raise ValueError("MANOData.hand_pose is required")
        blocks = dict(global_orient=init_params.global_orient, hand_pose=init_params.hand_pose, betas=init_params.betas,
                      transl=init_params.transl)
        p, loss, _evals, _pts = self._fit(blocks, j3d, conf_3d, seq_ind, target_model_indices, joint_loss_weight,
                                          pose_preserve_weight, freeze_betas)
        joints, verts = self._mesh(torch.cat([p["global_orient"], p["hand_pose"]], dim=1), p["betas"], p["transl"])
        B = p["betas"].shape[0]
        fitted = MANOData(betas=p["betas"], global_orient=p["global_orient"],
                          body_pose=torch.zeros((B, 0), device=self.device), transl=p["transl"], hand_pose=p["hand_pose"])
        return BodyModelFitResult(params=fitted, vertices=verts, joints=joints, loss=loss.sum())


class FLAMEFitter(_ArticulatedFitter):
    """Optimisation fitter for the FLAME head model (misc_models.py:170-359)."""

    model_type = "flame"

    def fit_frame(self, init_params: FLAMEData, j3d: torch.Tensor, conf_3d: Optional[torch.Tensor] = None, seq_ind: int = 0,
                  target_model_indices: Optional[torch.Tensor] = None, joint_loss_weight: float = 600.0,
                  pose_preserve_weight: float = 5.0, freeze_betas: bool = False) -> BodyModelFitResult:
        B = _f32(j3d, self.device).shape[0]

        def opt(v, dim):      # absent blocks start at zero and ARE optimised (misc_models.py:213-232)
            return v if v is not None else torch.zeros(B, dim, device=self.device)

        blocks = dict(global_orient=init_params.global_orient, transl=init_params.transl,
                      jaw_pose=opt(init_params.jaw_pose, 3), expression=opt(init_params.expression, 10),
                      neck_pose=opt(init_params.neck_pose, 3), leye_pose=opt(init_params.leye_pose, 3),
                      reye_pose=opt(init_params.reye_pose, 3), betas=init_params.betas)
        p, loss, _evals, _pts = self._fit(blocks, j3d, conf_3d, seq_ind, target_model_indices, joint_loss_weight,
                                          pose_preserve_weight, freeze_betas)
        full_pose = torch.cat([p["global_orient"], p["neck_pose"], p["jaw_pose"], p["leye_pose"], p["reye_pose"]], dim=1)
        joints, verts = self._mesh(full_pose, torch.cat([p["betas"], p["expression"]], dim=1), p["transl"])
        fitted = FLAMEData(betas=p["betas"], global_orient=p["global_orient"],
                           body_pose=torch.zeros((B, 0), device=self.device), transl=p["transl"],
                           expression=p["expression"], jaw_pose=p["jaw_pose"], neck_pose=p["neck_pose"],
                           leye_pose=p["leye_pose"], reye_pose=p["reye_pose"])
        return BodyModelFitResult(params=fitted, vertices=verts, joints=joints, loss=loss.sum())
