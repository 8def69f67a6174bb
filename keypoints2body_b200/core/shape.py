"""Shared-shape pre-pass over several frames (reference core/shape.py:10-115).

Objective per frame t (reference shape.py:71-100): model joints at the fixed pose, translated so
the model root meets the target root, squared error times conf^2 summed over the observed
joints, plus ``shape_prior_weight^2 * |betas|^2`` (added once PER FRAME); L-BFGS over betas,
``lr = 0.1``.  The CUDA kernel is ``k2b_shape_pass`` (one warp per sequence, lanes = frames).
"""

from __future__ import annotations

from typing import Optional

import torch


def optimize_shape_multi_frame(fitter, init_betas, pose_init, j3d_world, joints_category="SMPL24", num_iters=20,
                               step_size=1e-1, use_lbfgs=True, device=None, frame_indices: Optional[list] = None,
                               joints3d_conf: Optional[torch.Tensor] = None, shape_prior_weight=5.0):
    if joints_category not in ("SMPL24", "AMASS"):
        raise ValueError(f"No such joints category: {joints_category}")
    if not use_lbfgs:
        # the reference's Adam branch raises (never steps): shape.py:110-113, SURVEY.md Appendix A.1
        raise RuntimeError("element 0 of tensors does not require grad and does not have a grad_fn "
                           "(the reference's Adam shape pass is broken; use use_lbfgs=True)")
    return fitter.shape_pass(init_betas, pose_init, j3d_world, frame_indices=frame_indices,
                             conf=joints3d_conf, num_iters=num_iters, step_size=step_size,
                             shape_prior_weight=shape_prior_weight)
