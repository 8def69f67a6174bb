"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement of the reference's world-space fitting path on top of torch
autograd and torch.optim (the reference's own optimisers, present wherever
torch is).  It exists because ``/root/reference`` cannot travel to the GPU
box; it is pinned against the unmodified reference by
``tests/test_oracle_vs_reference.py`` using ``tests/golden/*.npz``.

Each function cites the reference lines it follows.  Differences are limited
to packaging: parameters travel as one dict of tensors instead of dataclasses,
and the GMM is passed as arrays instead of being unpickled from a CWD-relative
path.
"""

from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

PARAM_ORDER = ("global_orient", "body_pose", "transl", "left_hand_pose", "right_hand_pose",
               "expression", "jaw_pose", "leye_pose", "reye_pose", "betas")
"""Order of the optimiser's parameter list = L-BFGS flat-vector layout
(/root/reference/keypoints2body/core/fitters/world_space.py:215-229)."""


class GMMPrior:
    """``MaxMixturePrior`` with ``use_merged=True``
    (/root/reference/keypoints2body/core/prior.py:101-195)."""

    def __init__(self, gmm: dict, dtype=torch.float32):
        np_dt = np.float32 if dtype == torch.float32 else np.float64
        means = np.asarray(gmm["means"]).astype(np_dt)
        covs = np.asarray(gmm["covars"]).astype(np_dt)
        # prior.py:152-155 -- inverse of the (already cast) covariances, cast again
        prec = np.stack([np.linalg.inv(c) for c in covs]).astype(np_dt)
        # prior.py:158-163 -- constant uses the UNcast covariances / weights
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm["covars"]])
        const = (2 * np.pi) ** (69 / 2.0)
        nll_w = np.asarray(gmm["weights"] / (const * (sqrdets / sqrdets.min())))
        self.means = torch.tensor(means, dtype=dtype)
        self.precisions = torch.tensor(prec, dtype=dtype)
        self.nll_weights = torch.tensor(nll_w, dtype=dtype).unsqueeze(0)

    def __call__(self, pose: torch.Tensor, betas=None) -> torch.Tensor:
        # prior.py:182-195
        d = pose.unsqueeze(1) - self.means
        pd = torch.einsum("mij,bmj->bmi", [self.precisions, d])
        quad = (pd * d).sum(dim=-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)
        return torch.min(ll, dim=1)[0]


def gmof(x, sigma):
    """Geman-McClure (/root/reference/keypoints2body/core/losses.py:6-10)."""
    x2, s2 = x ** 2, sigma ** 2
    return (s2 * x2) / (s2 + x2)


def angle_prior(pose):
    """Elbow / knee bending penalty (losses.py:13-21)."""
    sign = torch.tensor([1.0, -1.0, -1.0, -1.0], device=pose.device)
    return torch.exp(pose[:, [52, 55, 9, 12]] * sign) ** 2


def body_fitting_loss_3d(body_pose, preserve_pose, betas, model_joints, j3d, pose_prior,
                         joints3d_conf, joint_loss_weight=500.0, pose_preserve_weight=0.0,
                         sigma=100, pose_prior_weight=4.78 * 1.5, shape_prior_weight=5.0,
                         angle_prior_weight=15.2, reduce=True):
    """Total loss (losses.py:24-67); ``reduce=False`` returns the per-frame terms."""
    if joints3d_conf.dim() == 1:
        joints3d_conf = joints3d_conf.view(1, -1)
    err = gmof(model_joints - j3d, sigma)
    joint = (joint_loss_weight ** 2) * ((joints3d_conf ** 2) * err.sum(dim=-1)).sum(dim=-1)
    prior = (pose_prior_weight ** 2) * pose_prior(body_pose, betas)
    angle = (angle_prior_weight ** 2) * angle_prior(body_pose).sum(dim=-1)
    shape = (shape_prior_weight ** 2) * (betas ** 2).sum(dim=-1)
    keep = (pose_preserve_weight ** 2) * ((body_pose - preserve_pose) ** 2).sum(dim=-1)
    total = joint + prior + angle + shape + keep
    return total.sum() if reduce else total


def _model_kwargs(p: dict) -> dict:
    return {k: v for k, v in p.items() if v is not None}


def frame_loss(model, prior, p: dict, preserve_pose, j3d, conf, joint_index,
               joint_loss_weight, pose_preserve_weight, reduce=True):
    """``compute_loss`` closure of the fitter (world_space.py:173-212)."""
    out = model(**_model_kwargs(p))
    sub = out.joints[:, joint_index, :]
    return body_fitting_loss_3d(p["body_pose"], preserve_pose, p["betas"], sub, j3d, prior,
                                conf, joint_loss_weight, pose_preserve_weight, reduce=reduce)


def evaluate(model, prior, params: dict, preserve_pose, j3d, conf, *, num_obs=22,
             joint_loss_weight=600.0, pose_preserve_weight=0.0):
    """One function evaluation: per-frame loss, d(sum loss)/d(params), model joints."""
    p = {k: (v.clone().detach().requires_grad_(True) if v is not None else None)
         for k, v in params.items()}
    idx = torch.arange(num_obs)
    out = model(**_model_kwargs(p))
    sub = out.joints[:, idx, :]
    per_frame = body_fitting_loss_3d(p["body_pose"], preserve_pose, p["betas"], sub, j3d, prior,
                                     conf, joint_loss_weight, pose_preserve_weight, reduce=False)
    per_frame.sum().backward()
    grads = {k: (v.grad if v is not None and v.grad is not None else
                 (torch.zeros_like(v) if v is not None else None)) for k, v in p.items()}
    return per_frame.detach(), grads, out.joints.detach()


def fit_frame(model, prior, init: dict, j3d, conf, *, seq_ind=0, num_obs=22, use_lbfgs=True,
              step_size=1e-2, num_iters_first=30, num_iters_followup=10,
              joint_loss_weight=600.0, pose_preserve_weight=5.0, freeze_betas=False,
              trace: Optional[list] = None):
    """``WorldSpaceFitter.fit_frame`` (world_space.py:93-323).

    ``init`` maps PARAM_ORDER names to (B,dim) tensors (absent / None blocks are
    skipped exactly like the reference's ``is not None`` tests); ``j3d`` is
    (B,K,3); ``conf`` is (K,) (a 2-D conf is reduced to its first row,
    world_space.py:163-164).  Returns dict(params, joints, vertices, loss).
    ``trace`` (optional list) receives (loss, flat_grad) of every closure call.
    """
    p = {k: (init[k].clone().detach() if init.get(k) is not None else None) for k in PARAM_ORDER}
    for k, v in p.items():
        if v is not None and k != "betas":
            v.requires_grad_(True)
    p["betas"].requires_grad_(not freeze_betas)
    preserve = p["body_pose"].clone().detach()
    if conf is None:
        conf = torch.ones(j3d.shape[1])
    elif conf.dim() == 2:
        conf = conf[0]
    idx = torch.arange(num_obs)
    w_keep = pose_preserve_weight if seq_ind > 0 else 0.0

    def loss_fn():
        return frame_loss(model, prior, p, preserve, j3d[:, idx], conf[idx], idx,
                          joint_loss_weight, w_keep)

    num_iters = num_iters_first if seq_ind == 0 else num_iters_followup
    opt_params = [p[k] for k in PARAM_ORDER if p[k] is not None and (k != "betas" or not freeze_betas)]
    if use_lbfgs:
        opt = torch.optim.LBFGS(opt_params, max_iter=num_iters, lr=step_size,
                                line_search_fn="strong_wolfe")

        def closure():
            opt.zero_grad()
            loss = loss_fn()
            loss.backward()
            if trace is not None:
                trace.append((float(loss.detach()), torch.cat([q.grad.reshape(-1) for q in opt_params]).clone()))
            return loss

        opt.step(closure)
        with torch.no_grad():
            final_loss = loss_fn()
    else:
        opt = torch.optim.Adam(opt_params, lr=step_size, betas=(0.9, 0.999))
        final_loss = None
        for _ in range(num_iters):
            opt.zero_grad()
            loss = loss_fn()
            loss.backward()
            opt.step()
            final_loss = loss.detach()
    with torch.no_grad():
        out = model(**_model_kwargs(p))
    return {
        "params": {k: (v.detach() if v is not None else None) for k, v in p.items()},
        "joints": out.joints.detach(),
        "vertices": out.vertices.detach(),
        "loss": final_loss,
    }


def camera_fitting_loss_3d(model_joints, camera_t, camera_t_est, j3d, depth_loss_weight=100.0):
    """Stage-1 camera loss (core/losses.py:70-93); RHip, LHip, RShoulder, LShoulder = 2, 1, 17, 16 in both
    joint maps.  Note the (B,4,3) + (B,3) broadcast: the depth term is added to each of the 4 joint rows."""
    sel = [2, 1, 17, 16]
    mj = model_joints + camera_t
    err = (j3d[:, sel] - mj[:, sel]) ** 2
    depth = (depth_loss_weight ** 2) * (camera_t - camera_t_est) ** 2
    return (err + depth).sum()


def fit_frame_camera(model, prior, init: dict, j3d, conf, *, seq_ind=0, num_obs=22, use_lbfgs=True, step_size=1e-2,
                     num_iters=100, joint_loss_weight=600.0, pose_preserve_weight=5.0, freeze_betas=True,
                     init_cam_t=None):
    """``CameraSpaceFitter.fit_frame`` (core/fitters/camera_space.py:81-339), B = 1.

    ``init_cam_t`` (camera_space.py:91,133) is both the start and the depth reference.  When it is left
    to the stage-0 estimate the stage-1 translation gradient is analytically zero at the start, so Adam's
    first step there is decided by rounding noise (g / (|g| + 1e-8)): only this restatement, which
    issues the same torch ops in the same order, follows that trajectory; an independent implementation
    is compared step for step from a caller-supplied ``init_cam_t`` instead."""
    w_keep = pose_preserve_weight if seq_ind > 0 else 0.0
    body_pose = init["body_pose"].detach().clone()
    go = init["global_orient"].detach().clone()
    betas = init["betas"].detach().clone()
    idx = torch.arange(num_obs)
    with torch.no_grad():
        mj = model(global_orient=go, body_pose=body_pose, betas=betas).joints
    sel = [2, 1, 17, 16]
    if init_cam_t is None:
        init_cam_t = ((j3d[:, sel] - mj[:, sel]).sum(dim=1) / 4.0).detach()      # guess_init_3d :16-41
    else:
        init_cam_t = init_cam_t.detach().clone()
    cam_t = init_cam_t.clone()
    preserve = body_pose.detach().clone()
    go.requires_grad_(True)
    cam_t.requires_grad_(True)

    def make_opt(params):
        if use_lbfgs:
            return torch.optim.LBFGS(params, max_iter=num_iters, lr=step_size, line_search_fn="strong_wolfe")
        return torch.optim.Adam(params, lr=step_size, betas=(0.9, 0.999))

    def run(opt, loss_fn):
        if use_lbfgs:
            def closure():
                opt.zero_grad()
                loss = loss_fn()
                loss.backward()
                return loss
            opt.step(closure)
        else:
            for _ in range(num_iters):
                loss = loss_fn()
                opt.zero_grad()
                loss.backward()
                opt.step()

    def stage1():
        j = model(global_orient=go, body_pose=body_pose, betas=betas).joints
        return camera_fitting_loss_3d(j[:, idx], cam_t, init_cam_t, j3d[:, idx])

    run(make_opt([go, cam_t]), stage1)
    body_pose.requires_grad_(True)
    move_betas = seq_ind == 0 or not freeze_betas
    betas.requires_grad_(move_betas)
    params = [body_pose, betas, go, cam_t] if move_betas else [body_pose, go, cam_t]
    if conf is None:
        conf = torch.ones(j3d.shape[1])

    def stage2(jw=joint_loss_weight, wk=w_keep):
        j = model(global_orient=go, body_pose=body_pose, betas=betas).joints
        return body_fitting_loss_3d(body_pose, preserve, betas, j[:, idx] + cam_t, j3d[:, idx], prior, conf[idx], jw, wk)

    run(make_opt(params), stage2)
    with torch.no_grad():
        out = model(global_orient=go, body_pose=body_pose, betas=betas)
        final_loss = stage2(600.0, 0.0)                                           # :316-326
    return {"params": {"global_orient": go.detach(), "body_pose": body_pose.detach(), "betas": betas.detach(),
                       "transl": cam_t.detach()},
            "joints": out.joints.detach(), "vertices": out.vertices.detach(), "loss": final_loss}


def guess_transl(model, pose, betas, j3d):
    """``guess_init_transl_from_root`` (world_space.py:13-50): root-joint alignment."""
    with torch.no_grad():
        out = model(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    return (j3d[:, 0, :] - out.joints[:, 0, :]).detach()


def optimize_shape(model, init_betas, pose, j3d_seq, conf, *, frame_indices, num_obs=22,
                   num_iters=40, step_size=1e-1, shape_prior_weight=5.0):
    """``optimize_shape_multi_frame`` L-BFGS branch (core/shape.py:10-115)."""
    betas = init_betas.clone().detach().requires_grad_(True)
    idx = torch.arange(num_obs)
    opt = torch.optim.LBFGS([betas], max_iter=num_iters, lr=step_size, line_search_fn="strong_wolfe")

    def closure():
        opt.zero_grad()
        total = betas.new_tensor(0.0)
        for t in frame_indices:
            out = model(global_orient=pose[t:t + 1, :3], body_pose=pose[t:t + 1, 3:], betas=betas)
            j = out.joints
            shift = j3d_seq[t:t + 1, 0, :] - j[:, 0, :]
            e = ((j + shift.unsqueeze(1))[:, idx] - j3d_seq[t:t + 1][:, idx]) ** 2
            total = total + ((conf[idx] ** 2) * e.sum(dim=-1)).sum() \
                + (shape_prior_weight ** 2) * (betas ** 2).sum()
        total.backward()
        return total

    opt.step(closure)
    return betas.detach()


def flatten(params: dict, freeze_betas=False) -> torch.Tensor:
    """Concatenate the optimised blocks in PARAM_ORDER (the L-BFGS flat layout)."""
    keys = [k for k in PARAM_ORDER if params.get(k) is not None and (k != "betas" or not freeze_betas)]
    return torch.cat([params[k] for k in keys], dim=1)


# ---- MPJAE evaluation (cli/eval.py:88-157), numpy float32 like the reference ------------------------
def eval_rotmat(rotvec, eps=1e-8):
    """``rotvec_to_rotmat`` (cli/eval.py:88-127): closed-form Rodrigues with a Taylor branch at theta <= eps."""
    import numpy as np
    rv = np.asarray(rotvec, dtype=np.float32)
    x, y, z = rv[..., 0], rv[..., 1], rv[..., 2]
    t2 = x * x + y * y + z * z
    th = np.sqrt(t2)
    big = th > eps
    ths = np.where(big, th, 1.0).astype(np.float32)
    a = np.sin(ths) / ths
    b = (1.0 - np.cos(ths)) / (ths * ths)
    a = np.where(big, a, 1.0 - t2 / 6.0 + t2 * t2 / 120.0).astype(np.float32)
    b = np.where(big, b, 0.5 - t2 / 24.0 + t2 * t2 / 720.0).astype(np.float32)
    # products are formed first, then scaled (b * (x*y), not (b*x) * y), as cli/eval.py:106-126 does
    xy, xz, yz, xx, yy, zz = x * y, x * z, y * z, x * x, y * y, z * z
    r = np.empty(rv.shape[:-1] + (3, 3), dtype=np.float32)
    r[..., 0, 0] = 1.0 - b * (yy + zz); r[..., 0, 1] = b * xy - a * z; r[..., 0, 2] = b * xz + a * y
    r[..., 1, 0] = b * xy + a * z; r[..., 1, 1] = 1.0 - b * (xx + zz); r[..., 1, 2] = b * yz - a * x
    r[..., 2, 0] = b * xz - a * y; r[..., 2, 1] = b * yz + a * x; r[..., 2, 2] = 1.0 - b * (xx + yy)
    return r


def angular_error_deg(pred_rotvec, gt_rotvec):
    """``compute_angular_error_deg`` (cli/eval.py:130-139)."""
    import numpy as np
    tr = np.sum(eval_rotmat(pred_rotvec) * eval_rotmat(gt_rotvec), axis=(-1, -2))
    c = np.clip((tr - 1.0) * 0.5, -1.0 + 1e-6, 1.0 - 1e-6)
    return np.degrees(np.arccos(c))


def evaluate_pose_pair(pred_pose, gt_pose):
    """``evaluate_pose_pair`` (cli/eval.py:142-157) -> (mean_deg, sum_deg, count)."""
    import numpy as np
    n = min(gt_pose.shape[0], pred_pose.shape[0])
    d = min(gt_pose.shape[1], pred_pose.shape[1]) // 3 * 3
    ang = angular_error_deg(pred_pose[:n, :d].reshape(n, d // 3, 3), gt_pose[:n, :d].reshape(n, d // 3, 3))
    s = float(np.sum(ang, dtype=np.float64))
    return s / ang.size, s, int(ang.size)
