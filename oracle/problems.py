"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Seeded fitting problems shared by the golden generators (``tests/golden/make_goldens_r2.py``, which feeds
them to the UNMODIFIED reference in the authoring container) and by the tests that replay them through the
CUDA path on the GPU box.  Everything is regenerated from seeds, so only outputs are stored in the goldens.
"""

from __future__ import annotations

import torch

from keypoints2body_b200 import synthetic as syn


def frame_problem(weights, n, seed, noise=0.005, init_noise=0.1):
    """``n`` independent frames: noisy AMASS-22 targets from a smooth random motion and an initialisation
    ``init_noise`` rad away from the ground truth (betas 0, translation 3 cm off)."""
    mo = syn.make_motion(n, seed=seed)
    tgt = syn.kinematic_joints(weights, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    g = torch.Generator().manual_seed(seed + 1)
    tgt = tgt + noise * torch.randn(tgt.shape, generator=g)
    pose = mo["pose"] + init_noise * torch.randn(n, 72, generator=g)
    init = dict(global_orient=pose[:, :3].contiguous(), body_pose=pose[:, 3:].contiguous(),
                betas=torch.zeros(n, 10), transl=mo["transl"] + 0.03 * torch.randn(n, 3, generator=g))
    return tgt, init


def chain_problem(weights, num_sequences, frames, seed, noise=0.005):
    """``num_sequences`` smooth motions of ``frames`` frames each: targets (S, T, 22, 3).  The sequence driver's
    own initialisation applies (mean pose, zero betas, root-aligned translation; api/sequence.py:177-191)."""
    mo = syn.make_motion(frames, seed=seed, num_sequences=num_sequences)
    tgt = syn.kinematic_joints(weights, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    g = torch.Generator().manual_seed(seed + 1)
    tgt = tgt + noise * torch.randn(tgt.shape, generator=g)
    return tgt.reshape(num_sequences, frames, 22, 3)


def mean_joint_error(joints22, targets):
    """Mean over the 22 joints of the Euclidean distance, per frame."""
    return (joints22 - targets).norm(dim=-1).mean(dim=-1)


def articulated_problem(mt, B, seed):
    """Hands / face / MANO / FLAME fitting problem: returns (shim model, targets (B,K,3), model indices (K,), init dict, B).

    smplx (131 model joints: 55 kinematic + 76 vertex-picked): the reference's dict blocks body 0..21, left hand 25..45,
    right hand 46..66, face 67..86 (constants.py:65-71); smplh (73 joints): body + both hand blocks; mano (21 joints:
    16 kinematic + 5 vertex-picked tips) and flame (56: 5 + 51 landmarks): every model joint."""
    from oracle.smplx_shim import BodyModelShim

    w = syn.make_body_model(mt, seed=0, num_extra=syn.NUM_EXTRA_SMPLX_BLOCKS if mt == "smplx" else None)
    model = BodyModelShim(w)
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, s=1.0):
        return s * torch.randn(*shape, generator=g)

    if mt in ("smplx", "smplh"):
        idx = list(range(22)) + list(range(25, 67)) + (list(range(67, 87)) if mt == "smplx" else [])
        gt = dict(global_orient=rn(B, 3, s=0.3), body_pose=rn(B, 69, s=0.25), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2),
                  left_hand_pose=rn(B, 45, s=0.3), right_hand_pose=rn(B, 45, s=0.3))
        if mt == "smplx":
            gt.update(expression=rn(B, 10, s=0.5), jaw_pose=rn(B, 3, s=0.2), leye_pose=rn(B, 3, s=0.1), reye_pose=rn(B, 3, s=0.1))
        init = {k: (v + rn(*v.shape, s=0.05)) for k, v in gt.items()}
        init["betas"] = torch.zeros(B, 10)
    elif mt == "mano":
        idx = list(range(21))
        gt = dict(global_orient=rn(B, 3, s=0.3), hand_pose=rn(B, 45, s=0.3), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2))
        init = dict(global_orient=gt["global_orient"] + rn(B, 3, s=0.05), hand_pose=gt["hand_pose"] + rn(B, 45, s=0.05),
                    betas=torch.zeros(B, 10), transl=gt["transl"] + rn(B, 3, s=0.01), body_pose=torch.zeros(B, 0))
    else:
        idx = list(range(56))
        gt = dict(global_orient=rn(B, 3, s=0.3), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2), expression=rn(B, 10, s=0.5),
                  jaw_pose=rn(B, 3, s=0.2), neck_pose=rn(B, 3, s=0.2), leye_pose=rn(B, 3, s=0.1), reye_pose=rn(B, 3, s=0.1))
        init = {k: (v + rn(*v.shape, s=0.03)) for k, v in gt.items()}
        init["betas"] = torch.zeros(B, 10)
        init["body_pose"] = torch.zeros(B, 0)
    with torch.no_grad():
        kw = {k: v for k, v in gt.items() if k != "body_pose" or mt in ("smplx", "smplh")}
        joints = model(**kw).joints
    idx_t = torch.tensor(idx, dtype=torch.long)
    tgt = joints[:, idx_t] + 0.003 * rn(B, len(idx), 3)
    return model, tgt, idx_t, init, B
