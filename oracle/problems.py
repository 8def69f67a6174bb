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
