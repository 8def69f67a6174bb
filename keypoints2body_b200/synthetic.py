"""Seeded synthetic stand-ins for the licensed assets.

The SMPL / SMPL-H / SMPL-X model files, ``gmm_08.pkl`` and the mean-parameter
file cannot be shipped or downloaded, so benchmarks, tests and ``smoke()`` use
random weights of the official shapes (6890 / 6890 / 10475 vertices, 24 / 52 /
55 joints, 207 / 459 / 486 pose-blend rows) around a human-scale skeleton.
Generators are deterministic (``torch.Generator`` on CPU) so the GPU box and
the authoring container build bit-identical assets from the seed alone.
"""

from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

# --- skeletons --------------------------------------------------------------
# Kinematic parents [smplx-from-memory]; first 22 entries are common to all
# three models.
_BODY22_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]
SMPL_PARENTS = _BODY22_PARENTS + [20, 21]


def _hand_parents(wrist: int, first: int) -> list[int]:
    """Five 3-joint finger chains rooted at ``wrist``; ``first`` = index of the first joint."""
    out = []
    for f in range(5):
        base = first + 3 * f
        out += [wrist, base, base + 1]
    return out


SMPLH_PARENTS = _BODY22_PARENTS + _hand_parents(20, 22) + _hand_parents(21, 37)
SMPLX_PARENTS = _BODY22_PARENTS + [15, 15, 15] + _hand_parents(20, 25) + _hand_parents(21, 40)

# Canonical T-pose joint centres in metres (y up), roughly an adult.
_BODY22_CENTRES = [
    (0.00, -0.24, 0.03), (0.07, -0.33, 0.02), (-0.07, -0.33, 0.02), (0.00, -0.12, 0.00),
    (0.10, -0.71, 0.02), (-0.10, -0.71, 0.02), (0.00, 0.02, 0.00), (0.09, -1.11, -0.03),
    (-0.09, -1.11, -0.03), (0.00, 0.07, 0.02), (0.11, -1.17, 0.09), (-0.11, -1.17, 0.09),
    (0.00, 0.28, -0.02), (0.08, 0.19, -0.01), (-0.08, 0.19, -0.01), (0.00, 0.37, 0.03),
    (0.17, 0.22, -0.02), (-0.17, 0.22, -0.02), (0.43, 0.21, -0.04), (-0.43, 0.21, -0.04),
    (0.68, 0.22, -0.03), (-0.68, 0.22, -0.03),
]


def _hand_centres(sign: float) -> list[tuple]:
    out = []
    for f in range(5):
        z = -0.03 + 0.02 * (f - 2)
        for k in range(3):
            out.append((sign * (0.77 + 0.03 * k), 0.21 - 0.005 * f, z))
    return out


# MANO: wrist + five 3-joint finger chains (16 joints); FLAME: head root, neck, jaw, two eyes (5 joints) [smplx-from-memory]
MANO_PARENTS = [-1] + _hand_parents(0, 1)
FLAME_PARENTS = [-1, 0, 1, 1, 1]


def skeleton(model_type: str):
    """(parents list, (n_j,3) float64 T-pose joint centres) for a model type."""
    if model_type == "smpl":
        centres = _BODY22_CENTRES + [(0.77, 0.21, -0.04), (-0.77, 0.21, -0.04)]
        return SMPL_PARENTS, np.asarray(centres)
    if model_type == "smplh":
        centres = _BODY22_CENTRES + _hand_centres(1.0) + _hand_centres(-1.0)
        return SMPLH_PARENTS, np.asarray(centres)
    if model_type == "smplx":
        face = [(0.00, 0.33, 0.06), (0.03, 0.40, 0.09), (-0.03, 0.40, 0.09)]
        centres = _BODY22_CENTRES + face + _hand_centres(1.0) + _hand_centres(-1.0)
        return SMPLX_PARENTS, np.asarray(centres)
    if model_type == "mano":
        centres = [(0.0, 0.0, 0.0)]
        for f in range(5):
            for k in range(3):
                centres.append((0.09 + 0.03 * k, 0.005 * (2 - f), 0.02 * (f - 2)))
        return MANO_PARENTS, np.asarray(centres)
    if model_type == "flame":
        centres = [(0.0, 0.0, 0.0), (0.0, 0.06, -0.01), (0.0, 0.08, 0.05), (0.03, 0.13, 0.07), (-0.03, 0.13, 0.07)]
        return FLAME_PARENTS, np.asarray(centres)
    raise ValueError(f"no synthetic skeleton for model_type={model_type}")


_NUM_VERTS = {"smpl": 6890, "smplh": 6890, "smplx": 10475, "mano": 778, "flame": 5023}
_NUM_EXTRA = 21  # vertex-picked extra joints (nose/eyes/ears/feet/finger tips)
# SMPL-H / SMPL-X carry enough vertex-picked joints for the reference's dict-block indices (hands 25..66, face 67..,
# constants.py:65-71): 52 + 21 = 73 and 55 + 76 = 131 model joints; MANO: 5 finger tips; FLAME: 51 landmarks
_NUM_EXTRA_BY_TYPE = {"smpl": 21, "smplh": 21, "smplx": 21, "mano": 5, "flame": 51}
NUM_EXTRA_SMPLX_BLOCKS = 76     # make_body_model("smplx", num_extra=76): 131 model joints, enough for a 64-landmark face block


def make_body_model(model_type: str = "smpl", seed: int = 0, dtype=torch.float32, num_extra: Optional[int] = None,
                    skin_layout: str = "interleaved"):
    """Random body-model weights of the official shapes.

    Returns a namespace with the smplx buffer names (``v_template (V,3)``,
    ``shapedirs (V,3,S)``, ``posedirs (9(n_j-1), 3V)``, ``J_regressor (n_j,V)``,
    ``parents (n_j,)``, ``lbs_weights (V,n_j)``, ``extra_vertex_ids``) plus the
    attributes the reference reads (``num_betas``, ``NUM_HAND_JOINTS``,
    ``num_expression_coeffs``).  Vertex v is owned by joint ``v mod n_j``; the
    regressor averages the owned vertices and skinning weights are 0.7 owner /
    0.3 parent, so rest joints sit at the skeleton centres (+- noise).

    ``skin_layout="interleaved"`` (default; every golden uses it) is the worst case for the mesh kernels: neighbouring
    vertices never share a joint.  ``"coherent"`` owns vertices in contiguous runs (joint ``v * n_j // V``), which is
    how the real models are ordered (a body part's vertices are mostly consecutive); same arithmetic, same sizes.
    """
    parents, centres = skeleton(model_type)
    n_j, n_v = len(parents), _NUM_VERTS[model_type]
    n_shape = 20 if model_type in ("smplx", "flame") else 10
    scale = {"mano": 0.2, "flame": 0.25}.get(model_type, 1.0)      # hand / head sized vertex clouds and blend shapes
    g = torch.Generator().manual_seed(seed)
    if skin_layout not in ("interleaved", "coherent"):
        raise ValueError(f"skin_layout must be 'interleaved' or 'coherent', got {skin_layout}")
    owner = torch.arange(n_v) % n_j if skin_layout == "interleaved" else (torch.arange(n_v) * n_j) // n_v
    c = torch.as_tensor(centres, dtype=torch.float64)
    v_template = c[owner] + 0.04 * scale * torch.randn(n_v, 3, generator=g, dtype=torch.float64)
    J_regressor = torch.zeros(n_j, n_v, dtype=torch.float64)
    J_regressor[owner, torch.arange(n_v)] = 1.0
    J_regressor /= J_regressor.sum(dim=1, keepdim=True)
    par = torch.as_tensor(parents)
    par_owner = torch.where(par[owner] < 0, owner, par[owner])
    lbs = torch.zeros(n_v, n_j, dtype=torch.float64)
    lbs[torch.arange(n_v), owner] += 0.7
    lbs[torch.arange(n_v), par_owner] += 0.3
    shapedirs = 0.01 * scale * torch.randn(n_v, 3, n_shape, generator=g, dtype=torch.float64)
    posedirs = 0.001 * scale * torch.randn((n_j - 1) * 9, 3 * n_v, generator=g, dtype=torch.float64)
    perm = torch.randperm(n_v, generator=g)
    n_extra = _NUM_EXTRA_BY_TYPE[model_type] if num_extra is None else int(num_extra)
    # the first 21 are the round-1 set (goldens depend on them); further ones are appended
    extra = torch.cat([perm[:min(n_extra, _NUM_EXTRA)].sort().values, perm[_NUM_EXTRA:n_extra].sort().values])
    m = SimpleNamespace(
        model_type=model_type,
        v_template=v_template.to(dtype),
        shapedirs=shapedirs.to(dtype),
        posedirs=posedirs.to(dtype),
        J_regressor=J_regressor.to(dtype),
        parents=par.long(),
        lbs_weights=lbs.to(dtype),
        extra_vertex_ids=extra.long(),
        num_betas=10,
        NUM_HAND_JOINTS=15,
        num_expression_coeffs=10 if model_type in ("smplx", "flame") else 0,
        NUM_BODY_JOINTS=21 if model_type != "smpl" else 23,
    )
    return m


def make_gmm(seed: int = 0, num_gaussians: int = 8, dim: int = 69) -> dict:
    """Synthetic max-mixture prior in the ``gmm_XX.pkl`` dict format (float64)."""
    g = torch.Generator().manual_seed(seed)
    means = 0.2 * torch.randn(num_gaussians, dim, generator=g, dtype=torch.float64)
    A = 0.05 * torch.randn(num_gaussians, dim, dim, generator=g, dtype=torch.float64)
    covars = A @ A.transpose(1, 2) + 0.05 * torch.eye(dim, dtype=torch.float64)
    w = torch.rand(num_gaussians, generator=g, dtype=torch.float64)
    return {
        "means": means.numpy(),
        "covars": covars.numpy(),
        "weights": (w / w.sum()).numpy(),
    }


def write_assets(folder: str, seed: int = 0, num_gaussians: int = 8) -> str:
    """Write ``gmm_XX.pkl`` and ``neutral_smpl_mean_params.npz`` (zeros) into ``folder``."""
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, f"gmm_{num_gaussians:02d}.pkl"), "wb") as f:
        pickle.dump(make_gmm(seed, num_gaussians), f)
    np.savez(
        os.path.join(folder, "neutral_smpl_mean_params.npz"),
        pose=np.zeros(72, np.float32),
        shape=np.zeros(10, np.float32),
    )
    return folder


def make_motion(
    num_frames: int,
    seed: int = 1,
    num_sequences: int = 1,
    pose_scale: float = 0.25,
    drift: float = 0.03,
):
    """Ground-truth SMPL-family parameters of smooth random motions.

    Returns dict of float32 tensors with leading dim ``num_sequences*num_frames``:
    ``pose (N,72)``, ``betas (N,10)``, ``transl (N,3)`` (SURVEY.md section 8d:
    ``pose_t = 0.25 N(0,1) + cumsum 0.03 N(0,1)``, ``beta = 0.5 N(0,1)`` per
    sequence, ``transl_t = cumsum 0.01 N(0,1)``).
    """
    g = torch.Generator().manual_seed(seed)
    S, T = num_sequences, num_frames
    pose = pose_scale * torch.randn(S, 1, 72, generator=g) + torch.cumsum(
        drift * torch.randn(S, T, 72, generator=g), dim=1
    )
    betas = (0.5 * torch.randn(S, 1, 10, generator=g)).expand(S, T, 10)
    transl = torch.cumsum(0.01 * torch.randn(S, T, 3, generator=g), dim=1)
    return {
        "pose": pose.reshape(S * T, 72).contiguous(),
        "betas": betas.reshape(S * T, 10).contiguous(),
        "transl": transl.reshape(S * T, 3).contiguous(),
    }


def kinematic_joints(model, pose, betas, transl, num_joints: Optional[int] = None):
    """Posed joint centres of the first ``num_joints`` joints, in plain torch (any device).

    Used only to synthesise target keypoints from ground-truth parameters; it is
    a data generator, not the product path (which is the CUDA kernel).
    ``pose`` is ``(N, 3*n_rot)`` axis-angle for joints ``0..n_rot-1``.
    """
    dev, dt = pose.device, pose.dtype
    parents = model.parents.tolist()
    n = num_joints or len(parents)
    vt = model.v_template.to(dev, dt)
    sd = model.shapedirs.to(dev, dt)[..., : betas.shape[1]]
    Jr = model.J_regressor.to(dev, dt)[:n]
    J0 = Jr @ vt
    JS = torch.einsum("jv,vkl->jkl", Jr, sd)
    J = J0[None] + torch.einsum("jkl,bl->bjk", JS, betas)
    r = pose.reshape(pose.shape[0], -1, 3)[:, :n]
    theta = torch.linalg.norm(r + 1e-8, dim=-1, keepdim=True)
    k = r / theta
    K = torch.zeros(*k.shape[:2], 3, 3, device=dev, dtype=dt)
    K[..., 0, 1], K[..., 0, 2] = -k[..., 2], k[..., 1]
    K[..., 1, 0], K[..., 1, 2] = k[..., 2], -k[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -k[..., 1], k[..., 0]
    s, c = torch.sin(theta)[..., None], torch.cos(theta)[..., None]
    R = torch.eye(3, device=dev, dtype=dt) + s * K + (1 - c) * (K @ K)
    Rw, t = [R[:, 0]], [J[:, 0]]
    for j in range(1, n):
        p = parents[j]
        t.append(torch.einsum("bij,bj->bi", Rw[p], J[:, j] - J[:, p]) + t[p])
        Rw.append(Rw[p] @ R[:, j])
    return torch.stack(t, dim=1) + transl[:, None]
