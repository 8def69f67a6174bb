"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Torch restatement of the ``smplx`` body-model forward [smplx-from-memory] that
the reference calls as ``self.smpl(**kwargs)``
(/root/reference/keypoints2body/core/fitters/world_space.py:192,278;
engine.py:114-118; shape.py:75-77).  smplx itself is absent, so this module is
passed to the unmodified reference as ``model=`` and to ``reference_port``.

Forward (public SMPL/LBS formulation)::

    v_shaped = v_template + shapedirs . [betas; expression]
    J        = J_regressor . v_shaped
    R_j      = I + sin(a) K + (1 - cos(a)) K K,  a = ||r_j + 1e-8||, K = skew(r_j / a)
    v_posed  = v_shaped + ((R[1:] - I).flatten() @ posedirs)
    T_j      = T_parent(j) . [[R_j, J_j - J_parent(j)], [0, 1]]
    joints   = T[:, :, :3, 3];  A_j = T_j with translation reduced by R_wj J_j
    verts    = (sum_j W_vj A_j) . (v_posed; 1)
    joints   = cat(joints, verts[:, extra_vertex_ids]);  (+ transl on both)

SMPL-H / SMPL-X input semantics adopted by this build:
the model consumes ``body_pose[:, :63]``; hands are 45-D axis-angle (no PCA,
no mean pose); full-pose order is SMPL-H ``[go, body63, lh, rh]`` and SMPL-X
``[go, body63, jaw, leye, reye, lh, rh]``.
"""

from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn


def rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """(N,3) axis-angle -> (N,3,3), with smplx's per-component 1e-8 inside the norm."""
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    k = rot_vecs / angle
    kx, ky, kz = k[:, 0], k[:, 1], k[:, 2]
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=1).view(-1, 3, 3)
    s = torch.sin(angle).unsqueeze(-1)
    c = torch.cos(angle).unsqueeze(-1)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


class BodyModelShim(nn.Module):
    """Callable stand-in for ``smplx.SMPL / SMPLH / SMPLX`` built from raw weight buffers."""

    def __init__(self, weights):
        super().__init__()
        self.model_type = weights.model_type
        for name in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights"):
            self.register_buffer(name, getattr(weights, name).clone())
        self.register_buffer("parents", weights.parents.clone().long())
        self.register_buffer("extra_vertex_ids", weights.extra_vertex_ids.clone().long())
        self.num_betas = int(weights.num_betas)
        self.NUM_HAND_JOINTS = int(weights.NUM_HAND_JOINTS)
        self.num_expression_coeffs = int(weights.num_expression_coeffs)
        self.num_joints = int(self.parents.numel())

    def _full_pose(self, global_orient, body_pose, kw):
        B = global_orient.shape[0]
        if self.model_type in ("mano", "flame"):
            def need(name, dim):
                v = kw.get(name)
                return v if v is not None else torch.zeros(B, dim, dtype=global_orient.dtype, device=global_orient.device)
            if self.model_type == "mano":          # [go | hand45]
                return torch.cat([global_orient, need("hand_pose", 45)], dim=1)
            # FLAME joint order: head root, neck, jaw, left eye, right eye
            return torch.cat([global_orient, need("neck_pose", 3), need("jaw_pose", 3), need("leye_pose", 3),
                              need("reye_pose", 3)], dim=1)

        def opt(name, dim):
            v = kw.get(name)
            if v is None:
                return torch.zeros(B, dim, dtype=global_orient.dtype, device=global_orient.device)
            return v

        if self.model_type == "smpl":
            return torch.cat([global_orient, body_pose], dim=1)
        body = body_pose[:, :63]
        hands = [opt("left_hand_pose", 45), opt("right_hand_pose", 45)]
        if self.model_type == "smplh":
            return torch.cat([global_orient, body] + hands, dim=1)
        face = [opt("jaw_pose", 3), opt("leye_pose", 3), opt("reye_pose", 3)]
        return torch.cat([global_orient, body] + face + hands, dim=1)

    def forward(self, global_orient=None, body_pose=None, betas=None, transl=None, left_hand_pose=None,
                right_hand_pose=None, expression=None, jaw_pose=None, leye_pose=None, reye_pose=None,
                hand_pose=None, neck_pose=None, return_full_pose=False):
        # every block is a NAMED parameter: the reference's MANO / FLAME fitters keep only the keyword arguments that
        # inspect.signature(model.forward) lists (core/fitters/misc_models.py:12-15)
        kw = dict(left_hand_pose=left_hand_pose, right_hand_pose=right_hand_pose, expression=expression,
                  jaw_pose=jaw_pose, leye_pose=leye_pose, reye_pose=reye_pose, hand_pose=hand_pose, neck_pose=neck_pose)
        B = global_orient.shape[0]
        dt, dev = global_orient.dtype, global_orient.device
        full_pose = self._full_pose(global_orient, body_pose, kw)
        shape = betas
        if self.model_type in ("smplx", "flame"):
            expr = kw.get("expression")
            if expr is None:
                expr = torch.zeros(B, self.num_expression_coeffs, dtype=dt, device=dev)
            shape = torch.cat([betas, expr], dim=1)
        if shape.shape[0] != B:
            shape = shape.expand(B, -1)

        v_shaped = self.v_template + torch.einsum("bl,mkl->bmk", shape, self.shapedirs)
        J = torch.einsum("bik,ji->bjk", v_shaped, self.J_regressor)
        R = rodrigues(full_pose.reshape(-1, 3)).view(B, -1, 3, 3)
        eye = torch.eye(3, dtype=dt, device=dev)
        pose_feature = (R[:, 1:] - eye).reshape(B, -1)
        v_posed = v_shaped + torch.matmul(pose_feature, self.posedirs).view(B, -1, 3)

        parents = self.parents
        rel = J.clone()
        rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
        local = torch.cat([R, rel.unsqueeze(-1)], dim=-1)  # (B,n_j,3,4)
        bottom = torch.tensor([0, 0, 0, 1], dtype=dt, device=dev).expand(B, self.num_joints, 1, 4)
        local = torch.cat([local, bottom], dim=2)
        chain = [local[:, 0]]
        for j in range(1, self.num_joints):
            chain.append(torch.matmul(chain[int(parents[j])], local[:, j]))
        T = torch.stack(chain, dim=1)
        posed_joints = T[:, :, :3, 3]
        shift = torch.matmul(T[:, :, :3, :3], J.unsqueeze(-1))  # R_w J
        A = T.clone()
        A[:, :, :3, 3] = T[:, :, :3, 3] - shift.squeeze(-1)

        Tv = torch.matmul(self.lbs_weights.unsqueeze(0).expand(B, -1, -1),
                          A.reshape(B, self.num_joints, 16)).view(B, -1, 4, 4)
        homo = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt, device=dev)], dim=2)
        verts = torch.matmul(Tv, homo.unsqueeze(-1))[:, :, :3, 0]
        joints = torch.cat([posed_joints, verts[:, self.extra_vertex_ids]], dim=1)
        if transl is not None:
            joints = joints + transl.unsqueeze(1)
            verts = verts + transl.unsqueeze(1)
        return SimpleNamespace(vertices=verts, joints=joints,
                               full_pose=full_pose if return_full_pose else None)
