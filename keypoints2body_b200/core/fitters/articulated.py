"""General articulated fit: observations of hand joints / vertex-picked landmarks for SMPL-H / SMPL-X, MANO, FLAME.

Host side of ``k2b_artic_fit`` (csrc/artic_core.cuh).  Replaces, for inputs the body-keypoint kernels do not cover,
``WorldSpaceFitter.fit_frame`` with ``target_model_indices`` beyond the body joints
(/root/reference/keypoints2body/core/fitters/world_space.py:198-201) and the MANO / FLAME fitters
(core/fitters/misc_models.py:18-359).

A *layout* lists the parameter blocks in the order the reference hands them to its optimiser (the flat-vector order of
L-BFGS), says which model joints each rotation block drives, and carries the quadratic regularisers of the loss:

====== =============================================================== ====================================== ============
model  blocks (reference order)                                        regularisers (coefficient of x^2)      temporal term
====== =============================================================== ====================================== ============
smplh  go 3, body 69, transl 3, lh 45, rh 45, betas 10                 betas 5^2 (losses.py:56) + GMM + angle  body 1
smplx  go, body 69, transl, lh, rh, expr 10, jaw, leye, reye, betas    the same                                body 1
mano   go 3, hand 45, transl 3, betas 10 (misc_models.py:80-82)        hand 1e-2, betas 5 (:111)               hand 1
flame  go, transl, jaw, expr, neck, leye, reye, betas (:247-258)       jaw 1e-2, expr 1e-3, betas 5 (:286-291)  jaw 1, expr 0.2
====== =============================================================== ====================================== ============
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from ... import _native as nat
from ...body_model import BodyModelWeights

MAX_SKIN = 8

# (block name, width) in the reference's optimiser order
LAYOUTS = {
    "smpl": [("global_orient", 3), ("body_pose", 69), ("transl", 3), ("betas", 10)],
    "smplh": [("global_orient", 3), ("body_pose", 69), ("transl", 3), ("left_hand_pose", 45), ("right_hand_pose", 45),
              ("betas", 10)],
    "smplx": [("global_orient", 3), ("body_pose", 69), ("transl", 3), ("left_hand_pose", 45), ("right_hand_pose", 45),
              ("expression", 10), ("jaw_pose", 3), ("leye_pose", 3), ("reye_pose", 3), ("betas", 10)],
    "mano": [("global_orient", 3), ("hand_pose", 45), ("transl", 3), ("betas", 10)],
    "flame": [("global_orient", 3), ("transl", 3), ("jaw_pose", 3), ("expression", 10), ("neck_pose", 3),
              ("leye_pose", 3), ("reye_pose", 3), ("betas", 10)],
}
# model joints driven by each rotation block: (block, first joint, joints, offset inside the block)
POSE_MAP = {
    "smpl": [("global_orient", 0, 1, 0), ("body_pose", 1, 23, 0)],
    "smplh": [("global_orient", 0, 1, 0), ("body_pose", 1, 21, 0), ("left_hand_pose", 22, 15, 0), ("right_hand_pose", 37, 15, 0)],
    "smplx": [("global_orient", 0, 1, 0), ("body_pose", 1, 21, 0), ("jaw_pose", 22, 1, 0), ("leye_pose", 23, 1, 0),
              ("reye_pose", 24, 1, 0), ("left_hand_pose", 25, 15, 0), ("right_hand_pose", 40, 15, 0)],
    "mano": [("global_orient", 0, 1, 0), ("hand_pose", 1, 15, 0)],
    "flame": [("global_orient", 0, 1, 0), ("neck_pose", 1, 1, 0), ("jaw_pose", 2, 1, 0), ("leye_pose", 3, 1, 0),
              ("reye_pose", 4, 1, 0)],
}
SHAPE_BLOCKS = {"smpl": ["betas"], "smplh": ["betas"], "smplx": ["betas", "expression"], "mano": ["betas"], "flame": ["betas", "expression"]}
REG = {"smpl": {"betas": 25.0}, "smplh": {"betas": 25.0}, "smplx": {"betas": 25.0}, "mano": {"hand_pose": 1e-2, "betas": 5.0},
       "flame": {"jaw_pose": 1e-2, "expression": 1e-3, "betas": 5.0}}
KEEP = {"smpl": {"body_pose": 1.0}, "smplh": {"body_pose": 1.0}, "smplx": {"body_pose": 1.0}, "mano": {"hand_pose": 1.0},
        "flame": {"jaw_pose": 1.0, "expression": 0.2}}


def _f32(t, device):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def block_offsets(model_type: str):
    """(blocks, offset of each block, total length) of a model's parameter vector."""
    blocks = LAYOUTS[model_type]
    offset, o = {}, 0
    for name, w in blocks:
        offset[name] = o
        o += w
    return blocks, offset, o


def model_arrays(weights: BodyModelWeights, blocks=None, offset=None, n=None) -> dict:
    """Host arrays of ``k2b_artic_desc`` for one body model (contiguous numpy, the C ABI's dtypes)."""
    mt = weights.model_type
    if blocks is None:
        blocks, offset, n = block_offsets(mt)
    nj, ns, nv = weights.num_joints, weights.num_shape, weights.num_vertices
    Jr = weights.J_regressor.astype(np.float64)
    J0 = (Jr @ weights.v_template.astype(np.float64)).astype(np.float32)
    JS = np.einsum("jv,vkl->jkl", Jr, weights.shapedirs.astype(np.float64)).astype(np.float32)
    pose_src = np.full(3 * nj, -1, np.int32)
    for block, j0, count, off in POSE_MAP[mt]:
        base = offset[block] + off
        pose_src[3 * j0: 3 * (j0 + count)] = np.arange(base, base + 3 * count, dtype=np.int32)
    shape_src = np.concatenate([np.arange(offset[b], offset[b] + 10) for b in SHAPE_BLOCKS[mt]]).astype(np.int32)
    assert shape_src.shape[0] == ns, (shape_src.shape, ns)
    vid = weights.extra_vertex_ids.astype(np.int64)
    P = int(vid.shape[0])
    npf = 9 * (nj - 1)
    pv_t = weights.v_template[vid].astype(np.float32)
    pv_S = weights.shapedirs[vid].astype(np.float32)                                 # (P,3,ns)
    pd = weights.posedirs.reshape(npf, nv, 3)[:, vid, :]                               # (npf,P,3)
    pv_P = np.ascontiguousarray(np.transpose(pd, (1, 2, 0))).astype(np.float32)      # (P,3,npf)
    skin_idx = np.zeros((P, MAX_SKIN), np.int32)
    skin_w = np.zeros((P, MAX_SKIN), np.float32)
    for i, v in enumerate(vid):
        nz = np.nonzero(weights.lbs_weights[v])[0]
        if len(nz) > MAX_SKIN:
            raise NotImplementedError(f"vertex {v} is skinned to {len(nz)} joints; at most {MAX_SKIN} are supported")
        skin_idx[i, :len(nz)] = nz
        skin_w[i, :len(nz)] = weights.lbs_weights[v, nz]
    reg = np.zeros(n, np.float32)
    keep = np.zeros(n, np.float32)
    for name, w in blocks:
        reg[offset[name]: offset[name] + w] = REG[mt].get(name, 0.0)
        keep[offset[name]: offset[name] + w] = KEEP[mt].get(name, 0.0)
    out = dict(parents=weights.parents.astype(np.int32), J0=J0, JS=JS, pose_src=pose_src, shape_src=shape_src, pv_t=pv_t,
               pv_S=pv_S, pv_P=pv_P, skin_idx=skin_idx, skin_w=skin_w, reg=reg, keep=keep)
    return {k: np.ascontiguousarray(a) for k, a in out.items()}


class ArticulatedModel:
    """Device-resident ``k2b_artic`` for one body model; every vertex-picked joint of the model is observable."""

    def __init__(self, weights: BodyModelWeights, native_model: nat.NativeModel, with_body_priors: bool):
        mt = weights.model_type
        if mt not in LAYOUTS:
            raise ValueError(f"no articulated layout for model_type={mt}")
        self.model_type, self.native = mt, native_model
        self.lib, self.device = native_model.lib, native_model.device
        self.blocks = LAYOUTS[mt]
        self.offset, o = {}, 0
        for name, w in self.blocks:
            self.offset[name] = o
            o += w
        self.n = o
        nj, ns = weights.num_joints, weights.num_shape
        self.num_joints, self.num_points = nj, nj + weights.num_extra
        arrays = model_arrays(weights, self.blocks, self.offset, self.n)
        P = int(arrays["pv_t"].shape[0])
        hold = [arrays[k] for k in ("parents", "J0", "JS", "pose_src", "shape_src", "pv_t", "pv_S", "pv_P", "skin_idx",
                                    "skin_w", "reg", "keep")]
        fp, ip = nat._fp, nat._ip
        desc = nat.ArticDesc(
            num_joints=nj, num_shape=ns, num_params=self.n, num_picked=P, parents=ip(hold[0]), J0=fp(hold[1]), JS=fp(hold[2]),
            pose_src=ip(hold[3]), shape_src=ip(hold[4]), transl_src=self.offset["transl"],
            pv_template=fp(hold[5]) if P else None, pv_shapedirs=fp(hold[6]) if P else None,
            pv_posedirs=fp(hold[7]) if P else None, pv_skin_idx=ip(hold[8]) if P else None,
            pv_skin_w=fp(hold[9]) if P else None, reg_w=fp(hold[10]), keep_w=fp(hold[11]),
            body_off=self.offset["body_pose"] if with_body_priors else -1,
            prior_model=native_model.handle if with_body_priors else None)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(self.lib.k2b_artic_create(C.byref(desc), C.byref(handle)))
        self.handle = handle

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.k2b_artic_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------------
    def pack(self, blocks: dict, B: int) -> tuple[torch.Tensor, torch.Tensor]:
        """Blocks (absent / None = zeros, kept fixed) -> x (B,n) and the frozen mask (n,) of absent blocks."""
        dev = self.device
        parts, frozen = [], torch.zeros(self.n, dtype=torch.uint8)
        for name, w in self.blocks:
            v = _f32(blocks.get(name), dev)
            if v is None or v.shape[-1] == 0:
                v = torch.zeros(B, w, device=dev)
                frozen[self.offset[name]: self.offset[name] + w] = 1
            if v.dim() != 2 or v.shape[1] != w or v.shape[0] not in (1, B):
                raise ValueError(f"{name} must be (B={B}, {w}), got {tuple(v.shape)}")
            parts.append(v.expand(B, w))
        return torch.cat(parts, dim=1).contiguous(), frozen

    def unpack(self, x: torch.Tensor) -> dict:
        return {name: x[:, self.offset[name]: self.offset[name] + w].contiguous() for name, w in self.blocks}

    def run(self, mode: int, x0, j3d, conf, model_indices, *, keep_on: bool, num_iters: int, lr: float,
            joint_loss_weight: float, pose_preserve_weight: float, frozen: torch.Tensor, want_points=True):
        dev = self.device
        B, K = x0.shape[0], j3d.shape[1]
        idx = torch.as_tensor(model_indices).reshape(-1).to(torch.int64)
        if idx.numel() != K:
            raise ValueError(f"{K} observations for {idx.numel()} model indices")
        if int(idx.min()) < 0 or int(idx.max()) >= self.num_points:
            raise ValueError(f"model index out of range: this model has {self.num_points} joints "
                             f"({self.num_joints} kinematic + {self.num_points - self.num_joints} vertex-picked)")
        idx = idx.to(device=dev, dtype=torch.int32).contiguous()
        targets = _f32(j3d, dev)
        conf = _f32(conf, dev)
        conf_pf = conf is not None and conf.dim() == 2
        out = dict(x=torch.empty(B, self.n, device=dev), loss=torch.empty(B, device=dev),
                   grad=torch.empty(B, self.n, device=dev) if mode == nat.ARTIC_EVAL else None,
                   points=torch.empty(B, K, 3, device=dev) if want_points else None,
                   evals=torch.empty(B, dtype=torch.int32, device=dev))
        ws = self.native.workspace("artic", self.lib.k2b_artic_workspace_bytes(self.handle, B, mode, int(num_iters)))
        fz = frozen.to(dev).contiguous()
        a = nat.ArticFitArgs(
            num_frames=B, num_obs=K, mode=mode, num_iters=int(num_iters), conf_per_frame=int(conf_pf), lr=float(lr),
            joint_loss_weight=float(joint_loss_weight), keep_scale=float(pose_preserve_weight) ** 2 if keep_on else 0.0,
            obs_idx=nat.ptr(idx), targets=nat.ptr(targets), conf=nat.ptr(conf), init_x=nat.ptr(x0), keep_x=None,
            frozen=nat.ptr(fz), out_x=nat.ptr(out["x"]), out_loss=nat.ptr(out["loss"]), out_grad=nat.ptr(out["grad"]),
            out_points=nat.ptr(out["points"]), out_evals=nat.ptr(out["evals"]), out_gmm_component=None,
            workspace=nat.ptr(ws), workspace_bytes=ws.numel())
        with torch.cuda.device(dev):
            nat.check(self.lib.k2b_artic_fit(self.handle, C.byref(a), nat.current_stream()))
        return out


def get_articulated(native_model: nat.NativeModel, with_body_priors: bool) -> ArticulatedModel:
    """One ArticulatedModel per native model (built on first use)."""
    key = "_artic_prior" if with_body_priors else "_artic"
    am = getattr(native_model, key, None)
    if am is None:
        am = ArticulatedModel(native_model.weights, native_model, with_body_priors)
        setattr(native_model, key, am)
    return am


def mesh_forward(native: nat.NativeModel, full_pose, shape, transl, with_vertices=True):
    """Full mesh of any model the library holds (k2b_mesh_batch): joints (B, n_j + extras, 3), vertices (B, V, 3)."""
    dev = native.device
    B = full_pose.shape[0]
    joints = torch.empty(B, native.num_joints + native.num_extra, 3, device=dev)
    verts = torch.empty(B, native.num_vertices, 3, device=dev) if with_vertices else None
    ws = native.workspace("mesh", native.lib.k2b_mesh_workspace_bytes(native.handle, B))
    full_pose, shape = full_pose.contiguous(), shape.contiguous()
    transl = transl.contiguous() if transl is not None else None
    a = nat.MeshArgs(num_frames=B, full_pose=nat.ptr(full_pose), shape=nat.ptr(shape), transl=nat.ptr(transl),
                     out_vertices=nat.ptr(verts), out_joints=nat.ptr(joints), workspace=nat.ptr(ws),
                     workspace_bytes=ws.numel(), max_ctas=0)
    with torch.cuda.device(dev):
        nat.check(native.lib.k2b_mesh_batch(native.handle, C.byref(a), nat.current_stream()))
    return joints, verts


def articulated_fit(am: ArticulatedModel, blocks: dict, j3d, conf, model_indices, *, seq_ind: int, num_iters: int,
                    use_lbfgs: bool, lr: float, joint_loss_weight: float, pose_preserve_weight: float, freeze_betas: bool):
    """One fit_frame call (B independent frames) through k2b_artic_fit; returns (blocks dict, loss (B,), evals (B,))."""
    j3d = _f32(j3d, am.device)
    B = j3d.shape[0]
    x0, frozen = am.pack(blocks, B)
    if freeze_betas:
        frozen[am.offset["betas"]: am.offset["betas"] + 10] = 1
    out = am.run(nat.ARTIC_LBFGS if use_lbfgs else nat.ARTIC_ADAM, x0, j3d, conf, model_indices, keep_on=seq_ind > 0,
                 num_iters=num_iters, lr=lr, joint_loss_weight=joint_loss_weight, pose_preserve_weight=pose_preserve_weight,
                 frozen=frozen)
    return am.unpack(out["x"]), out["loss"], out["evals"], out["points"]
