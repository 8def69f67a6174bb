"""Articulated fitting problems for tools/bench_configs.py (product-side twin of oracle/problems.py: tools/ may not
import the oracle): ground-truth parameters, targets = the model's points at them through k2b_mesh_batch, an
initialisation a little off."""
import torch

from keypoints2body_b200 import _native as nat
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.body_model import extract_weights
from keypoints2body_b200.core.fitters.articulated import mesh_forward


def articulated_problem(mt, B, seed):
    w = syn.make_body_model(mt, seed=0, num_extra=syn.NUM_EXTRA_SMPLX_BLOCKS if mt == "smplx" else None)
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, s=1.0):
        return s * torch.randn(*shape, generator=g)

    if mt in ("smplx", "smplh"):
        idx = list(range(22)) + list(range(25, 67)) + (list(range(67, 87)) if mt == "smplx" else [])
        gt = dict(global_orient=rn(B, 3, s=0.3), body_pose=rn(B, 69, s=0.25), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2),
                  left_hand_pose=rn(B, 45, s=0.3), right_hand_pose=rn(B, 45, s=0.3))
        if mt == "smplx":
            gt.update(expression=rn(B, 10, s=0.5), jaw_pose=rn(B, 3, s=0.2), leye_pose=rn(B, 3, s=0.1), reye_pose=rn(B, 3, s=0.1))
        order = ["global_orient", "body_pose"] + (["jaw_pose", "leye_pose", "reye_pose"] if mt == "smplx" else []) + \
                ["left_hand_pose", "right_hand_pose"]
        full = torch.cat([gt[k][:, :63] if k == "body_pose" else gt[k] for k in order], dim=1)
        shape = torch.cat([gt["betas"]] + ([gt["expression"]] if mt == "smplx" else []), dim=1)
    elif mt == "mano":
        idx = list(range(21))
        gt = dict(global_orient=rn(B, 3, s=0.3), hand_pose=rn(B, 45, s=0.3), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2),
                  body_pose=torch.zeros(B, 0))
        full, shape = torch.cat([gt["global_orient"], gt["hand_pose"]], dim=1), gt["betas"]
    else:
        idx = list(range(56))
        gt = dict(global_orient=rn(B, 3, s=0.3), betas=rn(B, 10, s=0.5), transl=rn(B, 3, s=0.2), expression=rn(B, 10, s=0.5),
                  jaw_pose=rn(B, 3, s=0.2), neck_pose=rn(B, 3, s=0.2), leye_pose=rn(B, 3, s=0.1), reye_pose=rn(B, 3, s=0.1),
                  body_pose=torch.zeros(B, 0))
        full = torch.cat([gt[k] for k in ("global_orient", "neck_pose", "jaw_pose", "leye_pose", "reye_pose")], dim=1)
        shape = torch.cat([gt["betas"], gt["expression"]], dim=1)
    native = nat.NativeModel(extract_weights(w, mt), None, torch.device("cuda", torch.cuda.current_device()))
    joints, _ = mesh_forward(native, full.cuda(), shape.cuda(), gt["transl"].cuda(), with_vertices=True)
    idx_t = torch.tensor(idx, dtype=torch.long)
    tgt = joints[:, idx_t.cuda()].cpu() + 0.003 * rn(B, len(idx), 3)
    init = {k: (v + rn(*v.shape, s=0.03)) for k, v in gt.items()}
    init["betas"] = torch.zeros(B, 10)
    return w, tgt, idx_t, init
