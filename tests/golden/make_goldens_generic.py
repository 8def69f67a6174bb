"""Golden vectors for dict-block ("GENERIC") observations from the UNMODIFIED reference (authoring container only).

    python tests/golden/make_goldens_generic.py

The reference's public API with ``joints={"body": ...}`` (api/frame.py:79-104, adapters.py:224-304) fits the block
against explicit model-joint indices (world_space.py:198-201).  Stored in ``tests/golden/ref_goldens_generic.npz``:
one frame and one short sequence through ``optimize_params_frame`` / ``optimize_params_sequence`` (Adam), plus a
direct ``WorldSpaceFitter.fit_frame`` call with a partial, permuted ``target_model_indices``.
"""

from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from keypoints2body_b200 import synthetic as syn  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.smplx_shim import BodyModelShim  # noqa: E402

torch.set_num_threads(1)
OUT = os.path.join(os.path.dirname(__file__), "ref_goldens_generic.npz")
G = {}


def put(name, value):
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu().numpy()
    G[name] = np.asarray(value)


def main():
    ref = ref_loader.load_reference()
    from keypoints2body.core.fitters.world_space import WorldSpaceFitter
    from keypoints2body.models.smpl_data import SMPLData

    tmp = tempfile.mkdtemp()
    syn.write_assets(os.path.join(tmp, "data/models"), seed=0)
    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    mo = syn.make_motion(4, seed=51)
    tgt = syn.kinematic_joints(weights, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    g = torch.Generator().manual_seed(52)
    tgt = tgt + 0.005 * torch.randn(tgt.shape, generator=g)
    conf = 0.5 + 0.5 * torch.rand(4, 22, generator=g)
    block = torch.cat([tgt, conf[..., None]], dim=-1)          # (T, 22, 4): xyz + confidence
    put("gen_in_block", block)

    with ref_loader.reference_cwd(tmp):
        r = ref.optimize_params_frame({"body": block[0].numpy()}, body_model="smpl", model=model,
                                      config=dict(use_lbfgs=False))
        put("gen_frame_pose", r.params.pose)
        put("gen_frame_betas", r.params.betas)
        put("gen_frame_transl", r.params.transl)
        put("gen_frame_loss", r.loss)
        res = ref.optimize_params_sequence({"body": block.numpy()}, body_model="smpl", model=model,
                                           config=dict(frame=dict(use_lbfgs=False), use_shape_optimization=False))
        put("gen_seq_pose", torch.cat([x.params.pose for x in res]))
        put("gen_seq_betas", torch.cat([x.params.betas for x in res]))
        put("gen_seq_transl", torch.cat([x.params.transl for x in res]))
        put("gen_seq_loss", torch.stack([x.loss.reshape(()) for x in res]))

        # explicit, partial and permuted model indices straight into the fitter
        idx = torch.tensor([16, 0, 7, 21, 4, 12, 18, 1, 2, 20, 15, 8])
        fitter = WorldSpaceFitter(model, num_iters_first=12, use_lbfgs=False, joints_category="GENERIC")
        init = SMPLData(betas=torch.zeros(1, 10), global_orient=0.9 * mo["pose"][1:2, :3], body_pose=0.9 * mo["pose"][1:2, 3:],
                        transl=mo["transl"][1:2] + 0.02)
        rr = fitter.fit_frame(init, tgt[1:2, idx], conf[1, idx], seq_ind=0, target_model_indices=idx)
        put("gen_idx", idx)
        put("gen_idx_in_pose", init.pose)
        put("gen_idx_in_transl", init.transl)
        put("gen_idx_pose", rr.params.pose)
        put("gen_idx_betas", rr.params.betas)
        put("gen_idx_transl", rr.params.transl)
        put("gen_idx_loss", rr.loss)
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, {k: v.shape for k, v in G.items()})


if __name__ == "__main__":
    main()
