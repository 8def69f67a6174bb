"""Large randomized Adam parity sweep: CUDA path vs the oracle (test infrastructure; run on a GPU box).

    python tests/parity_sweep.py [frame|warp] > profiles/rNN_parity_sweep[_warp].json

``frame`` (default) = the one-thread-per-frame kernel (k2b_fit_batch), ``warp`` = the warp-per-frame kernel
(k2b_fit_chain), which additionally runs serial chains (frame t starts from frame t-1's result, 30 then 10
iterations with the temporal term) against the oracle's frame-by-frame loop.

For each (model, observations, schedule position) case: N random frames (random motions, 5 mm keypoint noise,
0.1 rad initial pose noise, random per-joint confidences incl. zeros), 10 or 30 Adam iterations on both sides,
then max / 99.9th percentile / median absolute differences of parameters, joints and vertices.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter
from oracle import reference_port as rp
from oracle.smplx_shim import BodyModelShim


def stats(a, b):
    d = np.abs(a - b).reshape(len(a), -1).max(axis=1)
    return {"max": float(d.max()), "p999": float(np.quantile(d, 0.999)), "median": float(np.median(d))}


def chain_case(mt, S, T, gmm, prior):
    """S sequences x T frames walked serially inside one launch vs the oracle's loop (batched over the sequences)."""
    w = syn.make_body_model(mt)
    g = torch.Generator().manual_seed(77 + S + T)
    mo = syn.make_motion(T, seed=4000 + S, num_sequences=S)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    tgt = (tgt + 0.005 * torch.randn(tgt.shape, generator=g)).reshape(S, T, 22, 3)
    conf = 0.3 + 0.7 * torch.rand(S, T, 22, generator=g)
    pose0 = mo["pose"].reshape(S, T, 72)[:, 0] + 0.1 * torch.randn(S, 72, generator=g)
    init = {k: None for k in rp.PARAM_ORDER}
    init.update(global_orient=pose0[:, :3].contiguous(), body_pose=pose0[:, 3:].contiguous(), betas=torch.zeros(S, 10),
                transl=mo["transl"].reshape(S, T, 3)[:, 0] + 0.02 * torch.randn(S, 3, generator=g))
    if mt in ("smplh", "smplx"):
        init.update(left_hand_pose=torch.zeros(S, 45), right_hand_pose=torch.zeros(S, 45))
    if mt == "smplx":
        init.update(expression=torch.zeros(S, 10), jaw_pose=torch.zeros(S, 3), leye_pose=torch.zeros(S, 3),
                    reye_pose=torch.zeros(S, 3))
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type=mt, gmm=gmm, use_lbfgs=False)
    ours = f.fit_chain({k: v for k, v in init.items() if v is not None}, tgt, conf)
    torch.cuda.synchronize()
    model, t0 = BodyModelShim(w), time.perf_counter()
    prev, rows = init, []
    for t in range(T):
        # the reference takes one (K,) confidence per call (world_space.py:163-164): sequences one by one
        outs = [rp.fit_frame(model, prior, {k: (v[s:s + 1] if v is not None else None) for k, v in prev.items()},
                             tgt[s:s + 1, t], conf[s, t], seq_ind=t, use_lbfgs=False) for s in range(S)]
        prev = {k: (torch.cat([o["params"][k] for o in outs]) if outs[0]["params"][k] is not None else None)
                for k in rp.PARAM_ORDER}
        rows.append((prev, torch.cat([o["joints"] for o in outs]), torch.cat([o["vertices"] for o in outs])))
    cpu_s = time.perf_counter() - t0
    p = ours["params"]
    def seqmajor(x):
        return x.cpu().numpy().reshape(S, T, -1)
    ref_pose = np.stack([torch.cat([r[0]["global_orient"], r[0]["body_pose"]], 1).numpy() for r in rows], 1)
    row = {"model": mt, "observations": "AMASS", "sequences": S, "frames_per_sequence": T,
           "schedule": "serial chain, 30 then 10 Adam iterations, per-frame confidences", "oracle_cpu_seconds": round(cpu_s, 1),
           "pose_rad": stats(seqmajor(torch.cat([p["global_orient"], p["body_pose"]], 1)).reshape(S * T, -1),
                             ref_pose.reshape(S * T, -1)),
           "betas": stats(seqmajor(p["betas"]).reshape(S * T, -1), np.stack([r[0]["betas"].numpy() for r in rows], 1).reshape(S * T, -1)),
           "transl_m": stats(seqmajor(p["transl"]).reshape(S * T, -1), np.stack([r[0]["transl"].numpy() for r in rows], 1).reshape(S * T, -1)),
           "joints_m": stats(seqmajor(ours["joints"]).reshape(S * T, -1), np.stack([r[1].numpy() for r in rows], 1).reshape(S * T, -1)),
           "vertices_m": stats(seqmajor(ours["vertices"]).reshape(S * T, -1), np.stack([r[2].numpy() for r in rows], 1).reshape(S * T, -1))}
    return row


def main():
    kernel = sys.argv[1] if len(sys.argv) > 1 else "frame"
    gmm = syn.make_gmm(0)
    prior = rp.GMMPrior(gmm)
    cases = [("smpl", "AMASS", 2048, 10, 3, False), ("smpl", "AMASS", 1024, 30, 0, False), ("smpl", "SMPL24", 1024, 10, 0, False),
             ("smpl", "AMASS", 1024, 10, 2, True), ("smplh", "AMASS", 512, 10, 1, False), ("smplx", "AMASS", 512, 5, 0, False)]
    report = []
    for mt, layout, n, iters, seq_ind, freeze in cases:
        w = syn.make_body_model(mt)
        K = 24 if layout == "SMPL24" else 22
        g = torch.Generator().manual_seed(1000 + n + iters + seq_ind)
        mo = syn.make_motion(n, seed=2000 + iters + seq_ind)
        tgt = syn.kinematic_joints(w, mo["pose"][:, :72] if K == 24 else mo["pose"][:, :66], mo["betas"], mo["transl"], K)
        tgt = tgt + 0.005 * torch.randn(tgt.shape, generator=g)
        pose = mo["pose"] + 0.1 * torch.randn(n, 72, generator=g)
        conf = torch.rand(K, generator=g)
        conf[torch.randperm(K, generator=g)[:2]] = 0.0           # two missing keypoints
        init = {k: None for k in rp.PARAM_ORDER}
        init.update(global_orient=pose[:, :3].contiguous(), body_pose=pose[:, 3:].contiguous(),
                    betas=0.3 * torch.randn(n, 10, generator=g), transl=mo["transl"] + 0.03 * torch.randn(n, 3, generator=g))
        if mt in ("smplh", "smplx"):
            init.update(left_hand_pose=torch.zeros(n, 45), right_hand_pose=torch.zeros(n, 45))
        if mt == "smplx":
            init.update(expression=0.3 * torch.randn(n, 10, generator=g), jaw_pose=torch.zeros(n, 3),
                        leye_pose=torch.zeros(n, 3), reye_pose=torch.zeros(n, 3))
        f = WorldSpaceFitter(w, joints_category=layout, model_type=mt, gmm=gmm, use_lbfgs=False)
        ours = f.fit_batch({k: v for k, v in init.items() if v is not None}, tgt, conf, seq_ind=seq_ind, num_iters=iters,
                           freeze_betas=freeze, kernel=kernel)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = rp.fit_frame(BodyModelShim(w), prior, init, tgt, conf, seq_ind=seq_ind, num_obs=K, use_lbfgs=False,
                           num_iters_first=iters, num_iters_followup=iters, freeze_betas=freeze)
        cpu_s = time.perf_counter() - t0
        p, r = ours["params"], ref["params"]
        row = {"model": mt, "observations": layout, "frames": n, "adam_iters": iters, "seq_ind": seq_ind,
               "freeze_betas": freeze, "oracle_cpu_seconds": round(cpu_s, 1),
               "pose_rad": stats(torch.cat([p["global_orient"], p["body_pose"]], 1).cpu().numpy(),
                                 torch.cat([r["global_orient"], r["body_pose"]], 1).numpy()),
               "betas": stats(p["betas"].cpu().numpy(), r["betas"].numpy()),
               "transl_m": stats(p["transl"].cpu().numpy(), r["transl"].numpy()),
               "joints_m": stats(ours["joints"].cpu().numpy(), ref["joints"].numpy()),
               "vertices_m": stats(ours["vertices"].cpu().numpy(), ref["vertices"].numpy())}
        if mt == "smplx":
            row["expression"] = stats(p["expression"].cpu().numpy(), r["expression"].numpy())
        ol, rl = ours["loss"].cpu().numpy().astype(np.float64), ref["loss_per_frame"].numpy().astype(np.float64) \
            if "loss_per_frame" in ref else None
        if rl is not None:
            row["loss_rel"] = float(np.max(np.abs(ol - rl) / np.abs(rl)))
        report.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    if kernel == "warp":
        for mt, S, T in (("smpl", 48, 6), ("smplx", 16, 4)):
            row = chain_case(mt, S, T, gmm, prior)
            report.append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"kernel": kernel, "tolerances": {"joints_m": 1e-4, "pose_rad": 1e-4, "betas": 1e-4, "transl_m": 1e-5},
                      "cases": report}, indent=1))


if __name__ == "__main__":
    main()
