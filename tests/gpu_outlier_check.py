"""Diagnostic (test infrastructure): the one outlier of the SMPL24 case of `tests/parity_sweep.py warp` (frame 626,
0.02 rad on one pose entry, everything else <= 2e-5).  Re-runs the case on both kernels, finds the worst frame and
element, shows the error after 1, 2, 3, 5 Adam iterations for 0 / 1 / 2 helper warps, and follows the gap between
the two most likely mixture components along the ORACLE's trajectory.

Finding (B200, round 1): the deviation is exactly 2 * lr after the FIRST iteration on body-pose entry 52, for every
helper count, and the mixture gap is 10 % (no tie).  At the initial point the reference's own gradient of that entry
is 1.7e-3 while the median |gradient| is 3.3e3 (max 7.5e4): kinematic, prior and angle-prior terms cancel to a few
ulps, so the SIGN of Adam's first step (step = lr * g / (|g| + 1e-8) ~ +-lr) is rounding noise in the reference too.
The one-thread-per-frame kernel happens to round like the oracle here, the warp kernel (different summation order)
does not.  Same mechanism as the camera-space start documented in DESIGN.md (G6')."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter
from oracle import reference_port as rp
from oracle.smplx_shim import BodyModelShim

gmm = syn.make_gmm(0); prior = rp.GMMPrior(gmm)
mt, layout, n, iters, seq_ind = "smpl", "SMPL24", 1024, 10, 0
w = syn.make_body_model(mt); K = 24
g = torch.Generator().manual_seed(1000 + n + iters + seq_ind)
mo = syn.make_motion(n, seed=2000 + iters + seq_ind)
tgt = syn.kinematic_joints(w, mo["pose"][:, :72], mo["betas"], mo["transl"], K)
tgt = tgt + 0.005 * torch.randn(tgt.shape, generator=g)
pose = mo["pose"] + 0.1 * torch.randn(n, 72, generator=g)
conf = torch.rand(K, generator=g)
conf[torch.randperm(K, generator=g)[:2]] = 0.0
init = {k: None for k in rp.PARAM_ORDER}
init.update(global_orient=pose[:, :3].contiguous(), body_pose=pose[:, 3:].contiguous(),
            betas=0.3 * torch.randn(n, 10, generator=g), transl=mo["transl"] + 0.03 * torch.randn(n, 3, generator=g))
f = WorldSpaceFitter(w, joints_category=layout, model_type=mt, gmm=gmm, use_lbfgs=False)
live = {k: v for k, v in init.items() if v is not None}
res = {kern: f.fit_batch(live, tgt, conf, seq_ind=0, num_iters=iters, with_mesh=False, kernel=kern) for kern in ("warp", "frame")}
ref = rp.fit_frame(BodyModelShim(w), prior, init, tgt, conf, seq_ind=0, num_obs=K, use_lbfgs=False, num_iters_first=iters, num_iters_followup=iters)
rb = ref["params"]["body_pose"].numpy()
for kern, o in res.items():
    d = np.abs(o["params"]["body_pose"].cpu().numpy() - rb).max(axis=1)
    print(kern, "worst frames", np.argsort(d)[-3:], np.sort(d)[-3:])
b = int(np.argmax(np.abs(res["warp"]["params"]["body_pose"].cpu().numpy() - rb).max(axis=1)))
rp_ = torch.cat([ref["params"]["global_orient"], ref["params"]["body_pose"], ref["params"]["transl"], ref["params"]["betas"]], 1).numpy()
def flat(o):
    p = o["params"]
    return torch.cat([p["global_orient"], p["body_pose"], p["transl"], p["betas"]], 1).cpu().numpy()
e = np.abs(flat(res["warp"])[b] - rp_[b])
print("frame", b, "worst elements", np.argsort(e)[-8:], np.sort(e)[-8:])
for H in ("0", "1", "2"):
    os.environ["K2B_CHAIN_HELPERS"] = H
    o = f.fit_batch(live, tgt, conf, seq_ind=0, num_iters=iters, with_mesh=False, kernel="warp")
    d = np.abs(flat(o) - rp_).max(axis=1)
    one = f.fit_batch({k: v[b:b + 1] for k, v in live.items()}, tgt[b:b + 1], conf, seq_ind=0, num_iters=iters, with_mesh=False, kernel="warp")
    d1 = np.abs(flat(one) - rp_[b:b + 1]).max()
    print("helpers", H, "batch worst", np.argsort(d)[-2:], np.sort(d)[-2:], "| frame alone", d1)
    for it in (1, 2, 3, 5):
        one = f.fit_batch({k: v[b:b + 1] for k, v in live.items()}, tgt[b:b + 1], conf, seq_ind=0, num_iters=it, with_mesh=False, kernel="warp")
        sub = {k: (v[b:b + 1] if v is not None else None) for k, v in init.items()}
        r1 = rp.fit_frame(BodyModelShim(w), prior, sub, tgt[b:b + 1], conf, seq_ind=0, num_obs=K, use_lbfgs=False, num_iters_first=it, num_iters_followup=it)["params"]
        r1 = torch.cat([r1["global_orient"], r1["body_pose"], r1["transl"], r1["betas"]], 1).numpy()
        e1 = np.abs(flat(one) - r1)[0]
        print("   helpers", H, "alone, iters", it, "max err", e1.max(), "at element", int(e1.argmax()))
