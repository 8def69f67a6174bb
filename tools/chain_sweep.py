"""Times the warp-per-sequence kernel for several team shapes (evaluators x helpers), L-BFGS, schedule S1.

    python tools/chain_sweep.py [SxT ...]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from keypoints2body_b200 import synthetic as syn  # noqa: E402
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter  # noqa: E402

w, gmm = syn.make_body_model("smpl"), syn.make_gmm()
cfgs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(256, 64), (1, 256), (148, 64), (512, 32)]
shapes = [(1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0)]
f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=True)
for S, Tn in cfgs:
    mo = syn.make_motion(S * Tn, seed=3)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).reshape(S, Tn, 22, 3).cuda()
    init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                transl=mo["transl"].reshape(S, Tn, 3)[:, 0].contiguous())
    init = {k: v.cuda() for k, v in init.items()}
    for E, H in shapes:
        os.environ["K2B_CHAIN_TEAM"], os.environ["K2B_CHAIN_HELPERS"] = str(E), str(H)
        try:
            f.fit_chain(init, tgt, None, with_mesh=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            o = f.fit_chain(init, tgt, None, with_mesh=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        except Exception as e:      # noqa: BLE001
            print(f"S={S} T={Tn} E={E} H={H}: {e}", flush=True)
            continue
        ev = o["evals"].float().reshape(S, Tn)
        print(f"S={S} T={Tn} E={E} H={H}: {dt*1e3:.2f} ms, {dt/Tn*1e6:.1f} us per frame-step, {S*Tn/dt/1e3:.1f} k frames/s, "
              f"evals follow {float(ev[:, 1:].mean()):.2f}, loss {float(o['loss'].mean()):.3f}", flush=True)
