"""What the window boundaries of fit_chain cost: S sequences x T frames, fit only, 1 launch vs C windows (each window
is a launch that ends when its slowest sequence does).

    python tools/window_tail.py [S] [T] [chunks ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from keypoints2body_b200 import synthetic as syn  # noqa: E402
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
chunk_list = [int(a) for a in sys.argv[3:]] or [1, 4, 16, 64]
w, gmm = syn.make_body_model("smpl"), syn.make_gmm()
for lb in (True, False):
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=lb)
    mo = syn.make_motion(S * T, seed=3)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).reshape(S, T, 22, 3).cuda()
    init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                transl=mo["transl"].reshape(S, T, 3)[:, 0].contiguous())
    init = {k: v.cuda() for k, v in init.items()}
    for C in chunk_list:
        kw = dict(with_mesh=False, time_major=True, chunks=C, fit_joints=False)
        f.fit_chain(init, tgt, None, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        o = f.fit_chain(init, tgt, None, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{'lbfgs' if lb else 'adam'} S={S} T={T} windows={C}: {ms:.2f} ms, {ms / T * 1e3:.2f} us per frame-step, "
              f"{S * T / ms / 1e3:.3f} M frames/s, loss {float(o['loss'].mean()):.3f}", flush=True)
