"""Supplementary measurements of the other SURVEY 8(d) configurations (bench.py covers the headline one).

    python tools/bench_configs.py > profiles/rNN_configs.jsonl

One JSON line per configuration, timed with CUDA events after warm-up on one B200:
  config1  optimize_params_frame, SMPL AMASS, defaults (world, L-BFGS, 30 its): B = 1 latency
  config2  optimize_params_sequence, T = 4096, schedules S0 / S1 / S2, both optimisers
  config3  SMPL-X, 65 536 independent frames, num_iters_first = 5, freeze_betas = False, L-BFGS and Adam
  config5  SMPL-X / SMPL-H / SMPL full-mesh forward of fitted parameter sets (vertices + joints)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import keypoints2body_b200 as k2b
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.core.config import FrameOptimizeConfig, SequenceOptimizeConfig
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter

dev = torch.device("cuda")


def cuda_ms(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def wall_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    syn.write_assets("/tmp/k2b_assets/data/models", seed=0)   # the API keeps the reference's CWD-relative asset paths
    os.chdir("/tmp/k2b_assets")
    w = syn.make_body_model("smpl")
    gmm = syn.make_gmm(0)
    mo = syn.make_motion(4096, seed=1)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    tgt = (tgt + 0.005 * torch.randn(tgt.shape, generator=torch.Generator().manual_seed(2))).numpy()

    # ---- config 1: single-frame latency through the public API --------------------------------------
    for lb in (True, False):
        ms = wall_ms(lambda: k2b.optimize_params_frame(tgt[0], body_model="smpl", joint_layout="AMASS", model=w,
                                                       config=dict(use_lbfgs=lb)))
        emit(config="config1 optimize_params_frame SMPL AMASS B=1 (30 its, full mesh)", optimizer="lbfgs" if lb else "adam",
             ms_per_call=ms, frames_per_s=1e3 / ms, timing="wall clock incl. Python, model/prior construction per call")

    # ---- config 2: one 4096-frame sequence, all schedules --------------------------------------------
    for lb in (True, False):
        for name, kw in (("S2 two_sweep", dict(schedule="two_sweep")),
                         ("S0 independent", dict(schedule="reference", use_previous_frame_init=False)),
                         ("S1 serial chain (reference default)", dict(schedule="reference", use_previous_frame_init=True))):
            T = 4096       # S1 runs as one launch of the warp-per-sequence kernel (k2b_fit_chain)
            cfg = SequenceOptimizeConfig(frame=FrameOptimizeConfig(use_lbfgs=lb), use_shape_optimization=False, **kw)
            ms = wall_ms(lambda: k2b.optimize_params_sequence(tgt[:T], body_model="smpl", joint_layout="AMASS", model=w,
                                                              config=cfg), warm=1, reps=2)
            emit(config=f"config2 optimize_params_sequence SMPL T={T} {name}", optimizer="lbfgs" if lb else "adam",
                 ms_per_call=ms, frames_per_s=T * 1e3 / ms, timing="wall clock through the public API (results as per-frame objects)")

    # ---- row f1: camera-space two-stage fitter through the public API (B = 1 per frame by construction) -----
    for lb in (True, False):
        cfgc = dict(use_lbfgs=lb, coordinate_mode="camera", num_iters=30)
        ms = wall_ms(lambda: k2b.optimize_params_frame(tgt[0], body_model="smpl", joint_layout="AMASS", model=w, config=cfgc))
        emit(config="f1 camera-space optimize_params_frame SMPL AMASS B=1 (2 stages x 30 its, full mesh)",
             optimizer="lbfgs" if lb else "adam", ms_per_call=ms, frames_per_s=1e3 / ms,
             timing="wall clock incl. Python, model/prior construction per call")
        Tc = 128
        scfg = SequenceOptimizeConfig(frame=FrameOptimizeConfig(use_lbfgs=lb, coordinate_mode="camera", num_iters=30),
                                      use_shape_optimization=False)
        ms = wall_ms(lambda: k2b.optimize_params_sequence(tgt[:Tc], body_model="smpl", joint_layout="AMASS", model=w,
                                                          config=scfg), warm=1, reps=2)
        emit(config=f"f1 camera-space optimize_params_sequence SMPL T={Tc} (one launch: k2b_fit_chain camera_sequence)",
             optimizer="lbfgs" if lb else "adam", ms_per_call=ms, frames_per_s=Tc * 1e3 / ms,
             timing="wall clock through the public API")

    # ---- config 3: SMPL-X, 65 536 independent frames, 5 iterations -------------------------------------
    wx = syn.make_body_model("smplx")
    B = 65536
    mx = syn.make_motion(B, seed=4)
    tx = syn.kinematic_joints(wx, mx["pose"][:, :66], mx["betas"], mx["transl"], 22).to(dev)
    g = torch.Generator().manual_seed(6)
    init = dict(global_orient=mx["pose"][:, :3].contiguous(), body_pose=(mx["pose"][:, 3:] + 0.05 * torch.randn(B, 69, generator=g)),
                betas=torch.zeros(B, 10), transl=mx["transl"], left_hand_pose=torch.zeros(B, 45), right_hand_pose=torch.zeros(B, 45),
                expression=torch.zeros(B, 10), jaw_pose=torch.zeros(B, 3), leye_pose=torch.zeros(B, 3), reye_pose=torch.zeros(B, 3))
    init = {k: v.to(dev).contiguous() for k, v in init.items()}
    for lb in (True, False):
        f = WorldSpaceFitter(wx, joints_category="AMASS", model_type="smplx", gmm=gmm, use_lbfgs=lb, device=dev)
        for mesh in (False, True):
            out = {}
            def run():
                out.update(f.fit_batch(init, tx, None, seq_ind=0, num_iters=5, freeze_betas=False, with_mesh=mesh))
            ms = cuda_ms(run)
            emit(config=f"config3 SMPL-X {B} independent frames, 5 iterations, freeze_betas=False, " + ("full mesh" if mesh else "fit only"),
                 optimizer="lbfgs" if lb else "adam", ms=ms, frames_per_s=B * 1e3 / ms,
                 evals_per_frame=float(out["evals"].float().mean()), timing="CUDA events, inputs resident")

    # ---- config 5: full-mesh forward of fitted parameter sets -------------------------------------------
    for mt, nv in (("smpl", 6890), ("smplh", 6890), ("smplx", 10475)):
        wm = syn.make_body_model(mt)
        f = WorldSpaceFitter(wm, joints_category="AMASS", model_type=mt, gmm=gmm, device=dev)
        B = 1 << 17
        gg = torch.Generator().manual_seed(1)
        p = dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                 betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg))
        if mt != "smpl":
            p.update(left_hand_pose=0.2 * torch.randn(B, 45, generator=gg), right_hand_pose=0.2 * torch.randn(B, 45, generator=gg))
        if mt == "smplx":
            p.update(expression=torch.randn(B, 10, generator=gg), jaw_pose=0.2 * torch.randn(B, 3, generator=gg),
                     leye_pose=0.2 * torch.randn(B, 3, generator=gg), reye_pose=0.2 * torch.randn(B, 3, generator=gg))
        p = {k: v.to(dev) for k, v in p.items()}
        buf = torch.empty(B, nv, 3, device=dev)
        ms = cuda_ms(lambda: f.forward_batch(p, out_vertices=buf))
        emit(config=f"config5 {mt} full mesh of {B} parameter sets ({nv} vertices)", ms=ms, frames_per_s=B * 1e3 / ms,
             vertex_output_GBps=B * nv * 12 / ms / 1e6, hbm_frac_of_measured=B * nv * 12 / ms / 1e6 / 6551.7,
             timing="CUDA events, chunk of 131072 frames (1M frames = 8 such chunks)")
        if mt == "smplx":       # BASELINE configs[4] as written: 1 M fitted SMPL-X frames (8 chunks into one 15.7 GB window)
            def million():
                for _ in range(8):
                    f.forward_batch(p, out_vertices=buf)
            ms = cuda_ms(million, warm=1, reps=2)
            emit(config=f"config5 smplx full mesh of {8 * B} parameter sets (8 chunks of {B})", ms=ms, frames_per_s=8 * B * 1e3 / ms,
                 vertex_output_GBps=8 * B * nv * 12 / ms / 1e6, hbm_frac_of_measured=8 * B * nv * 12 / ms / 1e6 / 6551.7,
                 timing="CUDA events")
        wc_ = syn.make_body_model(mt, skin_layout="coherent")
        fc_ = WorldSpaceFitter(wc_, joints_category="AMASS", model_type=mt, gmm=gmm, device=dev)
        ms = cuda_ms(lambda: fc_.forward_batch(p, out_vertices=buf))
        emit(config=f"config5 {mt} full mesh of {B} parameter sets, coherent vertex -> joint assignment", ms=ms,
             frames_per_s=B * 1e3 / ms, vertex_output_GBps=B * nv * 12 / ms / 1e6,
             hbm_frac_of_measured=B * nv * 12 / ms / 1e6 / 6551.7, timing="CUDA events")

    # ---- rows f2 / f4: the general articulated fit through its fitters (B = 1 per call, like the reference) ----------
    from keypoints2body_b200.core.fitters.misc_models import FLAMEFitter, MANOFitter
    from keypoints2body_b200.models.smpl_data import FLAMEData, MANOData, SMPLHData, SMPLXData
    from oracle_free_problem import articulated_problem
    for mt in ("smplx", "smplh", "mano", "flame"):
        wmod, tgt_a, idx_a, init_a = articulated_problem(mt, 1, seed=700)
        cls = {"smplx": SMPLXData, "smplh": SMPLHData, "mano": MANOData, "flame": FLAMEData}[mt]
        for lb in (True, False):
            if mt in ("smplx", "smplh"):
                fit = WorldSpaceFitter(wmod, joints_category="GENERIC", model_type=mt, gmm=gmm, use_lbfgs=lb, device=dev)
            else:
                fit = (MANOFitter if mt == "mano" else FLAMEFitter)(wmod, coordinate_mode="world", use_lbfgs=lb, device=dev)
            p0 = cls(**{k: v.to(dev) for k, v in init_a.items()})
            ms = wall_ms(lambda: fit.fit_frame(p0, tgt_a, torch.ones(len(idx_a)), seq_ind=0, target_model_indices=idx_a), warm=2, reps=10)
            emit(config=f"f2/f4 {mt} fit_frame, {len(idx_a)} observed model points (30 its, full mesh), general articulated fit",
                 optimizer="lbfgs" if lb else "adam", ms_per_call=ms, frames_per_s=1e3 / ms, timing="wall clock incl. Python")


if __name__ == "__main__":
    main()
