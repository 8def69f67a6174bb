"""Scratch GPU diagnostics (test infrastructure: it may use the oracle; the product and tools/ do not).

    python tests/gpu_debug.py eval|time|prof|mesh|meshx|meshprof|rounds
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter
from oracle import reference_port as rp

g = dict(np.load("tests/golden/ref_goldens.npz"))
w = syn.make_body_model("smpl"); gmm = syn.make_gmm(0)
T = torch.as_tensor
what = sys.argv[1] if len(sys.argv) > 1 else "eval"
if what == "eval":
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm)
    tag = "eval_smpl_22_w0"
    params = {k: T(g[f"{tag}_in_{k}"]) for k in rp.PARAM_ORDER if f"{tag}_in_{k}" in g}
    for conf in (T(g[tag + "_in_conf"]), None, torch.ones(22)):
        out = f.evaluate_batch(params, T(g[tag + "_in_target"]), conf)
        print("conf", None if conf is None else conf[:3], "loss", out["loss"].cpu().numpy(), "ref", g[tag + "_loss"].reshape(-1))
        print(" joints err", np.abs(out["joints"].cpu().numpy() - g[tag + "_joints"][:, :22]).max(), "comp", out["gmm_component"].cpu().numpy())
    # same frames through the Adam kernel with 0 iterations -> loss not returned; use 1 iteration
    fa = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=False)
    o = fa.fit_batch(params, T(g[tag + "_in_target"]), T(g[tag + "_in_conf"]), seq_ind=0, num_iters=1, with_mesh=False)
    print("adam 1-iter loss (pre-step)", o["loss"].cpu().numpy())
if what == "time":
    for opt in ("adam", "lbfgs"):
        f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=opt == "lbfgs")
        for B in (384, 148 * 384, 148 * 384 * 4):
            mo = syn.make_motion(B, seed=3)
            tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).cuda()
            init = dict(global_orient=mo["pose"][:, :3].contiguous(), body_pose=mo["pose"][:, 3:].contiguous() * 0.9,
                        betas=torch.zeros(B, 10), transl=mo["transl"])
            init = {k: v.cuda() for k, v in init.items()}
            for iters in (10, 30):
                f.fit_batch(init, tgt, None, seq_ind=int(os.environ.get("K2B_SEQ", "1")), num_iters=iters, with_mesh=False)
                torch.cuda.synchronize(); t0 = time.perf_counter()
                o = f.fit_batch(init, tgt, None, seq_ind=int(os.environ.get("K2B_SEQ", "1")), num_iters=iters, with_mesh=False)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
                ev = float(o["evals"].float().mean())
                print(f"{opt} B={B} iters={iters} evals/frame={ev:.1f} time={dt*1e3:.2f} ms  -> {dt/ (ev+1) * 1e6:.1f} us per eval-round, {B*(ev+1)/dt/1e6:.2f} M frame-evals/s")
if what == "prof":
    B = 148 * 384
    opt = sys.argv[2] if len(sys.argv) > 2 else "adam"
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=opt == "lbfgs")
    mo = syn.make_motion(B, seed=3)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).cuda()
    init = dict(global_orient=mo["pose"][:, :3].contiguous(), body_pose=mo["pose"][:, 3:].contiguous() * 0.9,
                betas=torch.zeros(B, 10), transl=mo["transl"])
    init = {k: v.cuda() for k, v in init.items()}
    for _ in range(2):
        o = f.fit_batch(init, tgt, None, seq_ind=1, num_iters=int(os.environ.get("K2B_ITERS", "10")), with_mesh=False)
    torch.cuda.synchronize()
    print("ok", float(o["loss"].mean()))
if what == "mesh":
    from oracle.smplx_shim import BodyModelShim
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm)
    shim = BodyModelShim(w)
    for B in (37, 300, 128 * 5):
        gg = torch.Generator().manual_seed(B)
        params = dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                      betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg))
        out = f.forward_batch(params)
        torch.cuda.synchronize()
        ref = shim(**params)
        dv = (out["vertices"].cpu() - ref.vertices).abs().max().item()
        dj = (out["joints"].cpu() - ref.joints).abs().max().item()
        print(f"mesh B={B}: max|verts-shim|={dv:.3e} max|joints-shim|={dj:.3e}", flush=True)
    B = 1 << 18
    gg = torch.Generator().manual_seed(1)
    params = {k: v.cuda() for k, v in dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                  betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg)).items()}
    buf = torch.empty(B, 6890, 3, device="cuda")
    for _ in range(2):
        f.forward_batch(params, out_vertices=buf)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        f.forward_batch(params, out_vertices=buf)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    print(f"mesh B={B}: {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M frames/s, {B*6890*12/dt/1e9:.0f} GB/s of vertex output")
if what == "meshprof":
    w = syn.make_body_model("smpl", skin_layout=os.environ.get("K2B_SKIN_LAYOUT", "interleaved"))
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm)
    B = 148 * 128 * 2
    gg = torch.Generator().manual_seed(1)
    params = {k: v.cuda() for k, v in dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                  betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg)).items()}
    buf = torch.empty(B, 6890, 3, device="cuda")
    for _ in range(2):
        f.forward_batch(params, out_vertices=buf)
    torch.cuda.synchronize()
    print("ok")
if what == "rounds":
    os.environ["K2B_DEBUG_ROUNDS"] = "1"
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=True)
    B = 148 * 384
    mo = syn.make_motion(B, seed=3)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).cuda()
    for (iters, seq, scale) in ((30, 0, 0.0), (10, 1, 0.9)):
        init = dict(global_orient=mo["pose"][:, :3].contiguous() * scale, body_pose=mo["pose"][:, 3:].contiguous() * scale,
                    betas=torch.zeros(B, 10), transl=mo["transl"])
        init = {k: v.cuda() for k, v in init.items()}
        o = f.fit_batch(init, tgt, None, seq_ind=seq, num_iters=iters, with_mesh=False)
        e = o["evals"].cpu().numpy()
        ev, rd = e & 0xFFFF, e >> 16
        print(f"iters={iters}: evals/frame mean {ev.mean():.2f} max {ev.max()}, rounds/warp mean {rd.mean():.2f} max {rd.max()} -> rounds/evals = {rd.mean()/ev.mean():.3f}")
if what == "meshx":
    # SMPL-H / SMPL-X full-mesh throughput: tcgen05 blend (64 frames per pass) vs the fused CUDA-core kernel
    for mt, nv in (("smplh", 6890), ("smplx", 10475)):
        wx = syn.make_body_model(mt)
        f = WorldSpaceFitter(wx, joints_category="AMASS", model_type=mt, gmm=gmm)
        B = 1 << 16
        gg = torch.Generator().manual_seed(1)
        params = dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                      betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg),
                      left_hand_pose=0.2 * torch.randn(B, 45, generator=gg), right_hand_pose=0.2 * torch.randn(B, 45, generator=gg))
        if mt == "smplx":
            params.update(expression=torch.randn(B, 10, generator=gg), jaw_pose=0.2 * torch.randn(B, 3, generator=gg),
                          leye_pose=0.2 * torch.randn(B, 3, generator=gg), reye_pose=0.2 * torch.randn(B, 3, generator=gg))
        params = {k: v.cuda() for k, v in params.items()}
        buf = torch.empty(B, nv, 3, device="cuda")
        res = {}
        for fp32 in ("0", "1"):
            os.environ["K2B_MESH_FP32"] = fp32
            for _ in range(2):
                f.forward_batch(params, out_vertices=buf)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3):
                f.forward_batch(params, out_vertices=buf)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            res[fp32] = buf[::997].clone()
            print(f"{mt} mesh B={B} {'cuda-core' if fp32 == '1' else 'tcgen05  '}: {dt*1e3:.2f} ms -> {B/dt/1e6:.3f} M frames/s, "
                  f"{B*nv*12/dt/1e9:.0f} GB/s of vertex output", flush=True)
        print(f"  max |tc - fp32| = {(res['0'] - res['1']).abs().max().item():.3e} m")
if what == "chain":
    # warp-per-sequence kernel: time S sequences x T frames (schedule S1) and the lone-frame latency
    cfgs = [(1, 1), (1, 64), (148, 64), (256, 64), (1024, 32), (2368, 16), (8192, 8)]
    if len(sys.argv) > 2:
        cfgs = [tuple(int(v) for v in a.split("x")) for a in sys.argv[2:]]
    for opt in ("adam", "lbfgs"):
        f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=opt == "lbfgs")
        for S, Tn in cfgs:
            mo = syn.make_motion(S * Tn, seed=3)
            tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).reshape(S, Tn, 22, 3).cuda()
            init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                        transl=mo["transl"].reshape(S, Tn, 3)[:, 0].contiguous())
            init = {k: v.cuda() for k, v in init.items()}
            f.fit_chain(init, tgt, None, with_mesh=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            o = f.fit_chain(init, tgt, None, with_mesh=False)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            ev = o["evals"].float().reshape(S, Tn)
            err = float((o["fit_joints"].reshape(S, Tn, 22, 3) - tgt).norm(dim=-1).mean())
            print(f"chain {opt} S={S} T={Tn}: {dt*1e3:.2f} ms, {S*Tn/dt/1e3:.1f} k frames/s, "
                  f"{dt/Tn*1e6:.1f} us per frame-step, evals first/follow {float(ev[:, 0].mean()):.1f}/"
                  f"{float(ev[:, 1:].mean()) if Tn > 1 else 0:.1f}, mean joint err {err*1e3:.2f} mm", flush=True)
    import keypoints2body_b200 as k2b, tempfile
    d = tempfile.mkdtemp(); syn.write_assets(d + "/data/models", seed=0); os.chdir(d)
    frame = g["seq_in_target"][0]
    for cfg in (None, dict(use_lbfgs=False)):
        for kern in ("1024", "0"):
            os.environ["K2B_WARP_MAX_FRAMES"] = kern
            r = k2b.optimize_params_frame(frame, body_model="smpl", joint_layout="AMASS", model=w, config=cfg)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                r = k2b.optimize_params_frame(frame, body_model="smpl", joint_layout="AMASS", model=w, config=cfg)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
            print(f"optimize_params_frame {'lbfgs' if cfg is None else 'adam'} kernel={'warp' if kern != '0' else 'frame'}: "
                  f"{dt*1e3:.2f} ms per call, loss {float(r.loss):.1f}", flush=True)
if what == "chainprof":
    opt = sys.argv[2] if len(sys.argv) > 2 else "adam"
    S, Tn = (int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "256x16").split("x"))
    f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=opt == "lbfgs")
    mo = syn.make_motion(S * Tn, seed=3)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).reshape(S, Tn, 22, 3).cuda()
    init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                transl=mo["transl"].reshape(S, Tn, 3)[:, 0].contiguous())
    init = {k: v.cuda() for k, v in init.items()}
    for _ in range(2):
        o = f.fit_chain(init, tgt, None, with_mesh=False)
    torch.cuda.synchronize()
    print("ok", float(o["loss"].mean()))
if what == "cross":
    # where does the warp-per-frame kernel stop beating the thread-per-frame kernel?
    for opt in ("lbfgs", "adam"):
        f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=opt == "lbfgs")
        for B in (1024, 4096, 8192, 16384, 32768, 65536):
            mo = syn.make_motion(B, seed=3)
            tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).cuda()
            init = dict(global_orient=torch.zeros(B, 3), body_pose=torch.zeros(B, 69), betas=torch.zeros(B, 10), transl=mo["transl"])
            init = {k: v.cuda() for k, v in init.items()}
            for seq in (0, 1):
                row = []
                for kern in ("frame", "warp"):
                    f.fit_batch(init, tgt, None, seq_ind=seq, with_mesh=False, kernel=kern)
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    for _ in range(3):
                        f.fit_batch(init, tgt, None, seq_ind=seq, with_mesh=False, kernel=kern)
                    torch.cuda.synchronize(); row.append((time.perf_counter() - t0) / 3 * 1e3)
                print(f"cross {opt} B={B} seq_ind={seq}: frame {row[0]:.2f} ms, warp {row[1]:.2f} ms", flush=True)

if what == "meshc":
    # full-mesh throughput with the interleaved (worst case) and the coherent (realistic) vertex -> joint assignment
    from oracle.smplx_shim import BodyModelShim
    for mt, nv in (("smpl", 6890), ("smplh", 6890), ("smplx", 10475)):
        for layout in ("interleaved", "coherent"):
            wx = syn.make_body_model(mt, skin_layout=layout)
            f = WorldSpaceFitter(wx, joints_category="AMASS", model_type=mt, gmm=gmm)
            B = 1 << 16
            gg = torch.Generator().manual_seed(1)
            params = dict(global_orient=0.3 * torch.randn(B, 3, generator=gg), body_pose=0.3 * torch.randn(B, 69, generator=gg),
                          betas=torch.randn(B, 10, generator=gg), transl=torch.randn(B, 3, generator=gg))
            if mt != "smpl":
                params.update(left_hand_pose=0.2 * torch.randn(B, 45, generator=gg), right_hand_pose=0.2 * torch.randn(B, 45, generator=gg))
            if mt == "smplx":
                params.update(expression=torch.randn(B, 10, generator=gg), jaw_pose=0.2 * torch.randn(B, 3, generator=gg),
                              leye_pose=0.2 * torch.randn(B, 3, generator=gg), reye_pose=0.2 * torch.randn(B, 3, generator=gg))
            ref = BodyModelShim(wx)(**{k: v[:200] for k, v in params.items()})
            params = {k: v.cuda() for k, v in params.items()}
            buf = torch.empty(B, nv, 3, device="cuda")
            for _ in range(2):
                out = f.forward_batch(params, out_vertices=buf)
            dv = (out["vertices"][:200].cpu() - ref.vertices).abs().max().item()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3):
                f.forward_batch(params, out_vertices=buf)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            print(f"{mt} {layout}: B={B} {dt*1e3:.2f} ms -> {B/dt/1e6:.2f} M frames/s, {B*nv*12/dt/1e9:.0f} GB/s of vertex output, max|verts-shim|={dv:.2e}", flush=True)
if what == "camseq":
    # a lone 128-frame camera-space sequence through the public API: one launch vs two launches per frame
    import keypoints2body_b200 as k2b
    import tempfile
    from oracle.problems import chain_problem
    tmp = tempfile.mkdtemp(); syn.write_assets(os.path.join(tmp, "data/models"), seed=0); os.chdir(tmp)
    tgt = chain_problem(w, 1, 128, seed=5)[0].numpy()
    for opt in ("adam", "lbfgs"):
        cfg = dict(frame=dict(use_lbfgs=opt == "lbfgs", coordinate_mode="camera", num_iters=10), use_shape_optimization=False)
        for mode in ("0", "1"):
            os.environ["K2B_CAMERA_LAUNCH_PER_FRAME"] = mode
            k2b.optimize_params_sequence(tgt, body_model="smpl", joint_layout="AMASS", model=w, config=cfg)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = k2b.optimize_params_sequence(tgt, body_model="smpl", joint_layout="AMASS", model=w, config=cfg)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            print(f"camera sequence {opt} 128 frames, {'launch per frame' if mode == '1' else 'one launch'}: {dt*1e3:.1f} ms, "
                  f"{128/dt:.0f} frames/s, median loss {np.median([float(r.loss) for r in res]):.1f}", flush=True)
if what == "chaincycles":
    # cycle accounting of the leading evaluator (library built with EXTRA=-DK2B_CHAIN_PROF)
    import ctypes as C
    from keypoints2body_b200 import _native as nat
    lib = nat.load_library()
    names = ["rounds", "next point + post", "line tables", "evaluation", "wait for the team", "machine", "outer update",
             "rest of the round", "frames"]
    for cfg in sys.argv[2:] or ["256x64"]:
        S, Tn = (int(v) for v in cfg.split("x"))
        f = WorldSpaceFitter(w, joints_category="AMASS", model_type="smpl", gmm=gmm, use_lbfgs=True)
        mo = syn.make_motion(S * Tn, seed=3)
        tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22).reshape(S, Tn, 22, 3).cuda()
        init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                    transl=mo["transl"].reshape(S, Tn, 3)[:, 0].contiguous())
        init = {k: v.cuda() for k, v in init.items()}
        f.fit_chain(init, tgt, None, with_mesh=False)
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 16)()
        lib.k2b_chain_prof(buf, 1)
        t0 = time.perf_counter()
        f.fit_chain(init, tgt, None, with_mesh=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        lib.k2b_chain_prof(buf, 1)
        v = list(buf)
        rounds, frames = v[0], v[8]
        cyc = sum(v[1:8])
        print(f"S={S} T={Tn}: {dt*1e3:.2f} ms, {dt/Tn*1e6:.1f} us per frame-step; {rounds/frames:.2f} rounds per frame, "
              f"{cyc/frames/1.965e3:.1f} us of leader cycles per frame")
        for i in range(1, 8):
            print(f"   {names[i]:20s} {v[i]/frames/1.965e3:7.2f} us per frame  {100*v[i]/cyc:5.1f} %  ({v[i]/rounds:7.0f} cycles per round)")
        if v[9]:
            print(f"   table updates per frame {v[9]/frames:.2f}: own share {v[10]/v[9]:.0f} cycles, wait for the team {v[11]/v[9]:.0f} cycles per update")
