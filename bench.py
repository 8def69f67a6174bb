#!/usr/bin/env python
"""bench.py -- frames fitted / second (SMPL, AMASS-22, reference iteration schedule).

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # CPU oracle port on the host cores

Workload (BASELINE.json configs[3] sharded; every sequence is configs[1]): per GPU
`--frames-per-gpu` frames (default 1 048 576 = 256 sequences x 4 096 frames; at 8 GPUs that is the
8M-frame config) of synthetic AMASS-22 keypoints generated from the synthetic SMPL model, fitted with the
reference default optimiser (L-BFGS / strong Wolfe), followed by the full-mesh output (6 890 vertices +
45 joints per frame).  One step = one pass over the batch.

--schedule reference (default) = S1, what optimize_params_sequence does by default (api/sequence.py:214-281):
    frame 0 gets 30 iterations, every later frame starts from the previous frame's result and gets 10
    iterations with the temporal pose-preserve term.  Serial in t, so each sequence is walked by one warp
    group inside the warp-per-sequence kernel; GPUs take whole sequences, no data-path collective.
--schedule two_sweep = S2, the frame-parallel variant (sweep 0: 30-iteration budget from the mean pose,
    sweep 1: 10-iteration budget from the neighbour's sweep-0 result): frames shard across GPUs, the sequence
    grid is shifted by half a sequence so every shard boundary falls inside a sequence and the one-frame NCCL
    halo exchange is really used.  Same throughput on one GPU, but it does 4x the evaluations and ends 5x
    further from the keypoints than the chain (18 cm vs 3.4 cm mean joint error on this workload).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames fitted/sec (SMPL, AMASS-22, fixed iters)"
UNIT = "frames/s"
SEQ_LEN = 4096
EVAL_FLOP = {"smpl": 98e3, "smplh": 110e3, "smplx": 115e3}   # algorithmic flop / evaluation, SURVEY.md 8(d)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner under torchrun)
    is sent to stderr instead."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=256 * SEQ_LEN)
    ap.add_argument("--optimizer", default="lbfgs", choices=["lbfgs", "adam"])
    ap.add_argument("--schedule", default="reference", choices=["reference", "two_sweep"],
                    help="reference = S1, the reference's own serial chain (frame t starts from frame t-1's result), "
                         "one warp group per sequence, whole sequences per GPU; two_sweep = S2 (frame-parallel "
                         "variant, frames sharded across GPUs with a one-frame NCCL halo exchange)")
    ap.add_argument("--chunks", type=int, default=16, help="schedule reference: time windows whose mesh pass overlaps the next window's fit")
    ap.add_argument("--no-vertices", action="store_true", help="skip the vertex output (joints only)")
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="frames of the first sequence the CPU arm fits (0 = 64)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--fp-steps", type=int, default=2, help="timed steps of the frame-parallel schedule measured beside S1")
    ap.add_argument("--no-frame-parallel", action="store_true")
    ap.add_argument("--no-e2e-vertices", action="store_true")
    return ap.parse_args()


def seq_index(lo: int, hi: int, device) -> torch.Tensor:
    """Index of global frames [lo, hi) inside their sequence: a first half-length sequence, then
    back-to-back sequences of SEQ_LEN frames (so multiples of 2^20 frames fall mid-sequence)."""
    g = torch.arange(lo, hi, device=device, dtype=torch.int64)
    half = SEQ_LEN // 2
    return torch.where(g < half, g, (g - half) % SEQ_LEN).to(torch.int32)


def make_targets(weights, lo: int, hi: int, device, chunk=1 << 16) -> torch.Tensor:
    """Synthetic keypoints of global frames [lo, hi): smooth random motions through the synthetic model
    (SURVEY.md 8(d)); deterministic per 65 536-frame chunk of the global frame axis."""
    from keypoints2body_b200 import synthetic as syn

    out = torch.empty(hi - lo, 22, 3, device=device)
    c0 = lo // chunk
    pos = lo
    while pos < hi:
        c = pos // chunk
        mo = syn.make_motion(SEQ_LEN, seed=1000 + c, num_sequences=chunk // SEQ_LEN)
        a, b = pos - c * chunk, min(hi, (c + 1) * chunk) - c * chunk
        sl = {k: v[a:b].to(device) for k, v in mo.items()}
        j = syn.kinematic_joints(weights, sl["pose"][:, :66], sl["betas"], sl["transl"], 22)
        g = torch.Generator(device="cpu").manual_seed(5000 + c)
        noise = 0.005 * torch.randn(chunk, 22, 3, generator=g)[a:b].to(device)
        out[pos - lo: pos - lo + (b - a)] = j + noise
        pos += b - a
    del c0
    return out


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_FRAMES = 64      # frames of the workload's first sequence the CPU arm fits per step (>= 64)


def workload_config(args, mesh=True) -> dict:
    """The workload definition, identical in both arms (`bench.py` and `bench.py --impl reference`)."""
    F = args.frames_per_gpu
    chain = args.schedule == "reference"
    return {
        "workload": (f"SMPL AMASS-22 sequence fit, {F} frames/GPU ({F // SEQ_LEN} sequences x {SEQ_LEN}; BASELINE "
                     "configs[3] shard, each sequence = configs[1]), "
                     + ("schedule S1 = the reference's own serial chain (frame t starts from frame t-1's result; 30 "
                        "iterations for frame 0, 10 + pose-preserve after), whole sequences per GPU, " if chain else
                        "schedule S2: sweep0 30-iteration budget + sweep1 10-iteration budget with pose-preserve, frames "
                        "sharded across GPUs with a one-frame halo, ")
                     + f"{args.optimizer}, " + ("full mesh (6890 verts + 45 joints) per frame" if mesh else "joints only")),
        "optimizer": args.optimizer, "schedule": args.schedule, "frames_per_gpu": F, "seq_len": SEQ_LEN,
        "l2_policy": "inputs larger than L2 (targets %.0f MB/GPU, outputs %.1f GB/GPU)"
                     % (F * 22 * 12 / 1e6, F * (6890 * 12 + 45 * 12) / 1e9 if mesh else F * 45 * 12 / 1e9),
    }


class LbfgsEvalCounter:
    """Counts closure evaluations of every torch.optim.LBFGS.step (both the reference and the port run torch's)."""

    def __init__(self):
        self.evals = []

    def __enter__(self):
        import torch.optim.lbfgs as L

        self.L, self.orig = L, L.LBFGS.step
        me = self

        def step(opt, closure):
            out = me.orig(opt, closure)
            me.evals.append(int(opt.state[opt._params[0]]["func_evals"]))
            return out

        L.LBFGS.step = step
        return self

    def __exit__(self, *exc):
        self.L.LBFGS.step = self.orig
        return False


class CpuArm:
    """The reference's implementation of the path on the host cores.

    kind "reference": the UNMODIFIED reference package (/root/reference in the authoring container, its copy
    oracle/_ref on the GPU box -- oracle/make_ref.py) through its public ``optimize_params_sequence``, body model =
    oracle/smplx_shim (smplx itself is not installable offline).  kind "port": oracle/reference_port.py, used only
    when neither exists.  B = 1 per frame and a full-mesh forward per evaluation, exactly like the reference."""

    def __init__(self, optimizer: str):
        from keypoints2body_b200 import synthetic as syn
        from oracle import ref_loader
        from oracle.smplx_shim import BodyModelShim

        self.lbfgs = optimizer == "lbfgs"
        self.weights = syn.make_body_model("smpl", seed=0)
        self.model = BodyModelShim(self.weights)
        self.gmm = syn.make_gmm(seed=0)
        self.kind = "port"
        self.ref = None
        if ref_loader.available():
            try:
                import tempfile

                self.ref = ref_loader.load_reference()
                self.cwd = tempfile.mkdtemp()
                syn.write_assets(os.path.join(self.cwd, "data/models"), seed=0)
                self.kind = "reference"
            except Exception as e:        # noqa: BLE001
                print("reference import failed, timing the port:", e, file=sys.stderr)
                self.ref = None

    def fit_chain(self, joints: torch.Tensor, threads: int) -> dict:
        """Schedule S1 on one sequence (T,22,3): frames/s, evaluations per frame, per-frame loss and joints."""
        from oracle import problems, ref_loader
        from oracle import reference_port as rp

        torch.set_num_threads(threads)
        T = joints.shape[0]
        with LbfgsEvalCounter() as cnt:
            t0 = time.perf_counter()
            if self.ref is not None:
                cfg = dict(frame=dict(use_lbfgs=self.lbfgs), use_shape_optimization=False)
                with ref_loader.reference_cwd(self.cwd):
                    res = self.ref.optimize_params_sequence(joints.numpy(), body_model="smpl", joint_layout="AMASS",
                                                            model=self.model, config=cfg)
                loss = torch.stack([r.loss.reshape(()) for r in res])
                j22 = torch.cat([r.joints[:, :22] for r in res])
                pose = torch.cat([r.params.pose for r in res])
            else:
                prior = rp.GMMPrior(self.gmm)
                z = dict(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10))
                prev = {k: None for k in rp.PARAM_ORDER}
                prev.update(z, transl=joints[0:1, 0] - self.model(**z).joints[:, 0])
                loss, j22, pose = [], [], []
                for t in range(T):
                    r = rp.fit_frame(self.model, prior, prev, joints[t:t + 1], torch.ones(22), seq_ind=t, use_lbfgs=self.lbfgs)
                    prev = r["params"]
                    loss.append(r["loss"].reshape(()))
                    j22.append(r["joints"][:, :22])
                    pose.append(torch.cat([prev["global_orient"], prev["body_pose"]], dim=1))
                loss, j22, pose = torch.stack(loss), torch.cat(j22), torch.cat(pose)
            dt = time.perf_counter() - t0
        evals = cnt.evals if self.lbfgs else [30] + [10] * (T - 1)
        return {"seconds": dt, "frames_per_s": T / dt, "evals_per_frame": float(sum(evals)) / T, "loss": loss,
                "joints22": j22, "pose": pose, "err": problems.mean_joint_error(j22, joints)}

    def describe(self, frames: int) -> str:
        return (f"{frames} frames of the workload's first sequence, schedule S1 (serial chain, 30 then 10 iterations), "
                f"{'lbfgs' if self.lbfgs else 'adam'}, B=1 per frame, "
                + ("the unmodified reference's optimize_params_sequence" if self.kind == "reference" else "oracle port")
                + f", torch {torch.__version__} CPU, full-mesh forward per evaluation")


def cpu_sample_targets(args) -> torch.Tensor:
    """The first CPU_SAMPLE_FRAMES frames of global sequence 0 -- the frames rank 0's first sequence starts with."""
    from keypoints2body_b200 import synthetic as syn

    n = args.cpu_sample_frames or CPU_SAMPLE_FRAMES
    return make_targets(syn.make_body_model("smpl", seed=0), 0, n, torch.device("cpu"))


def cpu_baseline_block(arm: CpuArm, sample: torch.Tensor, threads: int, all_thread_run: dict, with_demo: bool) -> dict:
    one = arm.fit_chain(sample[: max(16, len(sample) // 2)], 1)          # 1 thread: half the sample keeps the run short
    blk = {"value": all_thread_run["frames_per_s"], "unit": UNIT, "cores": threads, "kind": arm.kind,
           "sample": arm.describe(len(sample)), "evals_per_frame": all_thread_run["evals_per_frame"],
           "value_1_thread": one["frames_per_s"], "sample_1_thread": f"{len(one['loss'])} frames"}
    if with_demo:
        try:       # the reference's own demo sequence (real AMASS-22 keypoints; stored with the goldens)
            import numpy as np

            demo = torch.as_tensor(np.load(os.path.join(ROOT, "tests", "golden", "r2_adam.npz"))["demo1_in"])
            d = arm.fit_chain(demo, threads)
            blk["demo_sequence"] = {"value": d["frames_per_s"], "unit": UNIT, "frames": len(demo),
                                    "evals_per_frame": d["evals_per_frame"],
                                    "what": "data/demo/test_motion1.npy (195 real AMASS-22 frames), same call"}
        except Exception as e:      # noqa: BLE001
            blk["demo_sequence"] = {"unavailable": str(e)}
    return blk


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.schedule != "reference":
        raise SystemExit("--impl reference times schedule S1 (the reference has no other)")
    threads = os.cpu_count() or 1
    os.environ.pop("OMP_NUM_THREADS", None)        # torchrun pins it to 1; torch.set_num_threads decides here
    arm = CpuArm(args.optimizer)
    sample = cpu_sample_targets(args)
    n = len(sample)
    # B = 1 work is a stream of tiny ops: more threads are not always faster.  The arm gets the better of "all
    # host threads" and "one thread" (probe on 16 frames), and reports both.
    arm.fit_chain(sample[:4], threads)
    probe = {th: arm.fit_chain(sample[:16], th)["frames_per_s"] for th in (threads, 1)}
    if probe[1] > probe[threads]:
        threads = 1
    for _ in range(args.warmup):
        arm.fit_chain(sample, threads)
    runs = [arm.fit_chain(sample, threads) for _ in range(args.steps)]
    dt = sum(r["seconds"] for r in runs) / len(runs)
    val = n / dt
    last = dict(runs[-1], frames_per_s=val)
    blk = cpu_baseline_block(arm, sample, threads, last, with_demo=True)
    blk["probe_frames_per_s"] = {"all_threads_%d" % (os.cpu_count() or 1): probe[os.cpu_count() or 1], "one_thread": probe[1]}
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": blk,
        "cpu_baseline_batched_adam": cpu_reference_batched_adam(1024, threads),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def mesh_overlap_report(fitter) -> dict:
    """What fit_chain's event-timed overlap policy measured during the warm-up steps and chose for the timed ones: the
    share of the time windows whose mesh pass is held to the SMs the fit leaves free (rank 0's view)."""
    rep = {}
    for (S, T, chunks, lb), v in fitter.overlap_policy().items():
        rep["%dx%d/%d/%s" % (S, T, chunks, "lbfgs" if lb else "adam")] = {
            "capped_fraction": v["fraction"], "ms_per_candidate": {str(k): round(m, 2) for k, m in v["ms"].items()}}
    return rep


def cpu_reference_batched_adam(frames: int, threads: int, seed: int = 78) -> dict:
    """SURVEY 8(d) "reference, batched by hand": the reference's Adam path is batch-separable, so B frames can go
    through ONE fit_frame call per sweep (its L-BFGS path cannot: the line search couples the batch).  Times the
    oracle port that way, schedule S2, all host threads; reported next to the B = 1 number, never as the target."""
    from keypoints2body_b200 import synthetic as syn
    from oracle import reference_port as rp
    from oracle.smplx_shim import BodyModelShim

    torch.set_num_threads(threads)
    weights = syn.make_body_model("smpl", seed=0)
    model, prior = BodyModelShim(weights), rp.GMMPrior(syn.make_gmm(seed=0))
    mo = syn.make_motion(frames, seed=seed)
    tgt = syn.kinematic_joints(weights, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    conf = torch.ones(22)
    root0 = model(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10)).joints[:, 0]
    init = {k: None for k in rp.PARAM_ORDER}
    init.update(global_orient=torch.zeros(frames, 3), body_pose=torch.zeros(frames, 69), betas=torch.zeros(frames, 10),
                transl=tgt[:, 0] - root0)
    t0 = time.perf_counter()
    s0 = rp.fit_frame(model, prior, init, tgt, conf, seq_ind=0, use_lbfgs=False)
    prev = {k: (v[:-1] if v is not None else None) for k, v in s0["params"].items()}
    rp.fit_frame(model, prior, prev, tgt[1:], conf, seq_ind=1, use_lbfgs=False)
    dt = time.perf_counter() - t0
    return {"value": frames / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{frames} frames in one batched call per sweep (S2, 30 + 10 iterations), adam, torch CPU, "
                      "full-mesh forward per evaluation"}


class ChainRunner:
    """Schedule S1 for the benchmark: every sequence is walked serially by one warp group (k2b_fit_chain).  The time
    axis is cut into `chunks` windows: window c+1 is fitted on a high-priority stream while the mesh pass of
    window c runs behind it (the fit is latency-bound and leaves most of every SM idle).  Outputs are time-major
    (row t * S + s).  Same interface as SequenceBatchFitter as far as run_ours uses it."""

    def __init__(self, fitter, num_frames, with_vertices=True, chunks=8):
        self.f, self.F, self.S, self.chunks = fitter, num_frames, num_frames // SEQ_LEN, chunks
        dev = fitter.device
        self.vertices = torch.empty(num_frames, fitter.native.num_vertices, 3, device=dev) if with_vertices else None
        z = {"global_orient": torch.zeros(1, 3, device=dev), "body_pose": torch.zeros(1, 69, device=dev),
             "betas": torch.zeros(1, 10, device=dev)}
        self.root0 = fitter.forward_batch(z, with_vertices=False)["joints"][:, 0, :]
        self.init = {"global_orient": torch.zeros(self.S, 3, device=dev), "body_pose": torch.zeros(self.S, 69, device=dev),
                     "betas": torch.zeros(self.S, 10, device=dev)}
        self.kernel_events = None
        self.evals0 = self.evals1 = torch.zeros(1, device=dev)

    def run(self, targets, seq_ind=None, params_ready=None, window_done=None, window_ready=None):
        f = self.f
        tg = targets.view(self.S, SEQ_LEN, 22, 3)
        init = dict(self.init, transl=tg[:, 0, 0] - self.root0)     # root alignment of frame 0 (engine.py:89-128)
        f.chain_events = self.kernel_events                          # (start, end) of the fit on its own stream
        out = f.fit_chain(init, tg, None, with_mesh=True, out_vertices=self.vertices, time_major=True,
                          chunks=self.chunks, params_ready=params_ready, fit_joints=False, window_done=window_done,
                          window_ready=window_ready)
        f.chain_events = None
        p = out["params"]
        out["pose"] = torch.cat([p["global_orient"], p["body_pose"]], dim=1)
        self.evals0 = out["evals"]
        return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch.distributed as dist

    from keypoints2body_b200 import _native as nat
    from keypoints2body_b200 import synthetic as syn
    from keypoints2body_b200.api.batch import SequenceBatchFitter
    from keypoints2body_b200.core.config import FrameOptimizeConfig
    from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    F = args.frames_per_gpu
    lo, hi = rank * F, (rank + 1) * F

    weights = syn.make_body_model("smpl", seed=0)
    fitter = WorldSpaceFitter(weights, joints_category="AMASS", use_lbfgs=args.optimizer == "lbfgs",
                              model_type="smpl", gmm=syn.make_gmm(seed=0), device=dev)
    cfg = FrameOptimizeConfig()
    chain = args.schedule == "reference"
    if chain and F % SEQ_LEN:
        raise SystemExit("--schedule reference needs whole sequences per GPU")
    with_verts = not args.no_vertices
    if chain:
        sf = ChainRunner(fitter, F, with_verts, args.chunks)
        # the frame-parallel schedule (S2) is measured beside it, into the same vertex buffer
        sf2 = None if args.no_frame_parallel else SequenceBatchFitter(fitter, F, cfg, with_vertices=False)
        if sf2 is not None:
            sf2.vertices = sf.vertices
    else:
        sf, sf2 = SequenceBatchFitter(fitter, F, cfg, with_vertices=with_verts), None
    targets = make_targets(weights, lo, hi, dev)
    seq_ind = seq_index(lo, hi, dev)       # S2: frames shard across GPUs, boundaries fall mid-sequence; S1: whole sequences
    lib = fitter.native.lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for a, b in ev:
            a.record()
            fn()
            b.record()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / steps

    # ---- value: inputs resident in HBM -------------------------------------------------------
    def step_device():
        return sf.run(targets, seq_ind)

    if chain:
        # let fit_chain's event-timed mesh-overlap policy finish its trial calls (one cold call + one per candidate share)
        # before the warm-up steps, so that the timed steps all run with the share it settled on
        for _ in range(12):
            pol = fitter.overlap_policy()
            if pol and all(v["fraction"] is not None for v in pol.values()):
                break
            step_device()
    for _ in range(args.warmup):
        step_device()
    sf.kernel_events = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.k2b_launch_count()
    ms_step = timed(step_device, args.steps)
    launches = (lib.k2b_launch_count() - launches0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    fit_ms = [a.elapsed_time(b) for a, b in sf.kernel_events]          # S2: two fit launches per step; S1: one
    sf.kernel_events = None
    out = step_device()
    evals_total = float(out["evals"].sum())
    evals0, evals1 = float(sf.evals0.sum()), float(sf.evals1.sum())
    if chain:       # outputs are time-major (row t * S + s), targets sequence-major
        jt = out["joints"][:, :22].view(SEQ_LEN, F // SEQ_LEN, 22, 3).transpose(0, 1)
        mean_err = float((jt - targets.view(F // SEQ_LEN, SEQ_LEN, 22, 3)).norm(dim=-1).mean())
    else:
        mean_err = float((out["joints"][:, :22] - targets).norm(dim=-1).mean())

    # ---- e2e: pinned host inputs -> device, fit, results -> pinned host ----------------------------
    h_targets = torch.empty(targets.shape, pin_memory=True).copy_(targets)
    d_targets = torch.empty_like(targets)
    n_j = out["joints"].shape[1]
    h_pose = torch.empty(F, 72, pin_memory=True)
    h_betas = torch.empty(F, 10, pin_memory=True)
    h_transl = torch.empty(F, 3, pin_memory=True)
    h_loss = torch.empty(F, pin_memory=True)
    h_joints = torch.empty(F, n_j, 3, pin_memory=True)

    side = torch.cuda.Stream(device=dev)
    fitted = torch.cuda.Event()

    def window_home(c, rows, fit_done, outs, joints):
        """A time window's results go home while the later windows are fitted (schedule S1: outputs are time-major, so a
        window is one contiguous run of rows in every output array)."""
        meshed = torch.cuda.Event()
        meshed.record()                          # this window's mesh pass has been enqueued on the current stream
        with torch.cuda.stream(side):
            side.wait_event(fit_done)
            h_pose[rows].copy_(outs["pose"][rows], non_blocking=True)
            h_betas[rows].copy_(outs["betas"][rows], non_blocking=True)
            h_transl[rows].copy_(outs["transl"][rows], non_blocking=True)
            h_loss[rows].copy_(outs["loss"][rows], non_blocking=True)
            side.wait_event(meshed)
            h_joints[rows].copy_(joints[rows], non_blocking=True)

    # keypoints of time window c go up while window c-1 is fitted: the device layout is [sequence][frame], so a window is
    # a 2-D region (rows = sequences) -- one cudaMemcpy2DAsync per window on a copy stream
    up, memcpy2d = torch.cuda.Stream(device=dev), None
    if chain:
        try:
            import ctypes

            rt = ctypes.CDLL("libcudart.so.12")
            memcpy2d = rt.cudaMemcpy2DAsync
            memcpy2d.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                 ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
            memcpy2d.restype = ctypes.c_int
        except Exception:
            memcpy2d = None

    def upload_windows():
        S_, row = F // SEQ_LEN, 22 * 3 * 4
        bounds = [(SEQ_LEN * c) // args.chunks for c in range(args.chunks + 1)]
        events = []
        up.wait_stream(torch.cuda.current_stream())
        for c in range(args.chunks):
            a0, b0 = bounds[c], bounds[c + 1]
            rc = memcpy2d(d_targets.data_ptr() + a0 * row, SEQ_LEN * row, h_targets.data_ptr() + a0 * row, SEQ_LEN * row,
                          (b0 - a0) * row, S_, 1, up.cuda_stream)        # 1 = cudaMemcpyHostToDevice
            if rc != 0:
                raise RuntimeError(f"cudaMemcpy2DAsync failed: {rc}")
            ev = torch.cuda.Event()
            ev.record(up)
            events.append(ev)
        return events

    def step_host():
        if chain:
            if memcpy2d is not None:
                ready = upload_windows()
            else:
                d_targets.copy_(h_targets, non_blocking=True)
                ready = None
            if ready is not None:       # frame 0's root alignment reads the first window's keypoints on the current stream
                torch.cuda.current_stream().wait_event(ready[0])
            o = sf.run(d_targets, seq_ind, window_done=window_home, window_ready=ready)
            torch.cuda.current_stream().wait_stream(side)
            return o
        d_targets.copy_(h_targets, non_blocking=True)
        o = sf.run(d_targets, seq_ind, params_ready=fitted)
        with torch.cuda.stream(side):          # parameters and losses go home while the mesh pass runs
            side.wait_event(fitted)
            h_pose.copy_(o["pose"], non_blocking=True)
            h_betas.copy_(o["params"]["betas"], non_blocking=True)
            h_transl.copy_(o["params"]["transl"], non_blocking=True)
            h_loss.copy_(o["loss"], non_blocking=True)
        h_joints.copy_(o["joints"], non_blocking=True)
        torch.cuda.current_stream().wait_stream(side)
        return o

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    ms_e2e = timed(step_host, args.steps)
    h2d = h_targets.numel() * 4
    d2h = 4 * (h_pose.numel() + h_betas.numel() + h_transl.numel() + h_loss.numel() + h_joints.numel())
    # what arrived in host memory is what the device holds (one more step, untimed)
    h_joints.fill_(float("nan")); h_pose.fill_(float("nan")); h_loss.fill_(float("nan"))
    o_chk = step_host()
    torch.cuda.synchronize()
    e2e_check = {"joints_max_abs_diff": float((h_joints.to(dev) - o_chk["joints"]).abs().max()),
                 "pose_max_abs_diff": float((h_pose.to(dev) - o_chk["pose"]).abs().max()),
                 "loss_max_abs_diff": float((h_loss.to(dev) - o_chk["loss"]).abs().max()),
                 "targets_max_abs_diff": float((d_targets - targets).abs().max())}
    del o_chk

    # ---- frame-parallel schedule S2 beside it: frames sharded across the ranks, one-frame NCCL halo ------------
    def measure_fp(sfx):
        sfx.run(targets, seq_ind)
        sfx.kernel_events, sfx.halo_events = [], []
        ms_fp = timed(lambda: sfx.run(targets, seq_ind), args.fp_steps)
        torch.cuda.synchronize()
        fp_fit_ms = sum(a.elapsed_time(b) for a, b in sfx.kernel_events) / args.fp_steps
        halo_ms = sum(a.elapsed_time(b) for a, b in sfx.halo_events) / max(1, args.fp_steps)
        sfx.kernel_events = sfx.halo_events = None
        o2 = sfx.run(targets, seq_ind)
        res = {"ms_per_step": ms_fp, "fit_ms": fp_fit_ms, "halo_ms": halo_ms,
               "evals": float(sfx.evals0.sum()) + float(sfx.evals1.sum()),
               "mean_err": float((o2["joints"][:, :22] - targets).norm(dim=-1).mean())}
        del o2
        return res

    fp = measure_fp(sf2) if sf2 is not None else None
    # ... and with Adam (the strict-parity optimiser), so that the default line carries K1's Adam roofline too
    fp_adam = None
    if sf2 is not None and args.optimizer == "lbfgs":
        fitter_a = WorldSpaceFitter(weights, joints_category="AMASS", use_lbfgs=False, model_type="smpl",
                                    gmm=syn.make_gmm(seed=0), device=dev)
        sf2a = SequenceBatchFitter(fitter_a, F, cfg, with_vertices=False)
        sf2a.vertices = sf.vertices
        fp_adam = measure_fp(sf2a)
        del sf2a, fitter_a

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ------------------------------------------------------------------------------
    peak_tflops, _ = fma_peak(lib)
    flop_step = (evals0 + evals1 + F) * EVAL_FLOP["smpl"]     # + one final (loss / joints) forward per frame/sweep
    if chain and args.optimizer == "lbfgs":
        flop_step = evals0 * EVAL_FLOP["smpl"]                  # no extra forward: the returned loss is the accepted trial's
    fit_ms_step = sum(fit_ms) / max(1, args.steps)
    achieved = flop_step / (fit_ms_step * 1e-3) / 1e12 if fit_ms_step > 0 else None
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    mesh_ms = ms_step - fit_ms_step
    if chain:     # the mesh pass overlaps the fit inside the step; for its roofline it is timed alone
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fitter.forward_batch(out["params"], out_vertices=sf.vertices)      # sizes the workspace for one full pass
        torch.cuda.synchronize()
        e0.record()
        fitter.forward_batch(out["params"], out_vertices=sf.vertices)
        e1.record()
        torch.cuda.synchronize()
        mesh_ms = e0.elapsed_time(e1)
    mesh_bytes = F * (6890 * 3 * 4 + n_j * 3 * 4) if with_verts else F * n_j * 12
    kernel_name = ("chain_kernel<10,22,%s> (warp-per-sequence; %d window launches per step)" % (args.optimizer, args.chunks)
                   if chain else "fit_kernel<10,22,%s> (sweep 0 + sweep 1)" % args.optimizer)
    traffic, traffic_note = ncu_traffic("chain" if chain else "fit", args, F)
    line = {
        "metric": METRIC, "value": world * F / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (fit); fp16 x fp16 -> fp32 tcgen05 blend in the mesh pass", "data": "synthetic",
        "config": workload_config(args, with_verts),
        "fit_quality": {"evals_per_frame": evals_total / F, "mean_joint_error_m": mean_err},
        "mesh_overlap": mesh_overlap_report(fitter) if chain else None,
        "e2e": {"value": world * F / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                "host_copy_check": e2e_check,
                "note": "vertices stay in HBM (%.1f GB/GPU); params, loss and 45 joints are copied back; see "
                        "e2e_with_vertices.%s" % (mesh_bytes / 1e9, " Keypoints go up and results come home one time window "
                        "at a time (cudaMemcpy2DAsync / pinned copies on side streams) while the other windows are fitted."
                        if chain else "")},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "roofline": {
            "kernel": kernel_name, "bound": "fp32_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops if achieved and peak_tflops else None,
            "peak_source": "in-run FFMA micro-benchmark (k2b_fma_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f" % nominal,
            "frac_of_nominal": achieved / nominal if achieved else None,
            "flop_per_eval": EVAL_FLOP["smpl"], "evals_per_step": flop_step / EVAL_FLOP["smpl"],
            "ms_per_step_in_kernel": fit_ms_step, "share_of_step": fit_ms_step / ms_step,
            "traffic": traffic, "traffic_note": traffic_note, "launches_per_step": args.chunks if chain else 2,
            "note": ("latency-bound by construction: %d serial chains per GPU; the throughput kernel is under frame_parallel"
                     % (F // SEQ_LEN)) if chain else None,
        },
        "roofline_mesh": {
            "kernel": "mesh_pose_kernel + blend_skin_tc_kernel (tcgen05 blend, LBS in the epilogue) + gather_extra_kernel", "bound": "hbm",
            "achieved": mesh_bytes / (mesh_ms * 1e-3) / 1e9 if mesh_ms > 0 else None,
            "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
            "frac": (mesh_bytes / (mesh_ms * 1e-3) / 1e9) / peaks.get("hbm_gbs", 6650.0) if mesh_ms > 0 else None,
            "ms": mesh_ms, "traffic": None,
            "note": ("timed alone after the run; inside the step it overlaps the fit (exposed: %.1f ms)" % (ms_step - fit_ms_step))
                    if chain else "step time minus the fit launches",
        },
    }
    if fp is not None:
        fp_flop = (fp["evals"] + 2 * F) * EVAL_FLOP["smpl"]
        fp_ach = fp_flop / (fp["fit_ms"] * 1e-3) / 1e12
        t2, n2 = ncu_traffic("fit", args, F)
        line["frame_parallel"] = {
            "what": "schedule S2 on the same keypoints: frames sharded across the ranks (each shard boundary falls inside a "
                    "sequence), sweep 0 + one-frame halo over NCCL + sweep 1, full mesh; one thread per frame (fit_kernel)",
            "value": world * F / (fp["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": fp["ms_per_step"],
            "steps": args.fp_steps, "evals_per_frame": fp["evals"] / F, "mean_joint_error_m": fp["mean_err"],
            "halo_ms_per_step": fp["halo_ms"], "halo_bytes_per_boundary": 95 * 4, "n_boundaries": world - 1,
            "roofline": {"kernel": "fit_kernel<10,22,%s> (sweep 0 + sweep 1)" % args.optimizer, "bound": "fp32_fma",
                         "achieved": fp_ach, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": fp_ach / peak_tflops if peak_tflops else None, "flop_per_eval": EVAL_FLOP["smpl"],
                         "ms_per_step_in_kernel": fp["fit_ms"], "share_of_step": fp["fit_ms"] / fp["ms_per_step"],
                         "traffic": t2, "traffic_note": n2},
        }
    if fp_adam is not None:
        fa_ach = (fp_adam["evals"] + 2 * F) * EVAL_FLOP["smpl"] / (fp_adam["fit_ms"] * 1e-3) / 1e12
        line["frame_parallel_adam"] = {
            "what": "the same frame-parallel schedule with Adam (the strict-parity optimiser): fit_kernel<10,22,adam>",
            "value": world * F / (fp_adam["ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": fp_adam["ms_per_step"],
            "steps": args.fp_steps, "evals_per_frame": fp_adam["evals"] / F, "mean_joint_error_m": fp_adam["mean_err"],
            "halo_ms_per_step": fp_adam["halo_ms"],
            "roofline": {"kernel": "fit_kernel<10,22,adam> (sweep 0 + sweep 1)", "bound": "fp32_fma", "achieved": fa_ach,
                         "peak": peak_tflops, "unit": "TFLOP/s", "frac": fa_ach / peak_tflops if peak_tflops else None,
                         "flop_per_eval": EVAL_FLOP["smpl"], "ms_per_step_in_kernel": fp_adam["fit_ms"]},
        }
    # ---- the vertices were written: a sample re-evaluated by the FP32 CUDA-core mesh path ---------------------------
    if with_verts:
        g = torch.Generator().manual_seed(11)
        idx = torch.randint(0, F, (256,), generator=g).to(dev)
        sub = {k: v[idx].contiguous() for k, v in out["params"].items()}
        os.environ["K2B_MESH_FP32"] = "1"
        chk = fitter.forward_batch(sub)["vertices"]
        os.environ["K2B_MESH_FP32"] = "0"
        got = sf.vertices[idx]
        line["vertex_check"] = {"frames": 256, "max_abs_diff_m": float((chk - got).abs().max()),
                                "checksum": float(got.double().sum()),
                                "what": "256 random frames of the step's vertex output against the FP32 CUDA-core mesh path"}
        del chk, got
    # ---- e2e with the vertices streamed to pinned host memory in windows (PCIe-bound) ------------------------------
    if with_verts and world == 1 and not args.no_e2e_vertices:
        win = 16384
        pin = [torch.empty(win, 6890, 3, pin_memory=True) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)

        def step_host_vertices():
            step_host()
            copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                for i, a0 in enumerate(range(0, F, win)):
                    b0 = min(F, a0 + win)
                    pin[i % 2][: b0 - a0].copy_(sf.vertices[a0:b0], non_blocking=True)
            torch.cuda.current_stream().wait_stream(copy_stream)

        ms_v = timed(step_host_vertices, 1)
        line["e2e_with_vertices"] = {"value": F / (ms_v * 1e-3), "unit": UNIT, "ms_per_step": ms_v, "steps": 1,
                                     "d2h_bytes_per_step": d2h + F * 6890 * 12,
                                     "note": "every vertex leaves the GPU too (%.1f GB through two 1.35 GB pinned windows): "
                                             "PCIe-bound, %.1f GB/s" % (F * 6890 * 12 / 1e9, F * 6890 * 12 / (ms_v * 1e-3) / 1e9)}
        del pin
    # ---- CPU baseline on the same frames, and our result on them against it ------------------------------------------
    if world == 1 and not args.skip_cpu_baseline:
        threads = os.cpu_count() or 1
        arm = CpuArm(args.optimizer)
        sample = cpu_sample_targets(args)
        arm.fit_chain(sample[:8], threads)          # warm-up (thread pool, allocator)
        ref_run = arm.fit_chain(sample, threads)
        line["cpu_baseline"] = cpu_baseline_block(arm, sample, threads, ref_run, with_demo=False)
        n = len(sample)
        z = {"global_orient": torch.zeros(1, 3), "body_pose": torch.zeros(1, 69), "betas": torch.zeros(1, 10)}
        init = dict(z, transl=(sample[0:1, 0].to(dev) - sf.root0).cpu() if chain else sample[0:1, 0])
        if not chain:
            init["transl"] = sample[0:1, 0] - fitter.forward_batch(z, with_vertices=False)["joints"][:, 0].cpu()
        ours = fitter.fit_chain(init, sample[None].to(dev), None, with_mesh=False)
        pose = torch.cat([ours["params"]["global_orient"], ours["params"]["body_pose"]], dim=1).cpu()
        o_err = (ours["fit_joints"].cpu() - sample).norm(dim=-1).mean(dim=-1)
        line["parity_check"] = {
            "what": "our chain fit of the CPU sample's frames against the CPU arm's result on them (L-BFGS: distribution; "
                    "the reference differs from itself by a median 6.5 % per frame across thread counts)",
            "frames": n, "loss_median_ratio": float(ours["loss"].cpu().median() / ref_run["loss"].median()),
            "joint_error_median_ratio": float(o_err.median() / ref_run["err"].median()),
            "evals_per_frame": [float(ours["evals"].float().mean()), ref_run["evals_per_frame"]],
            "first_frames_pose_max_abs_diff": float((pose[:2] - ref_run["pose"][:2]).abs().max()),
        }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(which: str, args, frames: int):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    `ncu --set full` capture of this workload (profiles/r02_*_ncu_metrics.txt); null when the configuration differs."""
    name = {"chain": "r02_chain_ncu_metrics.txt", "fit": "r02_fit_ncu_metrics.txt"}[which]
    if args.optimizer != "lbfgs" or frames != 256 * SEQ_LEN:
        return None, "no committed capture for this configuration"
    try:
        total, seen = 0.0, 0
        for ln in open(os.path.join(ROOT, "profiles", name)):
            if ln.startswith("## launch 1"):
                break
            if ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum"):
                val, unit = ln.split("=")[1].split()[:2]
                total += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                seen += 1
        if seen == 2:
            return total, f"first launch in profiles/{name} (ncu --set full capture of this command)"
    except Exception:
        pass
    return None, "no committed capture"


def fma_peak(lib):
    import ctypes

    tf, ms = ctypes.c_double(), ctypes.c_double()
    rc = lib.k2b_fma_peak(20000, ctypes.byref(tf), ctypes.byref(ms), None)
    return (tf.value, ms.value) if rc == 0 else (None, None)


if __name__ == "__main__":
    a = parse()
    claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
