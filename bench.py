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
    ap.add_argument("--cpu-sample-frames", type=int, default=0, help="0 = choose for ~20 s of CPU work")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
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
def cpu_reference_rate(frames: int, threads: int, optimizer: str, seed: int = 77, schedule: str = "two_sweep"):
    """Times the oracle port (the reference's algorithm on torch CPU, B = 1 per frame like the API,
    full-mesh forward per evaluation like smplx) on `frames` frames with schedule S2 or S1."""
    from keypoints2body_b200 import synthetic as syn
    from oracle import reference_port as rp
    from oracle.smplx_shim import BodyModelShim

    torch.set_num_threads(threads)
    weights = syn.make_body_model("smpl", seed=0)
    model, prior = BodyModelShim(weights), rp.GMMPrior(syn.make_gmm(seed=0))
    mo = syn.make_motion(frames, seed=seed)
    tgt = syn.kinematic_joints(weights, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    lbfgs = optimizer == "lbfgs"
    conf = torch.ones(22)
    root0 = model(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10)).joints[:, 0]

    def chain_pass():      # S1: api/sequence.py:214-281
        prev = {k: None for k in rp.PARAM_ORDER}
        prev.update(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10),
                    transl=tgt[0:1, 0] - root0)
        for t in range(frames):
            prev = rp.fit_frame(model, prior, prev, tgt[t:t + 1], conf, seq_ind=t, use_lbfgs=lbfgs)["params"]

    if schedule == "reference":
        return chain_pass

    def one_pass():
        s0 = []
        for t in range(frames):
            init = {k: None for k in rp.PARAM_ORDER}
            init.update(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10),
                        transl=tgt[t:t + 1, 0] - root0)
            s0.append(rp.fit_frame(model, prior, init, tgt[t:t + 1], conf, seq_ind=0, use_lbfgs=lbfgs))
        for t in range(1, frames):
            rp.fit_frame(model, prior, s0[t - 1]["params"], tgt[t:t + 1], conf, seq_ind=t, use_lbfgs=lbfgs)

    return one_pass


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    probe = cpu_reference_rate(2, threads, args.optimizer, schedule=args.schedule)
    t0 = time.perf_counter()
    probe()
    per_frame = (time.perf_counter() - t0) / 2
    budget = 150.0 / max(1, args.steps + args.warmup)
    n = args.cpu_sample_frames or int(max(2, min(64, budget / per_frame)))
    one_pass = cpu_reference_rate(n, threads, args.optimizer, schedule=args.schedule)
    for _ in range(args.warmup):
        one_pass()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_pass()
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt
    sched = "S2 (30 + 10 iteration budgets)" if args.schedule == "two_sweep" else "S1 (serial chain, 30 then 10 iterations)"
    sample = (f"{n} frames/step, schedule {sched}, {args.optimizer}, B=1 per frame, "
              f"torch {torch.__version__} CPU, full-mesh forward per evaluation")
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "bounded sample of the ours-arm workload: " + sample, "optimizer": args.optimizer,
                   "schedule": args.schedule},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "cpu_baseline_batched_adam": cpu_reference_batched_adam(1024, threads),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


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

    def run(self, targets, seq_ind=None, params_ready=None):
        f = self.f
        tg = targets.view(self.S, SEQ_LEN, 22, 3)
        init = dict(self.init, transl=tg[:, 0, 0] - self.root0)     # root alignment of frame 0 (engine.py:89-128)
        f.chain_events = self.kernel_events                          # (start, end) of the fit on its own stream
        out = f.fit_chain(init, tg, None, with_mesh=True, out_vertices=self.vertices, time_major=True,
                          chunks=self.chunks, params_ready=params_ready, fit_joints=False)
        f.chain_events = None
        p = out["params"]
        out["pose"] = torch.cat([p["global_orient"], p["body_pose"]], dim=1)
        self.evals0 = out["evals"]
        return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
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
    sf = ChainRunner(fitter, F, not args.no_vertices, args.chunks) if chain else SequenceBatchFitter(
        fitter, F, cfg, with_vertices=not args.no_vertices)
    targets = make_targets(weights, lo, hi, dev)
    seq_ind = seq_index(lo, hi, dev)       # S2 only; S1 takes [lo, hi) as F / 4096 whole sequences
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

    def step_host():
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

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    ms_e2e = timed(step_host, args.steps)
    h2d = h_targets.numel() * 4
    d2h = 4 * (h_pose.numel() + h_betas.numel() + h_transl.numel() + h_loss.numel() + h_joints.numel())

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (fused fit kernel, two launches per step) -------------------
    tf = ctypes_double()
    peak_tflops, _ = fma_peak(lib)
    flop_step = (evals0 + evals1 + F) * EVAL_FLOP["smpl"]     # + one final (loss / joints) forward per frame/sweep
    fit_ms_step = sum(fit_ms) / max(1, args.steps)
    achieved = flop_step / (fit_ms_step * 1e-3) / 1e12 if fit_ms_step > 0 else None
    nominal = 148 * 128 * 2 * 1.965e9 / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    mesh_ms = ms_step - fit_ms_step
    if chain:     # the mesh pass overlaps the fit inside the step; for its roofline it is timed alone (rank 0
        # only: no collective here, the other ranks have already returned)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fitter.forward_batch(out["params"], out_vertices=sf.vertices)      # sizes the workspace for one full pass
        torch.cuda.synchronize()
        e0.record()
        fitter.forward_batch(out["params"], out_vertices=sf.vertices)
        e1.record()
        torch.cuda.synchronize()
        mesh_ms = e0.elapsed_time(e1)
    mesh_bytes = F * (6890 * 3 * 4 + n_j * 3 * 4) if not args.no_vertices else F * n_j * 12
    line = {
        "metric": METRIC, "value": world * F / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": (f"SMPL AMASS-22 sequence fit, {F} frames/GPU ({F // SEQ_LEN} sequences x {SEQ_LEN}; "
                         "BASELINE configs[3] shard, each sequence = configs[1]), "
                         + ("schedule S1 = the reference's own serial chain (frame t starts from frame t-1's result; 30 "
                            "iterations for frame 0, 10 + pose-preserve after), one warp group per sequence, whole "
                            f"sequences per GPU, {args.chunks} time windows (mesh of window c overlaps the fit of c+1), " if chain else
                            "schedule S2: sweep0 30-iteration budget + sweep1 10-iteration budget with pose-preserve, ")
                         + f"{args.optimizer}, "
                         + ("full mesh (6890 verts + 45 joints) per frame" if not args.no_vertices else "joints only")),
            "optimizer": args.optimizer, "schedule": args.schedule, "frames_per_gpu": F, "seq_len": SEQ_LEN,
            "l2_policy": "inputs larger than L2 (targets %.0f MB/GPU, outputs %.1f GB/GPU)" % (h2d / 1e6, mesh_bytes / 1e9),
            "evals_per_frame": evals_total / F, "mean_joint_error_m": mean_err,
        },
        "e2e": {"value": world * F / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e,
                "note": "vertices stay in HBM (%.1f GB/GPU); params, loss and 45 joints are copied back" % (mesh_bytes / 1e9)},
        "gpu_launches": int(launches) * args.steps,
        "clocks": clocks,
        "roofline": {
            "kernel": ("chain_kernel<10,22> (%s; one launch, latency-bound: F/4096 warps)" if chain else
                       "fit_kernel<10,22,%s> (sweep 0 + sweep 1)") % ("lbfgs" if args.optimizer == "lbfgs" else "adam"),
            "bound": "fp32_fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
            "frac": achieved / peak_tflops if achieved and peak_tflops else None,
            "peak_source": "in-run FFMA micro-benchmark (k2b_fma_peak); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = %.1f" % nominal,
            "frac_of_nominal": achieved / nominal if achieved else None,
            "flop_per_eval": EVAL_FLOP["smpl"], "evals_per_launch": [evals0 + F] if chain else [evals0 + F, evals1 + F],
            "ms_per_launch_pair": fit_ms_step, "share_of_step": fit_ms_step / ms_step,
            "traffic": ncu_chain_traffic(args.optimizer, F, args.chunks) if chain else ncu_fit_traffic(args.optimizer, F),
            "traffic_note": ("dram__bytes_read.sum + dram__bytes_write.sum of ONE of the %d window launches of a step, from "
                             "the committed ncu --set full capture of this command (profiles/r01_chain_ncu_metrics.txt): "
                             "the window's keypoints are read once (17.3 MB), everything else stays on chip or in L2"
                             % args.chunks) if chain else
                            "dram__bytes_read.sum + dram__bytes_write.sum of the sweep-0 launch from the committed ncu "
                            "--set full capture (profiles/r01_fit_lbfgs_ncu_metrics.txt, same frame count); the "
                            "algorithmic HBM bytes are ~1.5 KB/frame, the rest is L-BFGS (s, y) history that does not fit L2",
            "launches_per_step": args.chunks if chain else 2,
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
    del tf
    if world == 1 and not args.skip_cpu_baseline:
        threads = os.cpu_count() or 1
        probe = cpu_reference_rate(2, threads, args.optimizer, schedule=args.schedule)
        t0 = time.perf_counter()
        probe()
        per_frame = (time.perf_counter() - t0) / 2
        n = args.cpu_sample_frames or int(max(2, min(64, 20.0 / per_frame)))
        one_pass = cpu_reference_rate(n, threads, args.optimizer, schedule=args.schedule)
        t0 = time.perf_counter()
        one_pass()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} frames, schedule {'S1 (serial chain)' if chain else 'S2 (30 + 10 budgets)'}, {args.optimizer}, B=1 per frame, torch CPU, "
                      "full-mesh forward per evaluation like the reference"}
        line["cpu_baseline_batched_adam"] = cpu_reference_batched_adam(1024, threads)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def ncu_chain_traffic(optimizer: str, frames: int, chunks: int):
    """DRAM bytes of one window launch of the chain kernel from the committed ncu capture of the default bench; null
    when the configuration differs from the captured one."""
    if optimizer != "lbfgs" or frames != 256 * SEQ_LEN or chunks != 16:
        return None
    try:
        total, seen, inside = 0.0, 0, False
        for ln in open(os.path.join(ROOT, "profiles", "r01_chain_ncu_metrics.txt")):
            if ln.startswith("## bench launch 0"):
                inside = True
            elif ln.startswith("## bench launch 1"):
                break
            elif inside and (ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum")):
                val, unit = ln.split("=")[1].split()
                total += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                seen += 1
        return total if seen == 2 else None
    except Exception:
        return None


def ncu_fit_traffic(optimizer: str, frames: int):
    """DRAM bytes of the dominant launch (sweep 0 of the L-BFGS fit kernel) from the committed ncu capture; null
    when there is no capture for this optimiser / frame count."""
    if optimizer != "lbfgs" or frames != 256 * SEQ_LEN:
        return None
    try:
        total, seen = 0.0, 0
        for ln in open(os.path.join(ROOT, "profiles", "r01_fit_lbfgs_ncu_metrics.txt")):
            if ln.startswith("## launch 1"):
                break
            if ln.startswith("dram__bytes_read.sum") or ln.startswith("dram__bytes_write.sum"):
                val, unit = ln.split("=")[1].split()
                total += float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
                seen += 1
        return total if seen == 2 else None
    except Exception:
        return None


def ctypes_double():
    import ctypes

    return ctypes.c_double()


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
