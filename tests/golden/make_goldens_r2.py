"""Round-2 goldens from the UNMODIFIED reference (authoring container only; /root/reference cannot travel).

    python tests/golden/make_goldens_r2.py points   -> r2_points.npz   every L-BFGS trial point (x, f, g) of the reference
    python tests/golden/make_goldens_r2.py dist     -> r2_dist.npz     512 independent L-BFGS fits, both budgets
    python tests/golden/make_goldens_r2.py dist64   -> r2_dist64.npz   the same fits by the oracle port in float64
    python tests/golden/make_goldens_r2.py chains   -> r2_chains.npz   32 chains x 64 frames, L-BFGS, reference sequence driver
    python tests/golden/make_goldens_r2.py adam     -> r2_adam.npz     Adam chains: the two demo sequences (195 / 116 real
                                                                       AMASS-22 frames) and a 512-frame synthetic chain
    python tests/golden/make_goldens_r2.py adam_pert -> r2_adam_pert.npz  the same chains from an initial pose 1e-7 rad away

What runs is the reference's own WorldSpaceFitter / optimize_params_sequence / torch.optim (imported through
oracle.ref_loader's stubs; body model = oracle.smplx_shim on the seeded synthetic weights).  Inputs are regenerated
from seeds by oracle/problems.py; only the demo keypoints (reference data files, 80 KB) are stored with the outputs.
Single-threaded torch, so the L-BFGS runs are reproducible.
"""

from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from keypoints2body_b200 import synthetic as syn  # noqa: E402
from oracle import problems, ref_loader  # noqa: E402
from oracle import reference_port as rp  # noqa: E402
from oracle.smplx_shim import BodyModelShim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DIST_N, DIST_SEED = 512, 2024
CHAIN_S, CHAIN_T, CHAIN_SEED = 32, 64, 3030
LONG_T, LONG_SEED = 512, 4040


def save(name, arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", path, round(os.path.getsize(path) / 1e3, 1), "KB,", len(arrays), "arrays")


class LbfgsSpy:
    """Records, for every torch.optim.LBFGS.step, the number of closure evaluations and (optionally) the flat
    parameter vector, loss and flat gradient of every closure call."""

    def __init__(self, points=False):
        self.points, self.evals, self.trace = points, [], []

    def __enter__(self):
        import torch.optim.lbfgs as L

        self.L, self.orig = L, L.LBFGS.step
        spy = self

        def step(opt, closure):
            def wrapped():
                loss = closure()
                if spy.points:
                    x = torch.cat([p.detach().reshape(-1) for p in opt._params]).clone()
                    spy.trace.append((x, float(loss), opt._gather_flat_grad().clone()))
                return loss

            out = spy.orig(opt, wrapped)
            spy.evals.append(int(opt.state[opt._params[0]]["func_evals"]))
            return out

        L.LBFGS.step = step
        return self

    def __exit__(self, *exc):
        self.L.LBFGS.step = self.orig
        return False


def setup():
    ref = ref_loader.load_reference()
    tmp = tempfile.mkdtemp()
    syn.write_assets(os.path.join(tmp, "data/models"), seed=0)
    return ref, tmp


def smpl_init(SMPLData, init, b):
    return SMPLData(betas=init["betas"][b:b + 1], global_orient=init["global_orient"][b:b + 1],
                    body_pose=init["body_pose"][b:b + 1], transl=init["transl"][b:b + 1])


# ---------------------------------------------------------------------------------------------------------
def section_points():
    """Teacher-forced evaluation parity: the reference's L-BFGS runs of make_goldens.py (same seeds) with the
    flat x of every closure call, so each trial point can be fed to k2b_evaluate_batch."""
    ref, tmp = setup()
    from keypoints2body.core.fitters.world_space import WorldSpaceFitter
    from keypoints2body.models.smpl_data import SMPLData, SMPLXData

    from make_goldens import targets_for     # same inputs as the round-1 L-BFGS goldens

    G = {}
    with ref_loader.reference_cwd(tmp):
        for tag, mt, n_first, n_follow, seq_ind, B, seed in (("first", "smpl", 30, 10, 0, 4, 21),
                                                             ("follow", "smpl", 30, 10, 2, 4, 22),
                                                             ("smplx", "smplx", 5, 5, 0, 3, 23)):
            model = BodyModelShim(syn.make_body_model(mt, seed=0))
            mo, tgt = targets_for(model, B, seed=seed)
            g = torch.Generator().manual_seed(seed + 5)
            pose = mo["pose"] + 0.1 * torch.randn(B, 72, generator=g)
            betas = torch.zeros(B, 10)
            transl = mo["transl"] + 0.03 * torch.randn(B, 3, generator=g)
            fitter = WorldSpaceFitter(model, step_size=1e-2, num_iters_first=n_first, num_iters_followup=n_follow,
                                      use_lbfgs=True, joints_category="AMASS")
            xs, fs, gs, owner = [], [], [], []
            for b in range(B):
                base = dict(betas=betas[b:b + 1], global_orient=pose[b:b + 1, :3], body_pose=pose[b:b + 1, 3:],
                            transl=transl[b:b + 1])
                if mt == "smplx":
                    init = SMPLXData(**base, left_hand_pose=torch.zeros(1, 45), right_hand_pose=torch.zeros(1, 45),
                                     expression=torch.zeros(1, 10), jaw_pose=torch.zeros(1, 3),
                                     leye_pose=torch.zeros(1, 3), reye_pose=torch.zeros(1, 3))
                else:
                    init = SMPLData(**base)
                with LbfgsSpy(points=True) as spy:
                    fitter.fit_frame(init, tgt[b:b + 1], torch.ones(22), seq_ind=seq_ind)
                for x, f, gr in spy.trace:
                    xs.append(x); fs.append(f); gs.append(gr); owner.append(b)
            G[f"{tag}_x"] = torch.stack(xs)
            G[f"{tag}_f"] = np.asarray(fs, np.float64)
            G[f"{tag}_g"] = torch.stack(gs)
            G[f"{tag}_frame"] = np.asarray(owner, np.int32)
            G[f"{tag}_target"] = tgt
            G[f"{tag}_keep"] = pose[:, 3:]
            G[f"{tag}_seq_ind"] = np.asarray(seq_ind)
            print(tag, "trial points:", len(xs))
    save("r2_points.npz", G)


def section_dist():
    """G4(ii): 512 independent frames through the reference's L-BFGS fitter, first-frame and follow-up budgets."""
    ref, tmp = setup()
    from keypoints2body.core.fitters.world_space import WorldSpaceFitter
    from keypoints2body.models.smpl_data import SMPLData

    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    tgt, init = problems.frame_problem(weights, DIST_N, DIST_SEED)
    G = {"n": np.asarray(DIST_N), "seed": np.asarray(DIST_SEED), "target_sum": tgt.double().sum()}
    with ref_loader.reference_cwd(tmp):
        fitter = WorldSpaceFitter(model, step_size=1e-2, num_iters_first=30, num_iters_followup=10, use_lbfgs=True,
                                  joints_category="AMASS")
        for seq_ind in (0, 1):
            t0 = time.time()
            loss, err, evals, pose = [], [], [], []
            with LbfgsSpy() as spy:
                for b in range(DIST_N):
                    r = fitter.fit_frame(smpl_init(SMPLData, init, b), tgt[b:b + 1], torch.ones(22), seq_ind=seq_ind)
                    loss.append(float(r.loss))
                    err.append(float(problems.mean_joint_error(r.joints[:, :22], tgt[b:b + 1])))
                    pose.append(r.params.pose[0])
                evals = list(spy.evals)
            G[f"s{seq_ind}_loss"] = np.asarray(loss)
            G[f"s{seq_ind}_err"] = np.asarray(err)
            G[f"s{seq_ind}_evals"] = np.asarray(evals, np.int32)
            G[f"s{seq_ind}_pose"] = torch.stack(pose)
            print("dist seq_ind", seq_ind, "median loss", np.median(loss), "median err", np.median(err),
                  "mean evals", np.mean(evals), f"{time.time() - t0:.0f}s")
        # the reference against itself: the same follow-up fits on 4 threads (first 128 frames)
        torch.set_num_threads(4)
        loss4, err4 = [], []
        for b in range(128):
            r = fitter.fit_frame(smpl_init(SMPLData, init, b), tgt[b:b + 1], torch.ones(22), seq_ind=1)
            loss4.append(float(r.loss))
            err4.append(float(problems.mean_joint_error(r.joints[:, :22], tgt[b:b + 1])))
        torch.set_num_threads(1)
        G["s1_loss_4threads"] = np.asarray(loss4)
        G["s1_err_4threads"] = np.asarray(err4)
    save("r2_dist.npz", G)


def section_dist64():
    """G4(iii): the same 512 fits by the oracle port in float64 (the reference hard-codes float32 for its prior,
    world_space.py:87-91, so the float64 run is the pinned port's)."""
    w64 = syn.make_body_model("smpl", seed=0, dtype=torch.float64)
    w32 = syn.make_body_model("smpl", seed=0)
    model, prior = BodyModelShim(w64), rp.GMMPrior(syn.make_gmm(seed=0), dtype=torch.float64)
    tgt, init = problems.frame_problem(w32, DIST_N, DIST_SEED)
    tgt = tgt.double()
    G = {}
    for seq_ind in (0, 1):
        loss, err = [], []
        for b in range(DIST_N):
            sub = {k: None for k in rp.PARAM_ORDER}
            sub.update({k: v[b:b + 1].double() for k, v in init.items()})
            r = rp.fit_frame(model, prior, sub, tgt[b:b + 1], torch.ones(22, dtype=torch.float64), seq_ind=seq_ind,
                             use_lbfgs=True)
            loss.append(float(r["loss"]))
            err.append(float(problems.mean_joint_error(r["joints"][:, :22], tgt[b:b + 1])))
        G[f"s{seq_ind}_loss"] = np.asarray(loss)
        G[f"s{seq_ind}_err"] = np.asarray(err)
        print("dist64 seq_ind", seq_ind, "median loss", np.median(loss), "median err", np.median(err))
    save("r2_dist64.npz", G)


def run_sequence(ref, model, joints, cfg):
    with LbfgsSpy() as spy:
        res = ref.optimize_params_sequence(joints, body_model="smpl", joint_layout="AMASS", model=model, config=cfg)
    tgt = torch.as_tensor(np.asarray(joints), dtype=torch.float32)
    out = {
        "pose": torch.cat([r.params.pose for r in res]), "betas": torch.cat([r.params.betas for r in res]),
        "transl": torch.cat([r.params.transl for r in res]), "joints22": torch.cat([r.joints[:, :22] for r in res]),
        "loss": torch.stack([r.loss.reshape(()) for r in res]),
    }
    out["err"] = problems.mean_joint_error(out["joints22"], tgt)
    out["evals"] = np.asarray(spy.evals, np.int32)
    return out


def section_chains():
    """32 chains x 64 frames through the reference's own optimize_params_sequence (L-BFGS, S1)."""
    ref, tmp = setup()
    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    tgt = problems.chain_problem(weights, CHAIN_S, CHAIN_T, CHAIN_SEED)
    loss, err, evals = [], [], []
    with ref_loader.reference_cwd(tmp):
        for s in range(CHAIN_S):
            t0 = time.time()
            o = run_sequence(ref, model, tgt[s].numpy(), dict(frame=dict(use_lbfgs=True), use_shape_optimization=False))
            loss.append(o["loss"].numpy()); err.append(o["err"].numpy()); evals.append(o["evals"])
            print("chain", s, "median err", float(np.median(err[-1])), "mean evals", evals[-1][1:].mean(),
                  f"{time.time() - t0:.0f}s", flush=True)
    save("r2_chains.npz", {"S": np.asarray(CHAIN_S), "T": np.asarray(CHAIN_T), "seed": np.asarray(CHAIN_SEED),
                           "loss": np.stack(loss), "err": np.stack(err), "evals": np.stack(evals),
                           "target_sum": tgt.double().sum()})


def section_adam():
    """Long Adam chains (strict parity at every frame): the reference's two demo sequences and a 512-frame
    synthetic chain; plus the demo sequence with the default configuration (L-BFGS + shape pre-pass)."""
    ref, tmp = setup()
    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    G = {}
    demo = {n: np.load(f"/root/reference/data/demo/test_motion{n}.npy").astype(np.float32) for n in (1, 2)}
    long_tgt = problems.chain_problem(weights, 1, LONG_T, LONG_SEED)[0].numpy()
    with ref_loader.reference_cwd(tmp):
        for name, joints in (("demo1", demo[1]), ("demo2", demo[2]), ("long512", long_tgt)):
            t0 = time.time()
            o = run_sequence(ref, model, joints, dict(frame=dict(use_lbfgs=False), use_shape_optimization=False))
            for k in ("pose", "betas", "transl", "joints22", "loss"):
                G[f"{name}_{k}"] = o[k]
            if name.startswith("demo"):
                G[f"{name}_in"] = joints
            print(name, joints.shape, "median err", float(o["err"].median()), f"{time.time() - t0:.0f}s", flush=True)
        o = run_sequence(ref, model, demo[1], None)       # defaults: L-BFGS, shape pre-pass over the first 50 frames
        G["demo1_default_loss"] = o["loss"]
        G["demo1_default_err"] = o["err"]
        G["demo1_default_evals"] = o["evals"]               # entry 0 is the shape pass
        G["demo1_default_betas0"] = o["betas"][0]
        print("demo1 default: median err", float(o["err"].median()), "evals", o["evals"][:4])
    save("r2_adam.npz", G)


def section_adam_pert():
    """The reference against itself along the long Adam chains: the same runs from an initial body pose moved by 1e-7 rad.
    Unobserved leaf joints (feet, head, hands) see only the priors, their gradients hover at rounding level, and Adam's
    g / (sqrt(v) + eps) turns a sign flip there into a full-size step: the chain amplifies 1e-7 to centimetres within ~20
    frames.  These runs are the yardstick for a free-running comparison (tests/test_gpu_lbfgs_parity.py)."""
    ref, tmp = setup()
    from keypoints2body.models.smpl_data import SMPLData

    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    G = {}
    demo = {n: np.load(f"/root/reference/data/demo/test_motion{n}.npy").astype(np.float32) for n in (1, 2)}
    long_tgt = problems.chain_problem(weights, 1, LONG_T, LONG_SEED)[0].numpy()
    with torch.no_grad():
        root = model(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10)).joints[:, 0]
    with ref_loader.reference_cwd(tmp):
        for name, joints in (("demo1", demo[1]), ("demo2", demo[2]), ("long512", long_tgt)):
            init = SMPLData(betas=torch.zeros(1, 10), global_orient=torch.zeros(1, 3), body_pose=torch.full((1, 69), 1e-7),
                            transl=torch.as_tensor(joints[0:1, 0]) - root)
            res = ref.optimize_params_sequence(joints, init_params=init, body_model="smpl", joint_layout="AMASS", model=model,
                                               config=dict(frame=dict(use_lbfgs=False), use_shape_optimization=False))
            G[f"{name}_pose"] = torch.cat([r.params.pose for r in res])
            G[f"{name}_joints22"] = torch.cat([r.joints[:, :22] for r in res])
            print(name, "done", flush=True)
    save("r2_adam_pert.npz", G)


def generic_problem(mt, B, seed):
    """Ground-truth parameters with articulated hands / face, targets = the model's joints at them + 3 mm noise, and an
    initialisation a little off.  Shared with the tests (regenerated from the seed)."""
    from oracle.problems import articulated_problem

    return articulated_problem(mt, B, seed)


def section_generic():
    """Observations of hands / face (dict-block indices: kinematic finger joints and vertex-picked landmarks) through the
    reference's WorldSpaceFitter GENERIC path (world_space.py:198-201), and its MANO / FLAME fitters
    (core/fitters/misc_models.py:18-359).  Adam: strict goldens; L-BFGS: losses and evaluation counts."""
    ref, tmp = setup()
    from keypoints2body.core.fitters.misc_models import FLAMEFitter, MANOFitter
    from keypoints2body.core.fitters.world_space import WorldSpaceFitter
    from keypoints2body.models.smpl_data import FLAMEData, MANOData, SMPLHData, SMPLXData

    from oracle.problems import articulated_problem

    G = {}
    with ref_loader.reference_cwd(tmp):
        for mt in ("smplx", "smplh", "mano", "flame"):
            model, tgt, idx, init, B = articulated_problem(mt, 3, seed=700)
            cls = {"smplx": SMPLXData, "smplh": SMPLHData, "mano": MANOData, "flame": FLAMEData}[mt]
            for opt, lbfgs in (("adam", False), ("lbfgs", True)):
                for seq_ind in (0, 2):
                    if mt in ("smplx", "smplh"):
                        fit = WorldSpaceFitter(model, step_size=1e-2, num_iters_first=10, num_iters_followup=10, use_lbfgs=lbfgs,
                                               joints_category="GENERIC")
                    else:
                        fit = (MANOFitter if mt == "mano" else FLAMEFitter)(model, coordinate_mode="world", step_size=1e-2,
                                                                            num_iters_first=10, num_iters_followup=10,
                                                                            use_lbfgs=lbfgs)
                    outs = {}
                    with LbfgsSpy() as spy:
                        for b in range(B):
                            p0 = cls(**{k: (v[b:b + 1] if v is not None else None) for k, v in init.items()})
                            r = fit.fit_frame(p0, tgt[b:b + 1], torch.ones(len(idx)), seq_ind=seq_ind, target_model_indices=idx)
                            for k in init:
                                v = getattr(r.params, k)
                                if v is not None:
                                    outs.setdefault(k, []).append(v)
                            outs.setdefault("joints", []).append(r.joints)
                            outs.setdefault("loss", []).append(r.loss.reshape(1))
                            outs.setdefault("verts0", []).append(r.vertices[:, :64])
                    tag = f"{mt}_{opt}_s{seq_ind}"
                    for k, v in outs.items():
                        G[f"{tag}_{k}"] = torch.cat(v)
                    if lbfgs:
                        G[f"{tag}_evals"] = np.asarray(spy.evals, np.int32)
                    print(tag, "loss", [round(float(x), 2) for x in outs["loss"]], flush=True)
    save("r2_generic.npz", G)


def section_generic_api():
    """The reference's PUBLIC API on the articulated inputs: optimize_params_sequence for MANO / FLAME (raw model joint
    order, api/sequence.py:155-158, 192-281) and for SMPL-X dict blocks body + both hands + face (adapters.py:224-380),
    optimize_params_frame for SMPL-H dict blocks; Adam, 4-frame chains from the default initialisation."""
    ref, tmp = setup()
    from oracle.problems import articulated_problem

    G = {}
    with ref_loader.reference_cwd(tmp):
        for mt in ("mano", "flame"):
            model, tgt, idx, init, B = articulated_problem(mt, 4, seed=710)
            res = ref.optimize_params_sequence(tgt.numpy(), body_model=mt, model=model,
                                               config=dict(frame=dict(use_lbfgs=False)))
            names = [k for k in init if init[k].shape[-1] > 0]
            for k in names:
                G[f"api_{mt}_{k}"] = torch.cat([getattr(r.params, k) for r in res])
            G[f"api_{mt}_joints"] = torch.cat([r.joints for r in res])
            G[f"api_{mt}_loss"] = torch.stack([r.loss.reshape(()) for r in res])
            print(mt, "loss", [round(float(r.loss), 2) for r in res], flush=True)
        for mt in ("smplx", "smplh"):
            model, tgt, idx, init, B = articulated_problem(mt, 3, seed=711)
            blocks = {"body": tgt[:, :22].numpy(), "left_hand": tgt[:, 22:43].numpy(), "right_hand": tgt[:, 43:64].numpy()}
            if mt == "smplx":
                blocks["face"] = tgt[:, 64:84].numpy()
            if mt == "smplx":
                res = ref.optimize_params_sequence(blocks, body_model=mt, model=model,
                                                   config=dict(frame=dict(use_lbfgs=False)))
            else:
                res = [ref.optimize_params_frame({k: v[0] for k, v in blocks.items()}, body_model=mt, model=model,
                                                 config=dict(use_lbfgs=False))]
            for k in init:
                v0 = getattr(res[0].params, k, None)
                if v0 is not None:
                    G[f"api_{mt}_{k}"] = torch.cat([getattr(r.params, k) for r in res])
            G[f"api_{mt}_joints"] = torch.cat([r.joints for r in res])
            G[f"api_{mt}_loss"] = torch.stack([r.loss.reshape(()) for r in res])
            print(mt, "loss", [round(float(r.loss), 2) for r in res], flush=True)
    save("r2_generic_api.npz", G)


def section_camera_seq():
    """optimize_params_sequence(coordinate_mode="camera") of the unmodified reference on a smooth 12-frame sequence
    (Adam and L-BFGS, 15 iterations per stage): the loop api/sequence.py:214-281 over CameraSpaceFitter.fit_frame."""
    ref, tmp = setup()
    G = {}
    weights = syn.make_body_model("smpl", seed=0)
    model = BodyModelShim(weights)
    from oracle.problems import chain_problem

    tgt = chain_problem(weights, 1, 12, seed=77)[0]
    G["camseq_in_target"] = tgt
    with ref_loader.reference_cwd(tmp):
        for name, lb in (("camseq_adam", False), ("camseq_lbfgs", True)):
            res = ref.optimize_params_sequence(tgt.numpy(), body_model="smpl", joint_layout="AMASS", model=model,
                                               config=dict(frame=dict(use_lbfgs=lb, coordinate_mode="camera", num_iters=15),
                                                           use_shape_optimization=False))
            G[name + "_pose"] = torch.cat([r.params.pose for r in res])
            G[name + "_transl"] = torch.cat([r.params.transl for r in res])
            G[name + "_betas"] = torch.cat([r.params.betas for r in res])
            G[name + "_joints"] = torch.cat([r.joints for r in res])
            G[name + "_loss"] = torch.stack([r.loss.reshape(()) for r in res])
            print(name, "loss", [round(float(r.loss), 1) for r in res], flush=True)
    save("r2_camera_seq.npz", G)


if __name__ == "__main__":
    torch.set_num_threads(1)
    sys.path.insert(0, HERE)
    {"points": section_points, "dist": section_dist, "dist64": section_dist64, "chains": section_chains,
     "adam": section_adam, "adam_pert": section_adam_pert, "generic": section_generic, "generic_api": section_generic_api, "camera_seq": section_camera_seq}[sys.argv[1]]()
