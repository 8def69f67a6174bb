"""Generate golden vectors from the UNMODIFIED reference (authoring container only).

    python tests/golden/make_goldens.py

Runs /root/reference's own fitter / loss / prior / sequence driver (imported via
``oracle.ref_loader`` stubs, body model = ``oracle.smplx_shim``) on seeded
synthetic inputs and stores inputs + outputs in ``tests/golden/ref_goldens.npz``.
Synthetic weights are regenerated from seeds by the tests; only small arrays are
stored.  Torch runs single-threaded so the L-BFGS traces are reproducible.
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
OUT = os.path.join(os.path.dirname(__file__), "ref_goldens.npz")
G = {}


def put(name, value):
    if isinstance(value, torch.Tensor):
        value = value.detach().cpu().numpy()
    G[name] = np.asarray(value)


def targets_for(model, n_frames, seed, num_obs=22, noise=0.005):
    mo = syn.make_motion(n_frames, seed=seed)
    tgt = syn.kinematic_joints(model, mo["pose"][:, : 3 * 22], mo["betas"], mo["transl"], 22)
    if num_obs == 24:
        tgt = syn.kinematic_joints(model, mo["pose"], mo["betas"], mo["transl"], 24)
    g = torch.Generator().manual_seed(seed + 100)
    tgt = tgt + noise * torch.randn(tgt.shape, generator=g)
    return mo, tgt


def main():
    ref = ref_loader.load_reference()
    from keypoints2body.core.fitters.world_space import WorldSpaceFitter
    from keypoints2body.core.losses import body_fitting_loss_3d
    from keypoints2body.core.prior import MaxMixturePrior
    from keypoints2body.core.shape import optimize_shape_multi_frame
    from keypoints2body.models.smpl_data import SMPLData, SMPLHData, SMPLXData

    tmp = tempfile.mkdtemp()
    syn.write_assets(os.path.join(tmp, "data/models"), seed=0)
    models = {mt: BodyModelShim(syn.make_body_model(mt, seed=0)) for mt in ("smpl", "smplh", "smplx")}

    with ref_loader.reference_cwd(tmp):
        prior = MaxMixturePrior(prior_folder="./data/models/", num_gaussians=8, dtype=torch.float32)
        put("prior_means", prior.means)
        put("prior_precisions", prior.precisions)
        put("prior_nll_weights", prior.nll_weights)

        # ---- A. single-evaluation goldens (loss, grads, joints) -------------
        for mt in ("smpl", "smplh", "smplx"):
            model = models[mt]
            for num_obs in ((22, 24) if mt == "smpl" else (22,)):
                B = 6
                mo, tgt = targets_for(model, B, seed=1, num_obs=num_obs)
                g = torch.Generator().manual_seed(7)
                pose = mo["pose"] + 0.05 * torch.randn(B, 72, generator=g)
                pose[0] = 0.0                       # exactly r = 0 everywhere
                pose[1, 3:9] = 0.0                  # mixed zero / non-zero joints
                pose[2] = 0.3 * torch.randn(72, generator=g)  # far from target
                betas = mo["betas"] + 0.1 * torch.randn(B, 10, generator=g)
                transl = mo["transl"] + 0.02 * torch.randn(B, 3, generator=g)
                conf = 0.5 + torch.rand(num_obs, generator=g)
                keep = pose[:, 3:] + 0.1 * torch.randn(B, 69, generator=g)
                extra = {}
                if mt in ("smplh", "smplx"):
                    extra["left_hand_pose"] = 0.1 * torch.randn(B, 45, generator=g)
                    extra["right_hand_pose"] = 0.1 * torch.randn(B, 45, generator=g)
                if mt == "smplx":
                    extra["expression"] = 0.5 * torch.randn(B, 10, generator=g)
                    extra["jaw_pose"] = 0.1 * torch.randn(B, 3, generator=g)
                    extra["leye_pose"] = 0.1 * torch.randn(B, 3, generator=g)
                    extra["reye_pose"] = 0.1 * torch.randn(B, 3, generator=g)
                for w_keep in (0.0, 5.0):
                    p = dict(global_orient=pose[:, :3].clone(), body_pose=pose[:, 3:].clone(),
                             betas=betas.clone(), transl=transl.clone(), **{k: v.clone() for k, v in extra.items()})
                    for v in p.values():
                        v.requires_grad_(True)
                    out = model(**p)
                    idx = torch.arange(num_obs)
                    # per-frame loss: call the reference loss frame by frame
                    losses = []
                    for b in range(B):
                        lb = body_fitting_loss_3d(
                            body_pose=p["body_pose"][b:b + 1], preserve_pose=keep[b:b + 1],
                            betas=p["betas"][b:b + 1], model_joints=out.joints[b:b + 1, idx],
                            j3d=tgt[b:b + 1], pose_prior=prior, joints3d_conf=conf,
                            joint_loss_weight=600.0, pose_preserve_weight=w_keep)
                        losses.append(lb)
                    torch.stack(losses).sum().backward()
                    tag = f"eval_{mt}_{num_obs}_w{int(w_keep)}"
                    put(tag + "_loss", torch.stack(losses))
                    for k, v in p.items():
                        put(f"{tag}_in_{k}", v)
                        put(f"{tag}_grad_{k}", v.grad)
                    put(tag + "_in_keep", keep)
                    put(tag + "_in_conf", conf)
                    put(tag + "_in_target", tgt)
                    put(tag + "_joints", out.joints[:, :24])

        # ---- B/C. fitter goldens (Adam strict, L-BFGS traced) ---------------
        def run_fitter(mt, use_lbfgs, n_first, n_follow, seq_ind, freeze, B, seed, num_obs=22, tag=""):
            model = models[mt]
            mo, tgt = targets_for(model, B, seed=seed, num_obs=num_obs)
            g = torch.Generator().manual_seed(seed + 5)
            pose = mo["pose"] + 0.1 * torch.randn(B, 72, generator=g)
            betas = torch.zeros(B, 10)
            transl = mo["transl"] + 0.03 * torch.randn(B, 3, generator=g)
            fitter = WorldSpaceFitter(model, step_size=1e-2, num_iters_first=n_first,
                                      num_iters_followup=n_follow, use_lbfgs=use_lbfgs,
                                      joints_category="AMASS" if num_obs == 22 else "SMPL24")
            base = dict(betas=betas, global_orient=pose[:, :3], body_pose=pose[:, 3:], transl=transl)
            if mt == "smpl":
                init = SMPLData(**base)
            elif mt == "smplh":
                init = SMPLHData(**base, left_hand_pose=torch.zeros(B, 45), right_hand_pose=torch.zeros(B, 45))
            else:
                init = SMPLXData(**base, left_hand_pose=torch.zeros(B, 45), right_hand_pose=torch.zeros(B, 45),
                                 expression=torch.zeros(B, 10), jaw_pose=torch.zeros(B, 3),
                                 leye_pose=torch.zeros(B, 3), reye_pose=torch.zeros(B, 3))
            conf = torch.ones(num_obs)
            put(tag + "_in_pose", pose)
            put(tag + "_in_betas", betas)
            put(tag + "_in_transl", transl)
            put(tag + "_in_target", tgt)
            if not use_lbfgs:
                res = fitter.fit_frame(init, tgt, conf, seq_ind=seq_ind, freeze_betas=freeze)
                put(tag + "_pose", res.params.pose)
                put(tag + "_betas", res.params.betas)
                put(tag + "_transl", res.params.transl)
                put(tag + "_joints", res.joints)
                put(tag + "_loss", res.loss)
                if mt == "smplx":
                    put(tag + "_expression", res.params.expression)
                    put(tag + "_lh", res.params.left_hand_pose)
                    put(tag + "_jaw", res.params.jaw_pose)
                put(tag + "_verts0", res.vertices[0])
                return
            # L-BFGS: frame by frame (B=1 each, as the API does) with a line-search trace
            import torch.optim.lbfgs as L
            outs = {k: [] for k in ("pose", "betas", "transl", "joints", "loss", "nevals")}
            traces = []
            ls_records = []
            for b in range(B):
                tr = []
                ls = []
                orig = L.LBFGS._directional_evaluate
                orig_sw = L._strong_wolfe

                def rec_sw(obj_func, x, t, d, f, g, gtd, c1=1e-4, c2=0.9, tolerance_change=1e-9,
                           max_ls=25, _o=orig_sw, _ls=ls, _tr=tr):
                    start = len(_tr)
                    out = _o(obj_func, x, t, d, f, g, gtd, c1, c2, tolerance_change, max_ls)
                    # t_in, f, gtd, d_norm, max_ls, first trace row, n_evals, t_out, f_out, t_is_tensor
                    _ls.append((float(t), float(f), float(gtd), float(d.abs().max()), max_ls, start,
                                out[3], float(out[2]), float(out[0]), float(isinstance(t, torch.Tensor))))
                    return out

                L._strong_wolfe = rec_sw

                def rec(self, closure, x, t, d, _orig=orig, _tr=tr):
                    loss, fg = _orig(self, closure, x, t, d)
                    _tr.append((float(t), loss, float(fg.dot(d))))
                    return loss, fg

                L.LBFGS._directional_evaluate = rec
                try:
                    sub = type(init)(**{k: (v[b:b + 1] if isinstance(v, torch.Tensor) else v)
                                        for k, v in init.__dict__.items() if k != "metadata"})
                    res = fitter.fit_frame(sub, tgt[b:b + 1], conf, seq_ind=seq_ind, freeze_betas=freeze)
                finally:
                    L.LBFGS._directional_evaluate = orig
                    L._strong_wolfe = orig_sw
                ls_arr = np.full((40, 10), np.nan)
                ls_arr[: len(ls)] = np.asarray(ls, dtype=np.float64)
                ls_records.append(ls_arr)
                outs["pose"].append(res.params.pose)
                outs["betas"].append(res.params.betas)
                outs["transl"].append(res.params.transl)
                outs["joints"].append(res.joints)
                outs["loss"].append(res.loss.reshape(1))
                outs["nevals"].append(torch.tensor([len(tr) + 1]))
                t_arr = np.full((48, 3), np.nan)
                t_arr[: len(tr)] = np.asarray(tr)
                traces.append(t_arr)
            for k, v in outs.items():
                put(f"{tag}_{k}", torch.cat(v))
            put(tag + "_trace", np.stack(traces))
            put(tag + "_linesearch", np.stack(ls_records))

        for n in (5, 10, 30):
            run_fitter("smpl", False, n, n, 0, False, 4, seed=11, tag=f"adam_smpl_n{n}")
        run_fitter("smpl", False, 30, 10, 3, False, 4, seed=12, tag="adam_smpl_follow")
        run_fitter("smpl", False, 30, 10, 0, True, 4, seed=13, tag="adam_smpl_freeze")
        run_fitter("smpl", False, 10, 10, 0, False, 4, seed=14, num_obs=24, tag="adam_smpl24")
        run_fitter("smplh", False, 10, 10, 2, False, 3, seed=15, tag="adam_smplh")
        run_fitter("smplx", False, 5, 5, 0, False, 3, seed=16, tag="adam_smplx")
        run_fitter("smpl", True, 30, 10, 0, False, 4, seed=21, tag="lbfgs_smpl_first")
        run_fitter("smpl", True, 30, 10, 2, False, 4, seed=22, tag="lbfgs_smpl_follow")
        run_fitter("smplx", True, 5, 5, 0, False, 3, seed=23, tag="lbfgs_smplx")

        # ---- D. public sequence API (S1 chain, S0 independent, shape pass) --
        model = models["smpl"]
        mo, tgt = targets_for(model, 6, seed=31)
        put("seq_in_target", tgt)
        for name, cfg in (
            ("seq_adam_chain", dict(frame=dict(use_lbfgs=False), use_shape_optimization=False)),
            ("seq_adam_indep", dict(frame=dict(use_lbfgs=False), use_shape_optimization=False,
                                    use_previous_frame_init=False)),
            ("seq_lbfgs_shape", dict(frame=dict(use_lbfgs=True), use_shape_optimization=True,
                                     num_shape_frames=4, num_shape_iters=10)),
        ):
            res = ref.optimize_params_sequence(tgt.numpy(), body_model="smpl", joint_layout="AMASS",
                                               model=model, config=cfg)
            put(name + "_pose", torch.cat([r.params.pose for r in res]))
            put(name + "_betas", torch.cat([r.params.betas for r in res]))
            put(name + "_transl", torch.cat([r.params.transl for r in res]))
            put(name + "_joints", torch.cat([r.joints for r in res]))
            put(name + "_loss", torch.stack([r.loss.reshape(()) for r in res]))
        # frame API, defaults (L-BFGS) and Adam
        for name, cfg in (("frame_lbfgs", None), ("frame_adam", dict(use_lbfgs=False))):
            r = ref.optimize_params_frame(tgt[0].numpy(), body_model="smpl", joint_layout="AMASS",
                                          model=model, config=cfg)
            put(name + "_pose", r.params.pose)
            put(name + "_betas", r.params.betas)
            put(name + "_transl", r.params.transl)
            put(name + "_loss", r.loss)

        # ---- F. camera-space two-stage fitter (reference CameraSpaceFitter, B = 1 per frame) -------
        from keypoints2body.core.fitters.camera_space import CameraSpaceFitter

        model = models["smpl"]
        mo, tgtc = targets_for(model, 3, seed=41)
        gcam = torch.Generator().manual_seed(42)
        posec = mo["pose"] + 0.1 * torch.randn(3, 72, generator=gcam)
        put("cam_in_target", tgtc)
        put("cam_in_pose", posec)
        for name, lb, iters, seq_ind, freeze in (("cam_adam", False, 15, 0, True), ("cam_adam_follow", False, 15, 2, True),
                                                  ("cam_lbfgs", True, 20, 0, True)):
            fit = CameraSpaceFitter(model, step_size=1e-2, num_iters=iters, use_lbfgs=lb, joints_category="AMASS")
            outs = {k: [] for k in ("pose", "betas", "transl", "joints", "loss")}
            for b in range(3):
                init = SMPLData(betas=torch.zeros(1, 10), global_orient=posec[b:b + 1, :3], body_pose=posec[b:b + 1, 3:])
                r = fit.fit_frame(init, tgtc[b:b + 1], torch.ones(22), seq_ind=seq_ind, freeze_betas=freeze)
                outs["pose"].append(r.params.pose); outs["betas"].append(r.params.betas)
                outs["transl"].append(r.params.transl); outs["joints"].append(r.joints)
                outs["loss"].append(r.loss.reshape(1))
            for k, v in outs.items():
                put(f"{name}_{k}", torch.cat(v))
        # caller-supplied init_cam_t (camera_space.py:91,133): start AND depth reference, away from the
        # stage-0 stationary point, so Adam's first step is not decided by rounding noise
        with torch.no_grad():
            j0 = model(global_orient=posec[:, :3], body_pose=posec[:, 3:], betas=torch.zeros(3, 10)).joints
        sel = [2, 1, 17, 16]
        given = (tgtc[:, sel] - j0[:, sel]).sum(dim=1) / 4.0 + torch.tensor([[0.012, -0.02, 0.016]])
        put("cam_given_init", given)
        for name, seq_ind in (("cam_given_adam", 0), ("cam_given_adam_follow", 2)):
            fit = CameraSpaceFitter(model, step_size=1e-2, num_iters=15, use_lbfgs=False, joints_category="AMASS")
            outs = {k: [] for k in ("pose", "betas", "transl", "joints", "loss")}
            for b in range(3):
                init = SMPLData(betas=torch.zeros(1, 10), global_orient=posec[b:b + 1, :3], body_pose=posec[b:b + 1, 3:])
                r = fit.fit_frame(init, tgtc[b:b + 1], torch.ones(22), seq_ind=seq_ind, freeze_betas=True,
                                  init_cam_t=given[b:b + 1])
                outs["pose"].append(r.params.pose); outs["betas"].append(r.params.betas)
                outs["transl"].append(r.params.transl); outs["joints"].append(r.joints)
                outs["loss"].append(r.loss.reshape(1))
            for k, v in outs.items():
                put(f"{name}_{k}", torch.cat(v))
        r = ref.optimize_params_frame(tgtc[0].numpy(), body_model="smpl", joint_layout="AMASS", model=model,
                                      config=dict(use_lbfgs=False, coordinate_mode="camera", num_iters=10))
        put("frame_cam_adam_pose", r.params.pose)
        put("frame_cam_adam_transl", r.params.transl)
        put("frame_cam_adam_loss", r.loss)

        # ---- E. shape pass alone --------------------------------------------
        b = optimize_shape_multi_frame(model, init_betas=torch.zeros(1, 10),
                                       pose_init=torch.zeros(6, 72), j3d_world=tgt,
                                       joints_category="AMASS", num_iters=40, step_size=1e-1,
                                       use_lbfgs=True, frame_indices=list(range(5)),
                                       joints3d_conf=torch.ones(22), shape_prior_weight=5.0)
        put("shape_pass_betas", b)

        # ---- G. MPJAE evaluation (cli/eval.py) and the on-disk formats (io/motion.py) -----------------
        from keypoints2body.cli.eval import compute_angular_error_deg, evaluate_pose_pair
        from keypoints2body.io.motion import load_motion_data, write_smplx_zip

        ge = np.random.default_rng(11)
        gt_pose = (0.6 * ge.standard_normal((17, 72))).astype(np.float32)
        pred_pose = (gt_pose[:, :66] + 0.05 * ge.standard_normal((17, 66))).astype(np.float32)
        pred_pose[0, :6] = gt_pose[0, :6]            # identical rotations -> clipped cosine
        pred_pose[1, :3] = 0.0                       # zero rotation vector (Taylor branch)
        gt_pose[1, 3:6] = 0.0
        pred_pose[1, 3:6] = 0.0
        gt_pose[2, :3] = np.array([3.1, 0.0, 0.0])   # near pi
        pred_pose[2, :3] = np.array([-3.1, 0.05, 0.0])
        pred_pose = pred_pose[:15]                   # fewer predicted frames than ground truth
        put("mpjae_in_pred", pred_pose)
        put("mpjae_in_gt", gt_pose)
        put("mpjae_angles", compute_angular_error_deg(pred_pose.reshape(15, 22, 3), gt_pose[:15, :66].reshape(15, 22, 3)))
        mean, total, count = evaluate_pose_pair(pred_pose, gt_pose)
        put("mpjae_summary", np.array([mean, total, count], np.float64))

        io_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "io")
        os.makedirs(io_dir, exist_ok=True)
        seq22 = (0.3 * ge.standard_normal((4, 22, 3))).astype(np.float32)
        seq25 = (0.3 * ge.standard_normal((3, 25, 3))).astype(np.float64)
        np.save(os.path.join(io_dir, "seq22.npy"), seq22)
        np.savez(os.path.join(io_dir, "seq25.npz"), joints=seq25)
        with open(os.path.join(io_dir, "seq22.csv"), "w") as fh:
            for h in range(5):
                fh.write(f"header line {h},,\n")
            for t in range(4):
                fh.write(",".join([str(t), f"{t / 120.0:.6f}"] + [repr(float(v)) for v in seq22[t].reshape(-1)]) + "\n")
        import warnings as _w
        with _w.catch_warnings():
            _w.simplefilter("ignore")
            for name, layout in (("seq22.npy", None), ("seq22.csv", "AMASS"), ("seq25.npz", None)):
                j, lay, k = load_motion_data(__import__("pathlib").Path(io_dir) / name, layout)
                put("io_" + name.replace(".", "_") + "_joints", np.asarray(j, np.float64))
                G["io_" + name.replace(".", "_") + "_layout"] = np.array([lay, str(k)])
        zp = write_smplx_zip(__import__("pathlib").Path(io_dir), gt_pose[:3].astype(np.float64), np.linspace(-1, 1, 10),
                             (0.1 * ge.standard_normal((3, 3))), zip_name="ref_params.zip", person_idx=1)
        print("reference zip:", zp, os.path.getsize(zp), "bytes")

    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT) / 1e3, "KB,", len(G), "arrays")


if __name__ == "__main__":
    main()
