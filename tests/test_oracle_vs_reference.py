"""Pins the oracle (oracle/reference_port.py + oracle/smplx_shim.py) against golden
vectors produced by the UNMODIFIED reference (tests/golden/make_goldens.py)."""

import numpy as np
import pytest
import torch

from oracle import reference_port as rp

T = torch.as_tensor


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def test_prior_constants_match_reference(goldens, oracle_prior):
    np.testing.assert_array_equal(oracle_prior.means.numpy(), goldens["prior_means"])
    np.testing.assert_array_equal(oracle_prior.precisions.numpy(), goldens["prior_precisions"])
    np.testing.assert_array_equal(oracle_prior.nll_weights.numpy(), goldens["prior_nll_weights"])


EXTRA = ("left_hand_pose", "right_hand_pose", "expression", "jaw_pose", "leye_pose", "reye_pose")


@pytest.mark.parametrize("tag", ["eval_smpl_22_w0", "eval_smpl_22_w5", "eval_smpl_24_w5",
                                 "eval_smplh_22_w5", "eval_smplx_22_w0", "eval_smplx_22_w5"])
def test_evaluation_matches_reference(goldens, shims, oracle_prior, tag):
    _, mt, nobs, w = tag.split("_")
    g = goldens
    params = {k: None for k in rp.PARAM_ORDER}
    for k in rp.PARAM_ORDER:
        if f"{tag}_in_{k}" in g:
            params[k] = T(g[f"{tag}_in_{k}"])
    loss, grads, joints = rp.evaluate(shims(mt), oracle_prior, params, T(g[tag + "_in_keep"]),
                                      T(g[tag + "_in_target"]), T(g[tag + "_in_conf"]),
                                      num_obs=int(nobs), pose_preserve_weight=float(w[1:]))
    np.testing.assert_allclose(loss.numpy(), g[tag + "_loss"].reshape(-1), rtol=1e-6)
    np.testing.assert_allclose(joints[:, :24].numpy(), g[tag + "_joints"], atol=1e-6)
    for k, v in grads.items():
        if v is None:
            continue
        ref = g[f"{tag}_grad_{k}"]
        np.testing.assert_allclose(v.numpy(), ref, atol=1e-5 * max(1.0, np.abs(ref).max()))


ADAM_CASES = {
    "adam_smpl_n5": ("smpl", 5, 5, 0, False, 22), "adam_smpl_n10": ("smpl", 10, 10, 0, False, 22),
    "adam_smpl_n30": ("smpl", 30, 30, 0, False, 22), "adam_smpl_follow": ("smpl", 30, 10, 3, False, 22),
    "adam_smpl_freeze": ("smpl", 30, 10, 0, True, 22), "adam_smpl24": ("smpl", 10, 10, 0, False, 24),
    "adam_smplh": ("smplh", 10, 10, 2, False, 22), "adam_smplx": ("smplx", 5, 5, 0, False, 22),
}


def _init_from_golden(g, tag, mt):
    pose, B = T(g[tag + "_in_pose"]), g[tag + "_in_pose"].shape[0]
    init = {k: None for k in rp.PARAM_ORDER}
    init.update(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=T(g[tag + "_in_betas"]),
                transl=T(g[tag + "_in_transl"]))
    if mt in ("smplh", "smplx"):
        init.update(left_hand_pose=torch.zeros(B, 45), right_hand_pose=torch.zeros(B, 45))
    if mt == "smplx":
        init.update(expression=torch.zeros(B, 10), jaw_pose=torch.zeros(B, 3),
                    leye_pose=torch.zeros(B, 3), reye_pose=torch.zeros(B, 3))
    return init


@pytest.mark.parametrize("tag", sorted(ADAM_CASES))
def test_adam_fit_matches_reference(goldens, shims, oracle_prior, tag):
    mt, n1, n2, seq_ind, freeze, nobs = ADAM_CASES[tag]
    g = goldens
    out = rp.fit_frame(shims(mt), oracle_prior, _init_from_golden(g, tag, mt), T(g[tag + "_in_target"]),
                       torch.ones(nobs), seq_ind=seq_ind, num_obs=nobs, use_lbfgs=False,
                       num_iters_first=n1, num_iters_followup=n2, freeze_betas=freeze)
    p = out["params"]
    pose = torch.cat([p["global_orient"], p["body_pose"]], dim=1)
    np.testing.assert_allclose(pose.numpy(), g[tag + "_pose"], atol=2e-6)
    np.testing.assert_allclose(p["betas"].numpy(), g[tag + "_betas"], atol=2e-6)
    np.testing.assert_allclose(p["transl"].numpy(), g[tag + "_transl"], atol=2e-6)
    np.testing.assert_allclose(out["joints"].numpy(), g[tag + "_joints"], atol=2e-6)
    np.testing.assert_allclose(float(out["loss"]), float(g[tag + "_loss"]), rtol=1e-5)
    np.testing.assert_allclose(out["vertices"][0].numpy(), g[tag + "_verts0"], atol=2e-6)
    if mt == "smplx":
        np.testing.assert_allclose(p["expression"].numpy(), g[tag + "_expression"], atol=2e-6)
        np.testing.assert_allclose(p["left_hand_pose"].numpy(), g[tag + "_lh"], atol=1e-7)


LBFGS_CASES = {"lbfgs_smpl_first": ("smpl", 30, 10, 0), "lbfgs_smpl_follow": ("smpl", 30, 10, 2),
               "lbfgs_smplx": ("smplx", 5, 5, 0)}


@pytest.mark.parametrize("tag", sorted(LBFGS_CASES))
def test_lbfgs_fit_matches_reference(goldens, shims, oracle_prior, tag):
    """Same torch build, same thread count, same op sequence -> the port reproduces the
    reference's L-BFGS runs (chaotic under perturbation) closely."""
    mt, n1, n2, seq_ind = LBFGS_CASES[tag]
    g = goldens
    init = _init_from_golden(g, tag, mt)
    B = g[tag + "_in_pose"].shape[0]
    for b in range(B):
        sub = {k: (v[b:b + 1] if v is not None else None) for k, v in init.items()}
        tr = []
        out = rp.fit_frame(shims(mt), oracle_prior, sub, T(g[tag + "_in_target"][b:b + 1]), torch.ones(22),
                           seq_ind=seq_ind, use_lbfgs=True, num_iters_first=n1, num_iters_followup=n2,
                           trace=tr)
        assert len(tr) == int(g[tag + "_nevals"][b])
        ref_f = g[tag + "_trace"][b][: len(tr) - 1, 1]
        np.testing.assert_allclose([f for f, _ in tr[1:]], ref_f, rtol=1e-5)
        np.testing.assert_allclose(float(out["loss"]), float(g[tag + "_loss"][b]), rtol=1e-4)
        p = out["params"]
        pose = torch.cat([p["global_orient"], p["body_pose"]], dim=1)
        np.testing.assert_allclose(pose.numpy(), g[tag + "_pose"][b:b + 1], atol=1e-4)
        np.testing.assert_allclose(out["joints"].numpy(), g[tag + "_joints"][b:b + 1], atol=1e-4)


def test_shape_pass_matches_reference(goldens, shims):
    g = goldens
    b = rp.optimize_shape(shims("smpl"), torch.zeros(1, 10), torch.zeros(6, 72), T(g["seq_in_target"]),
                          torch.ones(22), frame_indices=list(range(5)), num_iters=40)
    np.testing.assert_allclose(b.numpy(), g["shape_pass_betas"], atol=1e-5)


CAM_CASES = {"cam_adam": (False, 15, 0), "cam_adam_follow": (False, 15, 2), "cam_lbfgs": (True, 20, 0),
             "cam_given_adam": (False, 15, 0), "cam_given_adam_follow": (False, 15, 2)}


@pytest.mark.parametrize("tag", sorted(CAM_CASES))
def test_camera_fitter_matches_reference(goldens, shims, oracle_prior, tag):
    """Camera-space two-stage fitter (core/fitters/camera_space.py) restated in the oracle."""
    lbfgs, iters, seq_ind = CAM_CASES[tag]
    g = goldens
    pose, tgt = T(g["cam_in_pose"]), T(g["cam_in_target"])
    for b in range(3):
        init = dict(global_orient=pose[b:b + 1, :3], body_pose=pose[b:b + 1, 3:], betas=torch.zeros(1, 10))
        out = rp.fit_frame_camera(shims("smpl"), oracle_prior, init, tgt[b:b + 1], torch.ones(22), seq_ind=seq_ind,
                                  use_lbfgs=lbfgs, num_iters=iters,
                                  init_cam_t=T(g["cam_given_init"][b:b + 1]) if "given" in tag else None)
        p = out["params"]
        tol = 1e-4 if lbfgs else 2e-6
        np.testing.assert_allclose(torch.cat([p["global_orient"], p["body_pose"]], 1).numpy(), g[tag + "_pose"][b:b + 1], atol=tol)
        np.testing.assert_allclose(p["transl"].numpy(), g[tag + "_transl"][b:b + 1], atol=tol)
        np.testing.assert_allclose(out["joints"].numpy(), g[tag + "_joints"][b:b + 1], atol=tol)
        np.testing.assert_allclose(float(out["loss"]), float(g[tag + "_loss"][b]), rtol=1e-4)


def test_explicit_model_indices_equal_zero_confidence_slots(generic_goldens, shims, oracle_prior, weights):
    """The reference's GENERIC path (explicit ``target_model_indices``, world_space.py:198-201) sums the loss over the
    observed joints only.  The CUDA path keeps fixed observation slots and gives unobserved joints confidence 0;
    this pins that the two are the same fit: the oracle run on the scattered observations reproduces the
    reference's golden for a partial, permuted index set."""
    from keypoints2body_b200 import synthetic as syn

    g = generic_goldens
    idx = torch.as_tensor(g["gen_idx"]).long()
    block = torch.as_tensor(g["gen_in_block"])
    full = torch.zeros(1, 22, 3)
    full[:, idx] = block[1:2, idx, :3]
    conf = torch.zeros(22)
    conf[idx] = block[1, idx, 3]
    pose = torch.as_tensor(g["gen_idx_in_pose"])
    init = {k: None for k in rp.PARAM_ORDER}
    init.update(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=torch.zeros(1, 10),
                transl=torch.as_tensor(g["gen_idx_in_transl"]))
    out = rp.fit_frame(shims("smpl"), oracle_prior, init, full, conf, seq_ind=0, use_lbfgs=False, num_iters_first=12)
    got = torch.cat([out["params"]["global_orient"], out["params"]["body_pose"]], dim=1).numpy()
    assert np.abs(got - g["gen_idx_pose"]).max() < 2e-6
    assert np.abs(out["params"]["transl"].numpy() - g["gen_idx_transl"]).max() < 2e-6
    np.testing.assert_allclose(float(out["loss"]), float(g["gen_idx_loss"]), rtol=1e-5)
