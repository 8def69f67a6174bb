"""GPU parity tests: the CUDA path (through the C ABI in include/k2b_b200.h) vs the oracle
(oracle/reference_port.py + oracle/smplx_shim.py, pinned to the unmodified reference by
tests/test_oracle_vs_reference.py) and vs the committed reference goldens.

Tolerances (BASELINE.md section 3.5 / SURVEY.md section 8c):
  G1 evaluation : loss rel <= 1e-5, gradient <= 1e-4 of its max-abs, joints <= 1e-5 m
  G2 Adam       : joints <= 1e-4 m, pose <= 1e-4 rad, betas <= 1e-4, transl <= 1e-5 m, loss rel <= 1e-4
  G4 L-BFGS     : statistical (the reference disagrees with itself by centimetres)
  G5 mesh       : vertices / joints <= 1e-4 m (measured ~1e-6)
"""

import numpy as np
import pytest
import torch

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu
T = torch.as_tensor


@pytest.fixture(scope="module")
def fitters(weights, gmm):
    from keypoints2body_b200.core.fitters.world_space import WorldSpaceFitter

    cache = {}

    def get(mt, cat="AMASS", **kw):
        key = (mt, cat, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = WorldSpaceFitter(weights(mt), joints_category=cat, model_type=mt, gmm=gmm, **kw)
        cache[key].warp_kernel_max_frames = 0      # one thread per frame unless a test asks for the warp kernel
        return cache[key]

    return get


KERNELS = ("frame", "warp")      # k2b_fit_batch (one thread per frame) / k2b_fit_chain (one warp per frame)


def cpu(x):
    return x.detach().cpu().numpy()


EXTRA = ("left_hand_pose", "right_hand_pose", "expression", "jaw_pose", "leye_pose", "reye_pose")


@pytest.mark.parametrize("tag", ["eval_smpl_22_w0", "eval_smpl_22_w5", "eval_smpl_24_w5",
                                 "eval_smplh_22_w5", "eval_smplx_22_w0", "eval_smplx_22_w5"])
def test_evaluation_vs_reference_goldens(goldens, fitters, tag):
    _, mt, nobs, w = tag.split("_")
    g = goldens
    f = fitters(mt, "SMPL24" if nobs == "24" else "AMASS")
    params = {k: T(g[f"{tag}_in_{k}"]) for k in rp.PARAM_ORDER if f"{tag}_in_{k}" in g}
    keep_w = float(w[1:])
    out = f.evaluate_batch(params, T(g[tag + "_in_target"]), T(g[tag + "_in_conf"]),
                           preserve_pose=T(g[tag + "_in_keep"]), preserve_on=keep_w > 0,
                           pose_preserve_weight=keep_w)
    np.testing.assert_allclose(cpu(out["loss"]), g[tag + "_loss"].reshape(-1), rtol=1e-5)
    np.testing.assert_allclose(cpu(out["joints"]), g[tag + "_joints"][:, : int(nobs)], atol=1e-5)
    ref_pose = np.concatenate([g[tag + "_grad_global_orient"], g[tag + "_grad_body_pose"]], axis=1)
    assert (np.abs(cpu(out["grad_pose"]) - ref_pose) / np.abs(ref_pose).max(axis=1, keepdims=True)).max() < 1e-4
    for name in ("transl", "betas") + (("expression",) if mt == "smplx" else ()):
        ref = g[f"{tag}_grad_{name}"]
        got = cpu(out["grad_" + name])
        assert (np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max() < 1e-4, name


ADAM_CASES = {
    "adam_smpl_n5": ("smpl", 5, 0, False, 22), "adam_smpl_n10": ("smpl", 10, 0, False, 22),
    "adam_smpl_n30": ("smpl", 30, 0, False, 22), "adam_smpl_follow": ("smpl", 10, 3, False, 22),
    "adam_smpl_freeze": ("smpl", 30, 0, True, 22), "adam_smpl24": ("smpl", 10, 0, False, 24),
    "adam_smplh": ("smplh", 10, 2, False, 22), "adam_smplx": ("smplx", 5, 0, False, 22),
}


def golden_init(g, tag, mt):
    pose, B = T(g[tag + "_in_pose"]), g[tag + "_in_pose"].shape[0]
    init = dict(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=T(g[tag + "_in_betas"]),
                transl=T(g[tag + "_in_transl"]))
    if mt in ("smplh", "smplx"):
        init.update(left_hand_pose=torch.zeros(B, 45), right_hand_pose=torch.zeros(B, 45))
    if mt == "smplx":
        init.update(expression=torch.zeros(B, 10), jaw_pose=torch.zeros(B, 3), leye_pose=torch.zeros(B, 3),
                    reye_pose=torch.zeros(B, 3))
    return init


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("tag", sorted(ADAM_CASES))
def test_adam_fit_vs_reference_goldens(goldens, fitters, tag, kernel):
    mt, iters, seq_ind, freeze, nobs = ADAM_CASES[tag]
    g = goldens
    f = fitters(mt, "SMPL24" if nobs == 24 else "AMASS", use_lbfgs=False)
    out = f.fit_batch(golden_init(g, tag, mt), T(g[tag + "_in_target"]), torch.ones(nobs), seq_ind=seq_ind,
                      num_iters=iters, freeze_betas=freeze, kernel=kernel)
    p = out["params"]
    pose = np.concatenate([cpu(p["global_orient"]), cpu(p["body_pose"])], axis=1)
    assert np.abs(pose - g[tag + "_pose"]).max() < 1e-4
    assert np.abs(cpu(p["transl"]) - g[tag + "_transl"]).max() < 1e-5
    assert np.abs(cpu(p["betas"]) - g[tag + "_betas"]).max() < 1e-4
    assert np.abs(cpu(out["joints"]) - g[tag + "_joints"]).max() < 1e-4          # full joints incl. extras
    assert np.abs(cpu(out["fit_joints"]) - g[tag + "_joints"][:, :nobs]).max() < 1e-4
    assert np.abs(cpu(out["vertices"][0]) - g[tag + "_verts0"]).max() < 1e-4
    np.testing.assert_allclose(float(out["loss"].sum()), float(g[tag + "_loss"]), rtol=1e-4)
    if mt == "smplx":
        assert np.abs(cpu(p["expression"]) - g[tag + "_expression"]).max() < 1e-4
    assert (cpu(out["evals"]) == iters).all()


def make_problem(weights, n, seed, noise=0.005, init_noise=0.1):
    from keypoints2body_b200 import synthetic as syn

    w = weights("smpl")
    mo = syn.make_motion(n, seed=seed)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    g = torch.Generator().manual_seed(seed + 1)
    tgt = tgt + noise * torch.randn(tgt.shape, generator=g)
    pose = mo["pose"] + init_noise * torch.randn(n, 72, generator=g)
    init = dict(global_orient=pose[:, :3].contiguous(), body_pose=pose[:, 3:].contiguous(),
                betas=torch.zeros(n, 10), transl=mo["transl"] + 0.03 * torch.randn(n, 3, generator=g))
    return tgt, init


@pytest.mark.parametrize("kernel", KERNELS)
def test_adam_fit_vs_oracle_batch(fitters, shims, oracle_prior, weights, kernel):
    """Fresh seeded inputs, B = 96 (ragged vs the 128-frame tile), per-frame confidences."""
    tgt, init = make_problem(weights, 96, seed=101)
    f = fitters("smpl", use_lbfgs=False)
    out = f.fit_batch(init, tgt, torch.ones(22), seq_ind=4, num_iters=10, kernel=kernel)
    full = {k: None for k in rp.PARAM_ORDER}
    full.update(init)
    ref = rp.fit_frame(shims("smpl"), oracle_prior, full, tgt, torch.ones(22), seq_ind=4, use_lbfgs=False,
                       num_iters_followup=10)
    p = out["params"]
    assert (cpu(p["body_pose"]) - ref["params"]["body_pose"].numpy()).__abs__().max() < 1e-4
    assert (cpu(p["transl"]) - ref["params"]["transl"].numpy()).__abs__().max() < 1e-5
    assert (cpu(out["joints"]) - ref["joints"].numpy()).__abs__().max() < 1e-4
    assert (cpu(out["vertices"]) - ref["vertices"].numpy()).__abs__().max() < 1e-4


# L-BFGS: parity is established by tests/test_gpu_lbfgs_parity.py (teacher-forced evaluations, line-search replay on
# the device, 512-fit and 32 x 64-chain distributions against the reference, float64 check).


@pytest.mark.parametrize("fp32_path", [False, True])
@pytest.mark.parametrize("mt,B", [("smpl", 37), ("smplh", 37), ("smplx", 37), ("smpl", 300), ("smplx", 200)])
def test_mesh_vs_shim(fitters, shims, mt, B, fp32_path, monkeypatch):
    """Full mesh output against the smplx restatement: tcgen05 blend + in-place skinning (default) and the
    fused CUDA-core kernel (K2B_MESH_FP32=1), at frame counts that leave partial passes / tiles."""
    monkeypatch.setenv("K2B_MESH_FP32", "1" if fp32_path else "0")
    g = torch.Generator().manual_seed(7)
    params = dict(global_orient=0.3 * torch.randn(B, 3, generator=g), body_pose=0.3 * torch.randn(B, 69, generator=g),
                  betas=torch.randn(B, 10, generator=g), transl=torch.randn(B, 3, generator=g))
    if mt in ("smplh", "smplx"):
        params.update(left_hand_pose=0.2 * torch.randn(B, 45, generator=g),
                      right_hand_pose=0.2 * torch.randn(B, 45, generator=g))
    if mt == "smplx":
        params.update(expression=torch.randn(B, 10, generator=g), jaw_pose=0.2 * torch.randn(B, 3, generator=g),
                      leye_pose=0.2 * torch.randn(B, 3, generator=g), reye_pose=0.2 * torch.randn(B, 3, generator=g))
    out = fitters(mt).forward_batch(params)
    ref = shims(mt)(**params)
    assert (cpu(out["vertices"]) - ref.vertices.numpy()).__abs__().max() < 1e-4
    assert (cpu(out["joints"]) - ref.joints.numpy()).__abs__().max() < 1e-4


@pytest.mark.parametrize("mt", ["smpl", "smplh", "smplx"])
def test_mesh_tensor_core_path_is_deterministic_and_matches_fp32_path(fitters, mt, monkeypatch):
    """The fused tcgen05 blend + skinning kernels (lane = vertex tiles; SMPL 64 frames per pass, SMPL-H / SMPL-X 32)
    against the FP32 CUDA-core path on EVERY vertex of a batch that spans many passes per SM and ends in a partial
    pass, three times over: identical bits every time (no race between the TMA / MMA / epilogue roles), 1e-4 m."""
    B = 148 * 64 * 3 + 37
    g = torch.Generator().manual_seed(11)
    params = dict(global_orient=0.3 * torch.randn(B, 3, generator=g), body_pose=0.3 * torch.randn(B, 69, generator=g),
                  betas=torch.randn(B, 10, generator=g), transl=torch.randn(B, 3, generator=g))
    if mt in ("smplh", "smplx"):
        params.update(left_hand_pose=0.2 * torch.randn(B, 45, generator=g),
                      right_hand_pose=0.2 * torch.randn(B, 45, generator=g))
    if mt == "smplx":
        params.update(expression=torch.randn(B, 10, generator=g), jaw_pose=0.2 * torch.randn(B, 3, generator=g),
                      leye_pose=0.2 * torch.randn(B, 3, generator=g), reye_pose=0.2 * torch.randn(B, 3, generator=g))
    params = {k: v.cuda() for k, v in params.items()}
    f = fitters(mt)
    monkeypatch.setenv("K2B_MESH_FP32", "1")
    ref = f.forward_batch(params)["vertices"].clone()
    monkeypatch.setenv("K2B_MESH_FP32", "0")
    first = None
    for rep in range(3):
        v = f.forward_batch(params)["vertices"]
        d = float((v - ref).abs().max())
        assert d < 1e-4, (mt, rep, d)
        if first is None:
            first = v.clone()
        else:
            assert torch.equal(v, first), (mt, rep)
    print(mt, "tensor-core mesh vs FP32 path, every vertex of", B, "frames:", d)


def test_host_buffer_entry_matches_device_entry(fitters, weights):
    """k2b_fit_batch_host (host pointers, copies inside) == k2b_fit_batch (device pointers)."""
    import ctypes as C

    from keypoints2body_b200 import _native as nat

    tgt, init = make_problem(weights, 300, seed=303)
    f = fitters("smpl", use_lbfgs=False)
    dev_out = f.fit_batch(init, tgt, None, seq_ind=0, num_iters=5, with_mesh=False)
    B = 300
    pose = torch.cat([init["global_orient"], init["body_pose"]], dim=1).contiguous()
    outs = dict(pose=torch.empty(B, 72), betas=torch.empty(B, 10), transl=torch.empty(B, 3), loss=torch.empty(B),
                joints=torch.empty(B, 22, 3), evals=torch.empty(B, dtype=torch.int32))
    a = nat.FitArgs(num_frames=B, num_obs=22, optimizer=nat.OPT_ADAM, num_iters=5, freeze_betas=0, conf_per_frame=0,
                    lr=1e-2, joint_loss_weight=600.0, pose_preserve_weight=5.0, targets=nat.ptr(tgt.contiguous()),
                    conf=None, init_pose=nat.ptr(pose), init_betas=nat.ptr(init["betas"]),
                    init_transl=nat.ptr(init["transl"].contiguous()), init_expr=None, preserve_pose=None,
                    frame_iters=None, frame_preserve=None, preserve_all=0, out_pose=nat.ptr(outs["pose"]),
                    out_betas=nat.ptr(outs["betas"]), out_transl=nat.ptr(outs["transl"]), out_expr=None,
                    out_loss=nat.ptr(outs["loss"]), out_joints=nat.ptr(outs["joints"]),
                    out_evals=nat.ptr(outs["evals"]), workspace=None, workspace_bytes=0)
    nat.check(f.native.lib.k2b_fit_batch_host(f.native.handle, C.byref(a), nat.current_stream()))
    assert torch.equal(outs["pose"][:, 3:], dev_out["params"]["body_pose"].cpu())
    assert torch.equal(outs["loss"], dev_out["loss"].cpu())
    assert torch.equal(outs["joints"], dev_out["fit_joints"].cpu())


def test_batch_separability_and_ragged_tiles(fitters, weights):
    """Frames are independent units: any frame fitted inside a large ragged batch (70 001 frames,
    546.9 tiles) is bit-identical to the same frame fitted alone, for both optimisers."""
    n = 70001
    tgt, init = make_problem(weights, 257, seed=404)
    rep = (n + 256) // 257
    tgt_big = tgt.repeat(rep, 1, 1)[:n].contiguous()
    init_big = {k: v.repeat(rep, 1)[:n].contiguous() for k, v in init.items()}
    for lbfgs in (False, True):
        f = fitters("smpl", use_lbfgs=lbfgs)
        big = f.fit_batch(init_big, tgt_big, None, seq_ind=2, num_iters=6, with_mesh=False)
        idx = torch.tensor([0, 1, 127, 128, 256, 257, 40000, n - 1])
        small = f.fit_batch({k: v[idx] for k, v in init_big.items()}, tgt_big[idx], None, seq_ind=2, num_iters=6,
                            with_mesh=False)
        for k in ("body_pose", "global_orient", "transl", "betas"):
            assert torch.equal(big["params"][k][idx.cuda()], small["params"][k]), (lbfgs, k)
        assert torch.equal(big["loss"][idx.cuda()], small["loss"])
        # periodic inputs -> periodic outputs
        assert torch.equal(big["params"]["body_pose"][:257], big["params"]["body_pose"][257:514])


def test_fit_reduces_loss_and_recovers_joints_full_size(fitters, weights):
    """Size-independent properties at a BASELINE-sized batch (65 536 frames): Adam lowers the loss on
    every frame and the fitted joints approach the noise-free targets."""
    n = 65536
    tgt, init = make_problem(weights, 512, seed=505, noise=0.0)
    tgt = tgt.repeat(n // 512, 1, 1).contiguous()
    init = {k: v.repeat(n // 512, 1).contiguous() for k, v in init.items()}
    f = fitters("smpl", use_lbfgs=False)
    before = f.evaluate_batch(init, tgt)["loss"]
    out = f.fit_batch(init, tgt, None, seq_ind=0, num_iters=30, with_mesh=False)
    after = f.evaluate_batch(out["params"], tgt)["loss"]
    assert bool((after < before).all())
    err0 = (f.evaluate_batch(init, tgt)["joints"] - tgt.cuda()).norm(dim=-1).mean()
    err1 = (out["fit_joints"] - tgt.cuda()).norm(dim=-1).mean()
    assert float(err1) < 0.5 * float(err0)


def test_edge_cases(fitters, weights):
    tgt, init = make_problem(weights, 3, seed=606)
    f = fitters("smpl", use_lbfgs=False)
    # zero iterations: parameters pass through
    out = f.fit_batch(init, tgt, None, seq_ind=0, num_iters=0, with_mesh=False)
    assert torch.equal(out["params"]["body_pose"].cpu(), init["body_pose"])
    # a joint with zero confidence contributes neither loss nor gradient
    conf = torch.ones(22)
    conf[20] = 0.0
    a = f.evaluate_batch(init, tgt, conf)
    tgt2 = tgt.clone()
    tgt2[:, 20] += 5.0
    b = f.evaluate_batch(init, tgt2, conf)
    assert torch.equal(a["loss"], b["loss"]) and torch.equal(a["grad_pose"], b["grad_pose"])
    # SMPL24 observations on a non-SMPL tree are refused, not silently mis-fitted
    with pytest.raises(NotImplementedError):
        fitters("smplh", "SMPL24").fit_batch(
            dict(init, left_hand_pose=torch.zeros(3, 45), right_hand_pose=torch.zeros(3, 45)),
            torch.zeros(3, 24, 3), None, num_iters=1, with_mesh=False)
    with pytest.raises(ValueError):
        f.fit_batch({k: v for k, v in init.items() if k != "transl"}, tgt, None)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("tag,iters,seq_ind", [("cam_given_adam", 15, 0), ("cam_given_adam_follow", 15, 2),
                                               ("cam_adam", 15, 0), ("cam_adam_follow", 15, 2)])
def test_camera_fitter_adam_vs_reference_goldens(goldens, weights, gmm, tag, iters, seq_ind, kernel):
    """Camera-space two-stage fitter, Adam: reference CameraSpaceFitter goldens.  Strict from a
    caller-supplied ``init_cam_t``; from the stage-0 estimate (zero translation gradient -> Adam's first
    step is rounding noise in the reference too) only the outcome is compared."""
    from keypoints2body_b200.core.fitters.camera_space import CameraSpaceFitter
    from keypoints2body_b200.models.smpl_data import SMPLData

    g = goldens
    f = CameraSpaceFitter(weights("smpl"), num_iters=iters, use_lbfgs=False, joints_category="AMASS",
                          model_type="smpl", gmm=gmm)
    f.warp_kernel_max_frames = 0 if kernel == "frame" else 1 << 20     # both stages on that kernel
    pose, tgt = T(g["cam_in_pose"]), T(g["cam_in_target"])
    strict = "given" in tag
    for b in range(3):
        init = SMPLData(betas=torch.zeros(1, 10), global_orient=pose[b:b + 1, :3], body_pose=pose[b:b + 1, 3:])
        r = f.fit_frame(init, tgt[b:b + 1], torch.ones(22), seq_ind=seq_ind, freeze_betas=True,
                        init_cam_t=T(g["cam_given_init"][b:b + 1]) if strict else None)
        ref_loss = float(g[tag + "_loss"][b])
        if strict:
            assert np.abs(cpu(r.params.pose) - g[tag + "_pose"][b:b + 1]).max() < 1e-4
            assert np.abs(cpu(r.params.transl) - g[tag + "_transl"][b:b + 1]).max() < 1e-4
            assert np.abs(cpu(r.params.betas) - g[tag + "_betas"][b:b + 1]).max() < 1e-4
            assert np.abs(cpu(r.joints) - g[tag + "_joints"][b:b + 1]).max() < 1e-4   # no camera translation in joints
            np.testing.assert_allclose(float(r.loss), ref_loss, rtol=1e-4)
        else:
            assert abs(float(r.loss) - ref_loss) < 0.25 * ref_loss
            assert np.abs(cpu(r.params.pose) - g[tag + "_pose"][b:b + 1]).max() < 0.1


@pytest.mark.parametrize("kernel", KERNELS)
def test_camera_fitter_lbfgs_statistics(goldens, weights, gmm, kernel):
    from keypoints2body_b200.core.fitters.camera_space import CameraSpaceFitter

    g = goldens
    f = CameraSpaceFitter(weights("smpl"), num_iters=20, use_lbfgs=True, joints_category="AMASS", model_type="smpl", gmm=gmm)
    f.warp_kernel_max_frames = 0 if kernel == "frame" else 1 << 20
    pose, tgt = T(g["cam_in_pose"]), T(g["cam_in_target"])
    out = f.fit_batch(dict(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=torch.zeros(3, 10)), tgt,
                      torch.ones(22), seq_ind=0, freeze_betas=True)
    # three chaotic fits pin no number; the machine is the one tests/test_gpu_lbfgs_parity.py pins.  Contract only:
    # the stage-2 loss is far below the stage-0 loss and within the spread of the reference's three fits.
    assert np.all(np.isfinite(cpu(out["loss"]))) and cpu(out["loss"]).max() <= 2.0 * g["cam_lbfgs_loss"].max()
    assert int(out["evals"].max()) <= 2 * (20 * 5 // 4 + 1)


def test_extreme_inputs_terminate_and_match_oracle(fitters, shims, oracle_prior, weights):
    """Rotations beyond pi, all-zero targets (the reference's smoke input) and non-finite observations:
    evaluation parity where the oracle is finite, and guaranteed termination (bounded budgets) where it is not."""
    g = torch.Generator().manual_seed(99)
    B = 4
    pose = torch.zeros(B, 72)
    pose[0, :3] = torch.tensor([3.1, 0.2, -0.1])              # near pi
    pose[1, 3:6] = torch.tensor([0.0, 4.6, 0.0])               # beyond pi
    pose[2] = 1.5 * torch.randn(72, generator=g)               # large everywhere
    params = dict(global_orient=pose[:, :3].contiguous(), body_pose=pose[:, 3:].contiguous(),
                  betas=2.0 * torch.randn(B, 10, generator=g), transl=torch.randn(B, 3, generator=g))
    tgt = torch.zeros(B, 22, 3)                                  # tests/test_integration_smoke.py:17-20 feeds zeros
    f = fitters("smpl", use_lbfgs=False)
    keep = params["body_pose"] * 0.5
    ours = f.evaluate_batch(params, tgt, torch.ones(22), preserve_pose=keep, preserve_on=True, pose_preserve_weight=5.0)
    ref_loss, ref_grads, _ = rp.evaluate(shims("smpl"), oracle_prior, params, keep, tgt, torch.ones(22),
                                         pose_preserve_weight=5.0)
    np.testing.assert_allclose(cpu(ours["loss"]), ref_loss.numpy().reshape(-1), rtol=2e-5)
    gref = ref_grads["body_pose"].numpy()
    assert np.abs(cpu(ours["grad_pose"])[:, 3:] - gref).max() <= 2e-4 * np.abs(gref).max()
    # non-finite observations: both optimisers return (NaN results are fine, hanging is not)
    bad = tgt.clone()
    bad[0, 3, 1] = float("nan")
    bad[1, 7, 0] = float("inf")
    for lb in (False, True):
        for kern in KERNELS:
            out = fitters("smpl", use_lbfgs=lb).fit_batch(params, bad, None, seq_ind=0, num_iters=10, with_mesh=False,
                                                          kernel=kern)
            ev = cpu(out["evals"])
            assert ev.max() <= 10 * 5 // 4 + 2 and ev.min() >= 1
            assert np.isfinite(cpu(out["loss"])[2:]).all()       # finite frames are unaffected by their neighbours


# ---- warp-per-sequence kernel (k2b_fit_chain): the reference's serial frame loop in one launch ----------
def _chain_init(g, shims, S):
    tgt = T(g["seq_in_target"])
    with torch.no_grad():
        root = shims("smpl")(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69),
                             betas=torch.zeros(1, 10)).joints[0, 0]
    init = dict(global_orient=torch.zeros(S, 3), body_pose=torch.zeros(S, 69), betas=torch.zeros(S, 10),
                transl=(tgt[0, 0] - root).expand(S, 3).contiguous())
    return init, tgt


@pytest.mark.parametrize("name,chain", [("seq_adam_chain", True), ("seq_adam_indep", False)])
def test_chain_kernel_vs_reference_sequence_goldens(goldens, fitters, shims, name, chain):
    """Five copies of the golden sequence, one warp each, per-frame confidences: every copy must reproduce the
    reference's optimize_params_sequence (api/sequence.py:214-281) result, mesh included."""
    g, S = goldens, 5
    init, tgt = _chain_init(g, shims, S)
    f = fitters("smpl", use_lbfgs=False)
    out = f.fit_chain(init, tgt[None].expand(S, -1, -1, -1), torch.ones(S, tgt.shape[0], 22), chain=chain)
    Tn = tgt.shape[0]
    p = out["params"]
    pose = np.concatenate([cpu(p["global_orient"]), cpu(p["body_pose"])], axis=1).reshape(S, Tn, 72)
    for s in range(S):
        assert np.abs(pose[s] - g[name + "_pose"]).max() < 1e-4
        assert np.abs(cpu(p["betas"]).reshape(S, Tn, 10)[s] - g[name + "_betas"]).max() < 1e-4
        assert np.abs(cpu(p["transl"]).reshape(S, Tn, 3)[s] - g[name + "_transl"]).max() < 1e-5
        assert np.abs(cpu(out["joints"]).reshape(S, Tn, -1, 3)[s] - g[name + "_joints"]).max() < 1e-4
        np.testing.assert_allclose(cpu(out["loss"]).reshape(S, Tn)[s], g[name + "_loss"], rtol=1e-4)
    assert (cpu(out["evals"]).reshape(S, Tn) == np.array([30] + [10] * (Tn - 1))).all()


def test_chain_kernel_many_sequences_vs_frame_kernel(fitters, weights):
    """3 000 two-frame chains (more warps than one wave holds; 16 warps per CTA) against the same chains run as
    two launches of the one-thread-per-frame kernel.  Both are within 1e-4 of the oracle, so within 2e-4 of each other."""
    S = 3000
    tgt, init = make_problem(weights, 2 * S, seed=303)
    tgt = tgt.reshape(S, 2, 22, 3)
    init = {k: v[:S].contiguous() for k, v in init.items()}
    f = fitters("smpl", use_lbfgs=False)
    out = f.fit_chain(init, tgt, None, with_mesh=False)
    a = f.fit_batch(init, tgt[:, 0], None, seq_ind=0, with_mesh=False, kernel="frame")
    b = f.fit_batch(a["params"], tgt[:, 1], None, seq_ind=1, with_mesh=False, kernel="frame")
    for k in ("global_orient", "body_pose", "betas", "transl"):
        got = cpu(out["params"][k]).reshape(S, 2, -1)
        assert np.abs(got[:, 0] - cpu(a["params"][k])).max() < 2e-4, k
        assert np.abs(got[:, 1] - cpu(b["params"][k])).max() < 2e-4, k
    np.testing.assert_allclose(cpu(out["loss"]).reshape(S, 2)[:, 1], cpu(b["loss"]), rtol=1e-3)


@pytest.mark.parametrize("lbfgs", [False, True])
def test_chain_windows_time_major_identical_to_one_launch(goldens, fitters, shims, lbfgs):
    """Cutting the time axis into windows (one launch each, continuing from the previous window's last frame, mesh
    of a finished window overlapping the next fit) and writing time-major must not change a single bit."""
    g, S = goldens, 4
    init, tgt = _chain_init(g, shims, S)
    tgt = tgt[None].expand(S, -1, -1, -1) + 0.01 * torch.arange(S).view(S, 1, 1, 1)
    f = fitters("smpl", use_lbfgs=lbfgs)
    one = f.fit_chain(init, tgt, None)
    win = f.fit_chain(init, tgt, None, time_major=True, chunks=3)
    Tn = tgt.shape[1]
    for k in ("global_orient", "body_pose", "betas", "transl"):
        a = cpu(one["params"][k]).reshape(S, Tn, -1)
        b = cpu(win["params"][k]).reshape(Tn, S, -1).transpose(1, 0, 2)
        assert np.array_equal(a, b), k
    for k in ("loss", "joints", "vertices", "fit_joints"):
        a = cpu(one[k]).reshape(S, Tn, -1)
        b = cpu(win[k]).reshape(Tn, S, -1).transpose(1, 0, 2)
        assert np.array_equal(a, b), k


def test_mesh_overlap_policy_measures_then_settles(goldens, fitters, shims):
    """fit_chain's share of mesh passes held to the free SMs is chosen by event timing: one candidate per call, then the
    fastest; whichever share a call runs with, the results are the same bits.  A pinned share is not tuned."""
    S = 4
    init, tgt = _chain_init(goldens, shims, S)
    tgt = tgt[None].expand(S, -1, -1, -1) + 0.01 * torch.arange(S).view(S, 1, 1, 1)
    f = fitters("smpl", use_lbfgs=False)
    f.__dict__.pop("_overlap_tuners", None)
    runs = [f.fit_chain(init, tgt, None, time_major=True, chunks=3) for _ in range(11)]
    (key, pol), = f.overlap_policy().items()
    assert key == (S, tgt.shape[1], 3, False)
    assert sorted(pol["ms"]) == [0.55, 0.65, 0.75, 0.85] and all(v > 0 for v in pol["ms"].values())
    assert pol["fraction"] == min(pol["ms"], key=pol["ms"].get)
    for r in runs[1:]:
        for k in ("loss", "joints", "vertices"):
            assert torch.equal(r[k], runs[0][k]), k
    f.fit_chain(init, tgt, None, time_major=True, chunks=3, mesh_capped_fraction=0.3)
    assert f.overlap_policy()[key]["ms"] == pol["ms"]


@pytest.mark.parametrize("lbfgs", [False, True])
def test_warp_kernel_edge_cases(fitters, weights, lbfgs):
    """Zero iterations pass the parameters through; a zero-confidence joint does not influence the fit; an explicit
    temporal anchor (preserve_pose) and per-frame confidences are honoured exactly like the per-thread kernel does."""
    tgt, init = make_problem(weights, 7, seed=707)
    f = fitters("smpl", use_lbfgs=lbfgs)
    out = f.fit_batch(init, tgt, None, seq_ind=0, num_iters=0, with_mesh=False, kernel="warp")
    assert torch.equal(out["params"]["body_pose"].cpu(), init["body_pose"])
    assert torch.equal(out["params"]["transl"].cpu(), init["transl"])
    conf = torch.ones(7, 22)
    conf[:, 20] = 0.0
    tgt2 = tgt.clone()
    tgt2[:, 20] += 5.0
    a = f.fit_batch(init, tgt, conf, seq_ind=2, num_iters=4, with_mesh=False, kernel="warp")
    b = f.fit_batch(init, tgt2, conf, seq_ind=2, num_iters=4, with_mesh=False, kernel="warp")
    assert torch.equal(a["params"]["body_pose"], b["params"]["body_pose"]) and torch.equal(a["loss"], b["loss"])
    if not lbfgs:
        keep = init["body_pose"] + 0.05
        w = f.fit_batch(init, tgt, conf, seq_ind=2, num_iters=6, with_mesh=False, kernel="warp", preserve_pose=keep)
        t = f.fit_batch(init, tgt, conf, seq_ind=2, num_iters=6, with_mesh=False, kernel="frame", preserve_pose=keep)
        assert (w["params"]["body_pose"] - t["params"]["body_pose"]).abs().max() < 1e-4
        assert (w["params"]["body_pose"] - a["params"]["body_pose"]).abs().max() > 1e-4     # the anchor matters
        np.testing.assert_allclose(cpu(w["loss"]), cpu(t["loss"]), rtol=1e-4)


@pytest.mark.parametrize("mt", ["smplh", "smplx"])
def test_chain_kernel_smplh_smplx_vs_oracle(fitters, shims, oracle_prior, weights, mt):
    """Serial chains on the 52- / 55-joint models (expression is optimised for SMPL-X, hands / jaw / eyes pass
    through): 3 sequences x 3 frames, Adam, against the oracle's frame-by-frame loop."""
    from keypoints2body_b200 import synthetic as syn

    S, Tn = 3, 3
    w = weights(mt)
    mo = syn.make_motion(Tn, seed=808, num_sequences=S)
    g = torch.Generator().manual_seed(809)
    tgt = syn.kinematic_joints(w, mo["pose"][:, :66], mo["betas"], mo["transl"], 22)
    tgt = (tgt + 0.005 * torch.randn(tgt.shape, generator=g)).reshape(S, Tn, 22, 3)
    pose0 = mo["pose"].reshape(S, Tn, 72)[:, 0] + 0.05 * torch.randn(S, 72, generator=g)
    init = dict(global_orient=pose0[:, :3].contiguous(), body_pose=pose0[:, 3:].contiguous(), betas=torch.zeros(S, 10),
                transl=mo["transl"].reshape(S, Tn, 3)[:, 0].contiguous(),
                left_hand_pose=0.1 * torch.randn(S, 45, generator=g), right_hand_pose=0.1 * torch.randn(S, 45, generator=g))
    if mt == "smplx":
        init.update(expression=0.2 * torch.randn(S, 10, generator=g), jaw_pose=0.05 * torch.randn(S, 3, generator=g),
                    leye_pose=torch.zeros(S, 3), reye_pose=torch.zeros(S, 3))
    f = fitters(mt, use_lbfgs=False)
    out = f.fit_chain(init, tgt, torch.ones(22))
    prev = {k: None for k in rp.PARAM_ORDER}
    prev.update(init)
    for t in range(Tn):
        r = rp.fit_frame(shims(mt), oracle_prior, prev, tgt[:, t], torch.ones(22), seq_ind=t, use_lbfgs=False)
        prev = r["params"]
        rows = torch.arange(S) * Tn + t
        assert (cpu(out["params"]["body_pose"][rows]) - prev["body_pose"].numpy()).__abs__().max() < 1e-4
        assert (cpu(out["params"]["transl"][rows]) - prev["transl"].numpy()).__abs__().max() < 1e-5
        assert (cpu(out["joints"][rows]) - r["joints"].numpy()).__abs__().max() < 1e-4
        assert (cpu(out["vertices"][rows]) - r["vertices"].numpy()).__abs__().max() < 1e-4
        if mt == "smplx":
            assert (cpu(out["params"]["expression"][rows]) - prev["expression"].numpy()).__abs__().max() < 1e-4
    assert torch.equal(out["params"]["left_hand_pose"].cpu(), init["left_hand_pose"].repeat_interleave(Tn, dim=0))
