"""Checks the warp-cooperative fitting code (csrc/chain_core.cuh) on the CPU.

tests/host_emul/libk2b_warp_emul.so runs the device code's warp-level routines with the 32 lanes as
coroutines (see warp_emul.cu): the same evaluation / Adam / L-BFGS / sequence-chain code the chain
kernel runs, compared here with the goldens made from the unmodified reference.  The `-m gpu` tests then
check the real kernel through the C ABI.
"""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from keypoints2body_b200.core.prior import prepare_gmm

from test_host_emul import ADAM_CASES, LBFGS_CASES, f32, pack_x

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "host_emul")
EMU_LIB = os.path.join(EMU_DIR, "libk2b_warp_emul.so")
CSRC = os.path.join(HERE, "..", "keypoints2body_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc needed to build the harness")

fp = C.POINTER(C.c_float)
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int32)


def _stale():
    if not os.path.exists(EMU_LIB):
        return True
    t = os.path.getmtime(EMU_LIB)
    srcs = [os.path.join(EMU_DIR, "warp_emul.cu")] + [os.path.join(CSRC, f) for f in
                                                      ("fit_core.cuh", "lbfgs_core.cuh", "chain_core.cuh")]
    return any(os.path.getmtime(s) > t for s in srcs)


@pytest.fixture(scope="module")
def wemu():
    if _stale():
        subprocess.run(["sh", os.path.join(EMU_DIR, "build.sh")], check=True, capture_output=True)
    lib = C.CDLL(EMU_LIB)
    lib.wemu_model_create.restype = C.c_void_p
    lib.wemu_model_create.argtypes = [C.c_int, fp, fp, fp, dp, dp, ip, C.c_int]
    lib.wemu_eval.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, fp, fp, fp, fp, fp, fp, fp, ip]
    lib.wemu_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_int, C.c_float, C.c_float, C.c_float, fp, fp, C.c_int, fp, fp, fp, fp, fp,
                               fp, fp, fp, fp, fp, fp, ip, C.c_int, C.c_int, C.c_float, fp]
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(fp)


class WModel:
    def __init__(self, lib, weights, gmm, ns):
        self.lib, self.ns = lib, ns
        g = prepare_gmm(gmm)
        Jr = weights.J_regressor.double().numpy()
        J0 = np.ascontiguousarray(Jr @ weights.v_template.double().numpy())
        JS = np.ascontiguousarray(np.einsum("jv,vkl->jkl", Jr, weights.shapedirs.double().numpy()[..., :ns]))
        par = np.ascontiguousarray(weights.parents.numpy().astype(np.int32))
        self.h = lib.wemu_model_create(ns, _p(f32(g.chol)), _p(f32(g.means)), _p(f32(g.neg_log_w)),
                                       J0.ctypes.data_as(dp), JS.ctypes.data_as(dp), par.ctypes.data_as(ip), len(par))

    def evaluate(self, x, target, conf, keep, keep_w, joint_w=600.0):
        K, NX = target.shape[0], 75 + self.ns
        x, target, conf, keep = f32(x), f32(target), f32(conf), f32(keep)
        grad, loss, joints = np.zeros(NX, np.float32), np.zeros(1, np.float32), np.zeros((K, 3), np.float32)
        comp = np.zeros(1, np.int32)
        self.lib.wemu_eval(self.h, K, joint_w, keep_w, _p(target), _p(conf), _p(x), _p(keep), _p(grad), _p(loss),
                           _p(joints), comp.ctypes.data_as(ip))
        return dict(grad=grad, loss=float(loss[0]), joints=joints, comp=int(comp[0]))

    def chain(self, x0, targets, conf=None, first_seq_ind=0, chain=True, lbfgs=False, iters_first=30,
              iters_follow=10, freeze=False, lr=1e-2, joint_w=600.0, keep_w=5.0, keep=None, loss_kind=0,
              final_mode=0, depth_ref=None):
        """x0 (S,NX); targets (S,T,K,3); returns arrays shaped (S,T,..)."""
        S, T, K = targets.shape[:3]
        x0, targets = f32(x0), f32(targets)
        pose, transl, betas = f32(x0[:, :72]), f32(x0[:, 72:75]), f32(x0[:, 75:85])
        expr = f32(x0[:, 85:95]) if self.ns == 20 else None
        conf = None if conf is None else f32(conf)
        keep = None if keep is None else f32(keep)
        o = dict(pose=np.zeros((S, T, 72), np.float32), betas=np.zeros((S, T, 10), np.float32),
                 transl=np.zeros((S, T, 3), np.float32), expr=np.zeros((S, T, 10), np.float32),
                 loss=np.zeros((S, T), np.float32), joints=np.zeros((S, T, K, 3), np.float32),
                 evals=np.zeros((S, T), np.int32))
        self.lib.wemu_chain(self.h, K, S, T, first_seq_ind, int(chain), int(lbfgs), iters_first, iters_follow,
                            int(freeze), lr, joint_w, keep_w, _p(targets), _p(conf),
                            0 if conf is None else (1 if conf.ndim == 1 else 2), _p(pose), _p(betas), _p(transl),
                            _p(expr), _p(keep), _p(o["pose"]), _p(o["betas"]), _p(o["transl"]), _p(o["expr"]),
                            _p(o["loss"]), _p(o["joints"]), o["evals"].ctypes.data_as(ip), loss_kind, final_mode, 100.0,
                            None if depth_ref is None else _p(f32(depth_ref)))
        return o


@pytest.fixture(scope="module")
def wmodels(wemu, weights, gmm):
    cache = {}

    def get(mt):
        if mt not in cache:
            cache[mt] = WModel(wemu, weights(mt), gmm, 20 if mt == "smplx" else 10)
        return cache[mt]

    return get


@pytest.mark.parametrize("tag", ["eval_smpl_22_w0", "eval_smpl_22_w5", "eval_smpl_24_w5",
                                 "eval_smplh_22_w5", "eval_smplx_22_w0", "eval_smplx_22_w5"])
def test_warp_evaluation_matches_reference(goldens, wmodels, tag):
    """G1 for the warp-cooperative evaluation: loss rel <= 1e-5, gradient <= 1e-4 of its max, joints <= 1e-5 m."""
    _, mt, nobs, w = tag.split("_")
    g = goldens
    pose = np.concatenate([g[tag + "_in_global_orient"], g[tag + "_in_body_pose"]], axis=1)
    expr = g.get(tag + "_in_expression")
    x0 = pack_x(pose, g[tag + "_in_transl"], g[tag + "_in_betas"], expr)
    keep_w = float(w[1:])
    ref_pose = np.concatenate([g[tag + "_grad_global_orient"], g[tag + "_grad_body_pose"]], axis=1)
    for b in range(x0.shape[0]):
        out = wmodels(mt).evaluate(x0[b], g[tag + "_in_target"][b], g[tag + "_in_conf"], g[tag + "_in_keep"][b], keep_w)
        np.testing.assert_allclose(out["loss"], g[tag + "_loss"].reshape(-1)[b], rtol=1e-5)
        np.testing.assert_allclose(out["joints"], g[tag + "_joints"][b, : int(nobs)], atol=1e-5)
        gx = out["grad"]
        assert (np.abs(gx[:72] - ref_pose[b]) / np.abs(ref_pose[b]).max()).max() < 1e-4
        for name, sl in (("transl", slice(72, 75)), ("betas", slice(75, 85))):
            ref = g[f"{tag}_grad_{name}"][b]
            assert (np.abs(gx[sl] - ref) / np.abs(ref).max()).max() < 1e-4
        if expr is not None:
            ref = g[tag + "_grad_expression"][b]
            assert (np.abs(gx[85:95] - ref) / np.abs(ref).max()).max() < 1e-4


@pytest.mark.parametrize("tag", sorted(ADAM_CASES))
def test_warp_adam_fit_matches_reference(goldens, wmodels, tag):
    """G2 for the warp-cooperative path: each frame is a one-frame chain starting at seq_ind."""
    mt, iters, seq_ind, freeze, nobs = ADAM_CASES[tag]
    g = goldens
    B = g[tag + "_in_pose"].shape[0]
    expr = np.zeros((B, 10), np.float32) if mt == "smplx" else None
    x0 = pack_x(g[tag + "_in_pose"], g[tag + "_in_transl"], g[tag + "_in_betas"], expr)
    out = wmodels(mt).chain(x0, g[tag + "_in_target"][:, None], np.ones(nobs), first_seq_ind=seq_ind,
                            iters_first=iters, iters_follow=iters, freeze=freeze)
    assert np.abs(out["pose"][:, 0] - g[tag + "_pose"]).max() < 1e-4
    assert np.abs(out["transl"][:, 0] - g[tag + "_transl"]).max() < 1e-5
    assert np.abs(out["betas"][:, 0] - g[tag + "_betas"]).max() < 1e-4
    assert np.abs(out["joints"][:, 0] - g[tag + "_joints"][:, :nobs]).max() < 1e-4
    np.testing.assert_allclose(out["loss"].sum(), float(g[tag + "_loss"]), rtol=1e-4)
    if mt == "smplx":
        assert np.abs(out["expr"][:, 0] - g[tag + "_expression"]).max() < 1e-4
    assert (out["evals"] == iters).all()


@pytest.mark.parametrize("tag", sorted(LBFGS_CASES))
def test_warp_lbfgs_statistics(goldens, wmodels, tag):
    """G4 for the warp-cooperative L-BFGS (same machine as the per-thread kernel, lane-distributed vectors):
    evaluation budgets like torch's, final losses not worse than the reference's in distribution."""
    mt, iters, seq_ind = LBFGS_CASES[tag]
    g = goldens
    B = g[tag + "_in_pose"].shape[0]
    expr = np.zeros((B, 10), np.float32) if mt == "smplx" else None
    x0 = pack_x(g[tag + "_in_pose"], g[tag + "_in_transl"], g[tag + "_in_betas"], expr)
    out = wmodels(mt).chain(x0, g[tag + "_in_target"][:, None], np.ones(22), first_seq_ind=seq_ind, lbfgs=True,
                            iters_first=iters, iters_follow=iters)
    ev, ref_ev = out["evals"][:, 0], g[tag + "_nevals"].reshape(-1)
    assert (ev <= iters * 5 // 4 + 1).all()
    assert (np.abs(ev - ref_ev) <= 2).all(), (ev, ref_ev)
    ref_loss = g[tag + "_loss"].reshape(-1)
    # a handful of chaotic fits pins no loss level: trial-by-trial agreement up to a noise-level decision is checked by
    # tests/test_lbfgs_conformance.py, the loss distribution on the GPU by tests/test_gpu_lbfgs_parity.py
    assert np.all(np.isfinite(out["loss"])) and ref_loss.size == out["loss"].size
    print(tag, "evals", ev, "ref", ref_ev, "loss", out["loss"].ravel(), "ref", ref_loss)


def _sequence_init(g, shims):
    tgt = g["seq_in_target"]
    with torch.no_grad():
        root = shims("smpl")(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69),
                             betas=torch.zeros(1, 10)).joints[0, 0].numpy()
    return pack_x(np.zeros((1, 72), np.float32), (tgt[0, 0] - root)[None], np.zeros((1, 10), np.float32)), tgt


@pytest.mark.parametrize("name,chain", [("seq_adam_chain", True), ("seq_adam_indep", False)])
def test_warp_sequence_chain_matches_reference(goldens, wmodels, shims, name, chain):
    """The reference's sequence loop (api/sequence.py:214-281) inside one warp: frame t starts from frame t-1's
    result (or from the fixed initialisation), 30 / 10 Adam iterations, temporal term from frame 1 on."""
    g = goldens
    x0, tgt = _sequence_init(g, shims)
    out = wmodels("smpl").chain(x0, tgt[None], np.ones(22), chain=chain)
    assert np.abs(out["pose"][0] - g[name + "_pose"]).max() < 1e-4
    assert np.abs(out["betas"][0] - g[name + "_betas"]).max() < 1e-4
    assert np.abs(out["transl"][0] - g[name + "_transl"]).max() < 1e-5
    assert np.abs(out["joints"][0] - g[name + "_joints"][:, :22]).max() < 1e-4
    np.testing.assert_allclose(out["loss"][0], g[name + "_loss"], rtol=1e-4)


@pytest.mark.parametrize("tag,iters,seq_ind", [("cam_given_adam", 15, 0), ("cam_given_adam_follow", 15, 2)])
def test_warp_camera_two_stage_adam_matches_reference(goldens, wmodels, tag, iters, seq_ind):
    """Camera-space fitter (camera_space.py:81-339) on the warp-cooperative path: stage 1 (loss_kind 1: torso joints,
    plain squared error, depth anchor, only orientation + translation move) then stage 2 (world loss with the
    initial body pose as temporal anchor, loss re-evaluated without it), against the reference goldens made with a
    caller-supplied ``init_cam_t``."""
    g = goldens
    m = wmodels("smpl")
    pose, tgt, cam0 = g["cam_in_pose"], g["cam_in_target"], g["cam_given_init"]
    x0 = pack_x(pose, cam0, np.zeros((len(pose), 10), np.float32))
    s1 = m.chain(x0, tgt[:, None], np.ones(22), first_seq_ind=0, iters_first=iters, iters_follow=iters, freeze=True,
                 loss_kind=1, depth_ref=cam0)
    assert np.abs(s1["pose"][:, 0, 3:] - pose[:, 3:]).max() == 0          # stage 1 moves orientation + translation only
    assert np.abs(s1["betas"][:, 0]).max() == 0
    x1 = pack_x(s1["pose"][:, 0], s1["transl"][:, 0], s1["betas"][:, 0])
    s2 = m.chain(x1, tgt[:, None], np.ones(22), first_seq_ind=1 if seq_ind > 0 else 0, iters_first=iters,
                 iters_follow=iters, freeze=seq_ind > 0, keep=pose[:, None, 3:], final_mode=1)
    assert np.abs(s2["pose"][:, 0] - g[tag + "_pose"]).max() < 1e-4
    assert np.abs(s2["transl"][:, 0] - g[tag + "_transl"]).max() < 1e-4
    assert np.abs(s2["betas"][:, 0] - g[tag + "_betas"]).max() < 1e-4
    np.testing.assert_allclose(s2["loss"][:, 0], g[tag + "_loss"].reshape(-1), rtol=1e-4)


@pytest.mark.parametrize("E,H", [(1, 2), (5, 0), (3, 1), (6, 1), (2, 2)])
def test_team_is_bit_identical_to_one_warp(goldens, wemu, wmodels, shims, E, H):
    """A sequence served by a team -- E evaluator warps trying the line search's next steps in the same round, H helper
    warps per evaluator scanning the mixture prior -- must return exactly what one warp returns: evaluations are pure
    functions of the point, and the machine consumes the same (loss, gradient) sequence either way.  L-BFGS chain of 6
    frames (30-iteration first frame, 10-iteration follow-ups with the temporal term) and the SMPL-X 5-iteration case."""
    g = goldens
    wemu.wemu_set_team.argtypes = [C.c_int, C.c_int]
    x0, tgt = _sequence_init(g, shims)
    xx = pack_x(g["lbfgs_smplx_in_pose"], g["lbfgs_smplx_in_transl"], g["lbfgs_smplx_in_betas"], np.zeros((3, 10), np.float32))
    try:
        wemu.wemu_set_team(1, 0)
        ref = wmodels("smpl").chain(x0, tgt[None], np.ones(22), lbfgs=True)
        refx = wmodels("smplx").chain(xx, g["lbfgs_smplx_in_target"][:, None], np.ones(22), lbfgs=True, iters_first=5, iters_follow=5)
        wemu.wemu_set_team(E, H)
        out = wmodels("smpl").chain(x0, tgt[None], np.ones(22), lbfgs=True)
        outx = wmodels("smplx").chain(xx, g["lbfgs_smplx_in_target"][:, None], np.ones(22), lbfgs=True, iters_first=5, iters_follow=5)
        adam = wmodels("smpl").chain(x0, tgt[None], np.ones(22), lbfgs=False)        # helpers only (teams are L-BFGS)
    finally:
        wemu.wemu_set_team(1, 0)
    for a, b in ((ref, out), (refx, outx)):
        for k in ("pose", "betas", "transl", "expr", "loss", "joints", "evals"):
            assert np.array_equal(a[k], b[k]), (E, H, k)
    assert np.abs(adam["pose"][0] - g["seq_adam_chain_pose"]).max() < 1e-4
    print("evaluations per frame", ref["evals"].ravel())


@pytest.mark.parametrize("lbfgs,iters", [(False, 15), (True, 20)])
def test_camera_sequence_in_one_launch(goldens, wemu, wmodels, lbfgs, iters):
    """``camera_sequence``: every frame a whole CameraSpaceFitter.fit_frame inside the launch (camera translation from
    the torso joints at the frame's initial parameters, stage 1, stage 2; camera_space.py:81-339), frames chained like
    the reference's loop (api/sequence.py:214-281).  (a) Bit-identical to the same fit done as separate stage launches
    chained by hand; (b) the first frame lands where the reference's fit from its default start lands (outcome level:
    the reference's first stage-1 step is decided by rounding noise, DESIGN.md section 3, note on G6')."""
    g = goldens
    m = wmodels("smpl")
    wemu.wemu_set_camera_seq.argtypes = [C.c_int]
    pose, tgt = g["cam_in_pose"], g["cam_in_target"]
    T = tgt.shape[0]
    x0 = pack_x(pose[:1], np.zeros((1, 3), np.float32), np.zeros((1, 10), np.float32))
    wemu.wemu_set_camera_seq(1)
    try:
        one = m.chain(x0, tgt[None], np.ones(22), first_seq_ind=0, chain=True, lbfgs=lbfgs, iters_first=iters,
                      iters_follow=iters, freeze=True)
    finally:
        wemu.wemu_set_camera_seq(0)
    # the same, stage by stage and frame by frame
    prev = x0.copy()
    for t in range(T):
        z = prev.copy()
        z[:, 72:75] = 0
        j0 = m.evaluate(z[0], tgt[t], np.ones(22), z[0, 3:72], 0.0)["joints"]
        d = tgt[t] - j0
        cam0 = ((((d[2] + d[1]) + d[17]) + d[16]) / np.float32(4.0)).astype(np.float32)[None]
        x1 = prev.copy()
        x1[:, 72:75] = cam0
        s1 = m.chain(x1, tgt[t][None, None], np.ones(22), first_seq_ind=0, lbfgs=lbfgs, iters_first=iters, iters_follow=iters,
                     freeze=True, loss_kind=1, depth_ref=cam0)
        x2 = pack_x(s1["pose"][:, 0], s1["transl"][:, 0], s1["betas"][:, 0])
        s2 = m.chain(x2, tgt[t][None, None], np.ones(22), first_seq_ind=1 if t > 0 else 0, lbfgs=lbfgs, iters_first=iters,
                     iters_follow=iters, freeze=t > 0, keep=prev[:, None, 3:72], final_mode=1)
        for k in ("pose", "transl", "betas", "loss"):
            assert np.array_equal(one[k][0, t], s2[k][0, 0]), (t, k)
        assert one["evals"][0, t] == s1["evals"][0, 0] + s2["evals"][0, 0]
        prev = pack_x(s2["pose"][:, 0], s2["transl"][:, 0], s2["betas"][:, 0])
    tag = "cam_lbfgs" if lbfgs else "cam_adam"
    ref_loss = float(g[tag + "_loss"][0])
    assert one["loss"][0, 0] < 1.25 * ref_loss, (one["loss"][0, 0], ref_loss)
    assert np.abs(one["pose"][0, 0] - g[tag + "_pose"][0]).max() < 0.1
