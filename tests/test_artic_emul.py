"""The general articulated fit (csrc/artic_core.cuh, artic_kernel.cuh) on the CPU against the reference's goldens.

tests/host_emul/libk2b_artic_emul.so runs the kernel's __host__ __device__ per-frame routine frame by frame on the host
(artic_emul.cu), with the C ABI's own argument structures.  tests/golden/r2_generic.npz holds what the UNMODIFIED
reference returns for the same problems (make_goldens_r2.py generic): WorldSpaceFitter's GENERIC path with hand joints
and vertex-picked landmarks for SMPL-H / SMPL-X (core/fitters/world_space.py:198-201, core/joints/adapters.py:224-380)
and MANOFitter / FLAMEFitter (core/fitters/misc_models.py:18-359).  Adam is compared strictly (G2 bars), L-BFGS by its
evaluation budget and final loss.  The `-m gpu` twin is tests/test_gpu_articulated.py.
"""

import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from keypoints2body_b200 import _native as nat
from keypoints2body_b200 import synthetic as syn
from keypoints2body_b200.body_model import extract_weights
from keypoints2body_b200.core.fitters import articulated as art
from keypoints2body_b200.core.prior import prepare_gmm
from oracle.problems import articulated_problem

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "host_emul")
EMU_LIB = os.path.join(EMU_DIR, "libk2b_artic_emul.so")
CSRC = os.path.join(HERE, "..", "keypoints2body_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc needed to build the harness")
fp = C.POINTER(C.c_float)


def _stale():
    if not os.path.exists(EMU_LIB):
        return True
    t = os.path.getmtime(EMU_LIB)
    srcs = [os.path.join(EMU_DIR, "artic_emul.cu")] + [os.path.join(CSRC, f) for f in
                                                       ("artic_core.cuh", "artic_kernel.cuh", "fit_core.cuh", "lbfgs_core.cuh")]
    return any(os.path.getmtime(s) > t for s in srcs)


@pytest.fixture(scope="module")
def emu():
    if _stale():
        subprocess.run(["sh", os.path.join(EMU_DIR, "build.sh")], check=True, capture_output=True)
    lib = C.CDLL(EMU_LIB)
    lib.emu_artic_fit.argtypes = [C.POINTER(nat.ArticDesc), fp, fp, fp, C.POINTER(nat.ArticFitArgs)]
    lib.emu_artic_fit.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def G():
    return dict(np.load(os.path.join(HERE, "golden", "r2_generic.npz")))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostArtic:
    """Host twin of articulated.ArticulatedModel: same arrays, the emulation instead of the CUDA library."""

    def __init__(self, lib, mt, gmm):
        self.lib, self.mt = lib, mt
        w = syn.make_body_model(mt, seed=0, num_extra=syn.NUM_EXTRA_SMPLX_BLOCKS if mt == "smplx" else None)
        self.weights = extract_weights(w, mt)
        self.blocks, self.offset, self.n = art.block_offsets(mt)
        self.a = art.model_arrays(self.weights)
        self.with_prior = mt in ("smpl", "smplh", "smplx")
        if self.with_prior:
            g = prepare_gmm(gmm)
            P = np.zeros((8, 69, 72), np.float32)
            for m in range(8):
                L = g.chol[m].astype(np.float64)
                P[m, :, :69] = (L @ L.T).astype(np.float32)
            mu = np.zeros((8, 72), np.float32)
            mu[:, :69] = g.means
            self.P, self.mu, self.nlw = P, mu, np.ascontiguousarray(g.neg_log_w, dtype=np.float32)
        a = self.a
        nfp, nip = nat._fp, nat._ip
        self.desc = nat.ArticDesc(
            num_joints=self.weights.num_joints, num_shape=self.weights.num_shape, num_params=self.n,
            num_picked=a["pv_t"].shape[0], parents=nip(a["parents"]), J0=nfp(a["J0"]), JS=nfp(a["JS"]),
            pose_src=nip(a["pose_src"]), shape_src=nip(a["shape_src"]), transl_src=self.offset["transl"],
            pv_template=nfp(a["pv_t"]), pv_shapedirs=nfp(a["pv_S"]), pv_posedirs=nfp(a["pv_P"]),
            pv_skin_idx=nip(a["skin_idx"]), pv_skin_w=nfp(a["skin_w"]), reg_w=nfp(a["reg"]), keep_w=nfp(a["keep"]),
            body_off=self.offset["body_pose"] if self.with_prior else -1, prior_model=None)

    def pack(self, init, B):
        x = np.zeros((B, self.n), np.float32)
        frozen = np.zeros(self.n, np.uint8)
        for name, w in self.blocks:
            v = init.get(name)
            o = self.offset[name]
            if v is None or v.shape[-1] == 0:
                frozen[o:o + w] = 1
            else:
                x[:, o:o + w] = v.numpy()
        return x, frozen

    def run(self, mode, x0, frozen, tgt, idx, conf, seq_ind, iters):
        B, K = tgt.shape[0], tgt.shape[1]
        out = dict(x=np.zeros((B, self.n), np.float32), loss=np.zeros(B, np.float32), grad=np.zeros((B, self.n), np.float32),
                   points=np.zeros((B, K, 3), np.float32), evals=np.zeros(B, np.int32))
        hold = [np.ascontiguousarray(idx, dtype=np.int32), np.ascontiguousarray(tgt, dtype=np.float32),
                np.ascontiguousarray(conf, dtype=np.float32), np.ascontiguousarray(x0), np.ascontiguousarray(frozen)]
        args = nat.ArticFitArgs(
            num_frames=B, num_obs=K, mode=mode, num_iters=iters, conf_per_frame=0, lr=1e-2, joint_loss_weight=600.0,
            keep_scale=25.0 if seq_ind > 0 else 0.0, obs_idx=_ptr(hold[0]), targets=_ptr(hold[1]), conf=_ptr(hold[2]),
            init_x=_ptr(hold[3]), keep_x=None, frozen=_ptr(hold[4]), out_x=_ptr(out["x"]), out_loss=_ptr(out["loss"]),
            out_grad=_ptr(out["grad"]), out_points=_ptr(out["points"]), out_evals=_ptr(out["evals"]),
            out_gmm_component=None, workspace=None, workspace_bytes=0)
        gp = (self.P.ctypes.data_as(fp), self.mu.ctypes.data_as(fp), self.nlw.ctypes.data_as(fp)) if self.with_prior else (None,) * 3
        rc = self.lib.emu_artic_fit(C.byref(self.desc), *gp, C.byref(args))
        assert rc == 0, rc
        return out


@pytest.fixture(scope="module")
def hosts(emu, gmm):
    cache = {}

    def get(mt):
        if mt not in cache:
            cache[mt] = HostArtic(emu, mt, gmm)
        return cache[mt]

    return get


@pytest.mark.parametrize("mt", ["smplx", "smplh", "mano", "flame"])
@pytest.mark.parametrize("seq_ind", [0, 2])
def test_adam_fit_matches_reference(hosts, G, mt, seq_ind):
    """G2 bars on every optimised block, the returned loss and the observed model points."""
    h = hosts(mt)
    model, tgt, idx, init, B = articulated_problem(mt, 3, seed=700)
    x0, frozen = h.pack(init, B)
    out = h.run(nat.ARTIC_ADAM, x0, frozen, tgt.numpy(), idx.numpy(), np.ones(len(idx), np.float32), seq_ind, 10)
    tag = f"{mt}_adam_s{seq_ind}"
    worst = {}
    for name, w in h.blocks:
        key = f"{tag}_{name}"
        if key not in G:
            continue
        got = out["x"][:, h.offset[name]: h.offset[name] + w]
        worst[name] = float(np.abs(got - G[key]).max())
    print(tag, {k: f"{v:.1e}" for k, v in worst.items()})
    for name, v in worst.items():
        bar = 1e-5 if name == "transl" else 1e-4
        assert v < bar, (name, v)
    pts_ref = G[f"{tag}_joints"][:, idx.numpy()]
    assert np.abs(out["points"] - pts_ref).max() < 1e-4
    # per-frame losses; the reference returns the loss of the last iteration before its step, B = 1 per call
    assert np.allclose(out["loss"], G[f"{tag}_loss"], rtol=1e-4)


@pytest.mark.parametrize("mt", ["smplx", "smplh", "mano", "flame"])
@pytest.mark.parametrize("seq_ind", [0, 2])
def test_lbfgs_fit_budget_and_loss(hosts, G, mt, seq_ind):
    """torch's L-BFGS budget (max_iter 10 -> max_eval 12, the last line search may overshoot) and a final loss in the
    reference's range: the trajectories separate at the first accept test that is decided by float32 noise."""
    h = hosts(mt)
    model, tgt, idx, init, B = articulated_problem(mt, 3, seed=700)
    x0, frozen = h.pack(init, B)
    out = h.run(nat.ARTIC_LBFGS, x0, frozen, tgt.numpy(), idx.numpy(), np.ones(len(idx), np.float32), seq_ind, 10)
    tag = f"{mt}_lbfgs_s{seq_ind}"
    ref_loss, ref_evals = G[f"{tag}_loss"], G[f"{tag}_evals"]
    print(tag, "loss", out["loss"], "ref", ref_loss, "evals", out["evals"], "ref", ref_evals)
    assert np.all(out["evals"] >= 10) and np.all(out["evals"] <= 12 + 25)
    assert abs(int(out["evals"].sum()) - int(ref_evals.sum())) <= 2
    # three chaotic trajectories pin no level: a 1e-7 change of summation order moves single fits between the basins the
    # reference itself visits (e.g. 44 353 <-> 52 106 for the same frame at seq_ind 0 / 2); bound the damage instead
    assert np.all(out["loss"] < 1.5 * ref_loss) and np.median(out["loss"] / ref_loss) < 1.25


@pytest.mark.parametrize("mt", ["mano", "flame"])
def test_adam_chain_matches_reference_public_api(hosts, mt):
    """The reference's optimize_params_sequence for MANO / FLAME (zero initialisation, translation = first observed
    joint, 30 iterations for frame 0, then 10 with the temporal term, each frame starting from the previous result;
    api/sequence.py:192-281) replayed frame by frame on the emulation: G2 bars at every frame."""
    A = dict(np.load(os.path.join(HERE, "golden", "r2_generic_api.npz")))
    h = hosts(mt)
    model, tgt, idx, init, B = articulated_problem(mt, 4, seed=710)
    tgt = tgt.numpy()
    x = np.zeros((1, h.n), np.float32)
    x[0, h.offset["transl"]: h.offset["transl"] + 3] = tgt[0, 0]
    frozen = np.zeros(h.n, np.uint8)
    worst = 0.0
    for t in range(B):
        out = h.run(nat.ARTIC_ADAM, x, frozen, tgt[t:t + 1], idx.numpy(), np.ones(len(idx), np.float32), t, 30 if t == 0 else 10)
        x = out["x"].copy()
        for name, w in h.blocks:
            key = f"api_{mt}_{name}"
            if key in A:
                d = float(np.abs(x[0, h.offset[name]: h.offset[name] + w] - A[key][t]).max())
                worst = max(worst, d)
                assert d < 1e-4, (t, name, d)
        assert np.isclose(out["loss"][0], A[f"api_{mt}_loss"][t], rtol=1e-4), (t, out["loss"][0], A[f"api_{mt}_loss"][t])
    print(mt, "chain worst parameter difference", worst)
