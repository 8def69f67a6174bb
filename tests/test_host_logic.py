"""Host-side contracts, re-pointing the reference's own unit tests at this package
(/root/reference/tests/test_adapters.py, test_models.py, test_api_surface.py)."""

import numpy as np
import pytest
import torch

from keypoints2body_b200 import (FLAMEData, MANOData, SMPLData, SMPLHData, SMPLXData, optimize_params_frame,
                                 optimize_params_sequence)
from keypoints2body_b200.core.config import FrameOptimizeConfig, SequenceOptimizeConfig
from keypoints2body_b200.core.joints.adapters import (adapt_layout, adapt_layout_and_conf, normalize_frame_observations,
                                                      normalize_joints_frame, normalize_joints_sequence,
                                                      normalize_sequence_observations, resolve_adapter)
from keypoints2body_b200.core.prior import load_gmm, prepare_gmm


def test_api_exports_exist():
    assert callable(optimize_params_frame) and callable(optimize_params_sequence)


def test_normalize_joints_frame_shapes_and_conf():
    j3d, conf = normalize_joints_frame(np.zeros((22, 3), np.float32))
    assert tuple(j3d.shape) == (1, 22, 3) and tuple(conf.shape) == (22,)
    xyzc = np.zeros((22, 4), np.float32)
    xyzc[:, 3] = 0.7
    j3d, conf = normalize_joints_frame(xyzc)
    assert tuple(j3d.shape) == (1, 22, 3) and torch.allclose(conf, torch.full((22,), 0.7))
    with pytest.raises(ValueError):
        normalize_joints_frame(np.zeros((22, 5), np.float32))
    with pytest.raises(ValueError):
        normalize_joints_frame([[0, 0, 0]])


def test_normalize_joints_sequence():
    seq = np.zeros((5, 22, 4), np.float32)
    seq[:, :, 3] = 0.2
    xyz, conf = normalize_joints_sequence(seq)
    assert tuple(xyz.shape) == (5, 22, 3) and tuple(conf.shape) == (5, 22)


def test_layout_adapters():
    out, layout = adapt_layout(np.zeros((3, 25, 3), np.float32), "Manny25")
    assert tuple(out.shape) == (3, 22, 3) and layout == "AMASS"
    # OpenSim -> SMPL axes: (x, y, z) -> (z, y, -x)
    pts = np.zeros((1, 37, 3), np.float32)
    pts[0, 0] = (1.0, 2.0, 3.0)
    out, conf, layout = adapt_layout_and_conf(pts, np.ones((1, 37), np.float32), "SpineTrack37")
    assert np.allclose(out[0, 0], (3.0, 2.0, -1.0)) and conf.shape == (1, 22)
    assert resolve_adapter(24, None).name == "SMPL24"
    with pytest.raises(ValueError):
        resolve_adapter(23, None)
    with pytest.raises(ValueError):
        resolve_adapter(22, "SMPL24")
    with pytest.raises(ValueError):
        resolve_adapter(22, "nope")


def test_dict_observations():
    obs = {"body": np.zeros((22, 4), np.float32), "left_hand": np.zeros((21, 3), np.float32),
           "right_hand": np.zeros((21, 3), np.float32), "face": np.zeros((10, 3), np.float32)}
    j3d, conf, idx, layout = normalize_frame_observations(obs, layout=None, body_model="smplx")
    assert tuple(j3d.shape) == (1, 74, 3) and tuple(conf.shape) == (74,) and tuple(idx.shape) == (74,)
    assert layout == "GENERIC"
    seq = {k: np.zeros((4,) + v.shape, np.float32) for k, v in obs.items() if k != "face"}
    xyz, conf, idx, layout = normalize_sequence_observations(seq, layout=None, body_model="smplh")
    assert tuple(xyz.shape) == (4, 64, 3) and tuple(conf.shape) == (4, 64) and tuple(idx.shape) == (64,)
    with pytest.raises(ValueError):
        normalize_frame_observations({"left_hand": np.zeros((21, 3), np.float32)}, layout=None, body_model="smpl")
    with pytest.raises(ValueError):
        normalize_frame_observations({}, layout=None, body_model="smplx")


def test_param_containers():
    p = SMPLData(betas=torch.zeros(1, 10, requires_grad=True), global_orient=torch.zeros(1, 3),
                 body_pose=torch.zeros(1, 69), transl=torch.zeros(1, 3))
    p.validate()
    assert tuple(p.pose.shape) == (1, 72) and p.detach().betas.requires_grad is False
    n = SMPLData(betas=np.zeros((1, 10), np.float32), global_orient=np.zeros((1, 3), np.float32),
                 body_pose=np.zeros((1, 69), np.float32))
    assert n.pose.shape == (1, 72) and n.to(torch.device("cpu")).betas is n.betas
    assert issubclass(SMPLXData, SMPLHData) and issubclass(SMPLHData, SMPLData)
    mano = MANOData(betas=torch.zeros(1, 10), global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 0),
                    hand_pose=torch.zeros(1, 45))
    flame = FLAMEData(betas=torch.zeros(1, 10), global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 0),
                      expression=torch.zeros(1, 10))
    assert tuple(mano.pose.shape) == (1, 3) and tuple(flame.pose.shape) == (1, 3)
    with pytest.raises(ValueError):
        SMPLData(betas=None, global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69)).validate()


def test_config_defaults_match_reference():
    f = FrameOptimizeConfig()
    assert (f.estimator_type, f.input_type, f.coordinate_mode, f.use_lbfgs) == ("optimization", "joints3d", "world", True)
    assert (f.step_size, f.num_iters, f.num_iters_first, f.num_iters_followup) == (1e-2, 100, 30, 10)
    assert (f.joint_loss_weight, f.pose_preserve_weight, f.freeze_betas, f.shape_prior_weight) == (600.0, 5.0, False, 5.0)
    assert (f.pose_prior_num_gaussians, f.joints_category) == (8, "AMASS")
    s = SequenceOptimizeConfig()
    assert (s.num_shape_iters, s.num_shape_frames, s.use_shape_optimization, s.use_previous_frame_init) == (40, 50, True, True)
    assert s.fix_foot is False and s.limit_frames is None and s.schedule == "reference"


def test_prior_constants(tmp_path, gmm, goldens):
    from keypoints2body_b200 import synthetic as syn

    syn.write_assets(str(tmp_path), seed=0)
    c = prepare_gmm(load_gmm(str(tmp_path), 8))
    # the reference's own buffers (goldens) ...
    np.testing.assert_array_equal(c.precisions, goldens["prior_precisions"])
    np.testing.assert_array_equal(c.nll_weights, goldens["prior_nll_weights"].reshape(-1))
    # ... and the Cholesky form the kernel uses reproduces their symmetric part
    for m in range(8):
        P = c.precisions[m].astype(np.float64)
        L = c.chol[m].astype(np.float64)
        assert np.abs(L @ L.T - 0.5 * (P + P.T)).max() < 2e-5 * np.abs(P).max()
        assert np.allclose(np.triu(c.chol[m], 1), 0)
    with pytest.raises(FileNotFoundError):
        load_gmm(str(tmp_path), 6)


def test_torch_lbfgs_line_search_budget_is_max_eval_minus_evals():
    """The device machine caps a line search at ``max_ls = max_eval - evals`` (csrc/lbfgs_core.cuh, start_outer).
    That is what the torch build the reference runs on does: LBFGS.step passes ``max_ls=max_eval - current_evals``
    to _strong_wolfe (its signature default of 25 is never used by step).  Pinned here so a torch upgrade that
    changes it is noticed."""
    import inspect

    import torch.optim.lbfgs as L

    src = inspect.getsource(L.LBFGS.step)
    assert "max_ls=max_eval - current_evals" in src.replace("\n", " ").replace("  ", " ")


def test_parameter_width_validation():
    """Raw pointers cross the C ABI with fixed strides, so widths are checked on the host first (a genuine smplx
    SMPL-H / SMPL-X body_pose is 63 wide; this build follows the reference and expects 69)."""
    from keypoints2body_b200.core.fitters.world_space import _check_widths

    ok = dict(global_orient=(torch.zeros(4, 3), 3), body_pose=(torch.zeros(4, 69), 69), betas=(torch.zeros(1, 10), 10),
              transl=(torch.zeros(4, 3), 3), preserve_pose=(None, 69))
    _check_widths(4, **ok)
    for name, bad in (("body_pose", torch.zeros(4, 63)), ("betas", torch.zeros(4, 16)), ("transl", torch.zeros(4)),
                      ("global_orient", torch.zeros(3, 3)), ("preserve_pose", torch.zeros(4, 72))):
        with pytest.raises(ValueError, match=name):
            _check_widths(4, **dict(ok, **{name: (bad, ok[name][1])}))
