"""On-disk formats: our io.motion against files and outputs produced by the
reference's own io/motion.py (tests/golden/make_goldens.py section G).  Host-only code, runs on CPU."""
import io
import os
import warnings
import zipfile

import numpy as np
import pytest
import torch

from keypoints2body_b200.io import load_motion_data, write_smplx_zip, write_smplx_zip_from_smpl_data
from keypoints2body_b200.models.smpl_data import SMPLData

IO_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")


@pytest.mark.parametrize("name,layout", [("seq22.npy", None), ("seq22.csv", "AMASS"), ("seq25.npz", None)])
def test_load_motion_data_matches_reference(goldens, name, layout):
    key = "io_" + name.replace(".", "_")
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        joints, lay, k = load_motion_data(os.path.join(IO_DIR, name), layout)
    ref_layout, ref_k = goldens[key + "_layout"]
    assert lay == ref_layout and k == int(ref_k)
    np.testing.assert_array_equal(np.asarray(joints, np.float64), goldens[key + "_joints"])
    # a changed joint count is reported (motion.py:52-56)
    assert bool(caught) == (name == "seq25.npz")


def test_load_motion_data_errors(tmp_path):
    np.savez(tmp_path / "bad.npz", points=np.zeros((2, 22, 3)))
    with pytest.raises(ValueError):
        load_motion_data(tmp_path / "bad.npz")
    (tmp_path / "x.txt").write_text("1 2 3")
    with pytest.raises(ValueError):
        load_motion_data(tmp_path / "x.txt")


def _members(path):
    out = {}
    with zipfile.ZipFile(path) as zf:
        for n in zf.namelist():
            with np.load(io.BytesIO(zf.read(n))) as d:
                out[n] = {k: d[k] for k in d.files}
    return out


def test_write_smplx_zip_matches_reference(goldens, tmp_path):
    rng = np.random.default_rng(11)
    ref = _members(os.path.join(IO_DIR, "ref_params.zip"))
    # same inputs the golden script fed the reference: poses = gt_pose[:3] (float64), betas 1-D, transl (3,3)
    poses = goldens["mpjae_in_gt"][:3].astype(np.float64)
    transl = np.stack([ref[f"frame_{t:06d}/person_01.npz"]["transl"] for t in range(3)]).astype(np.float64)
    p = write_smplx_zip(tmp_path, poses, np.linspace(-1, 1, 10), transl, zip_name="ours.zip", person_idx=1)
    ours = _members(p)
    assert list(ours) == list(ref)
    for name, arrays in ref.items():
        assert set(ours[name]) == set(arrays)
        for k, v in arrays.items():
            assert ours[name][k].dtype == v.dtype and ours[name][k].shape == v.shape
            np.testing.assert_array_equal(ours[name][k], v)
    del rng


def test_write_smplx_zip_validation_and_dataclass(tmp_path):
    with pytest.raises(ValueError):
        write_smplx_zip(tmp_path, np.zeros((2, 66)), np.zeros(10), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        write_smplx_zip(tmp_path, np.zeros((2, 72)), np.zeros((3, 10)), np.zeros((2, 3)))
    with pytest.raises(ValueError):
        write_smplx_zip(tmp_path, np.zeros((2, 72)), np.zeros(10), np.zeros((3, 3)))
    data = SMPLData(betas=torch.zeros(2, 10), global_orient=torch.ones(2, 3), body_pose=torch.zeros(2, 69),
                    transl=torch.full((2, 3), 0.5))
    p = write_smplx_zip_from_smpl_data(tmp_path, data)
    m = _members(p)
    assert list(m) == ["frame_000000/person_00.npz", "frame_000001/person_00.npz"]
    assert m["frame_000001/person_00.npz"]["body_pose"].shape == (63,)
    np.testing.assert_array_equal(m["frame_000000/person_00.npz"]["global_orient"], np.ones(3, np.float32))
    with pytest.raises(ValueError):
        write_smplx_zip_from_smpl_data(tmp_path, SMPLData(betas=torch.zeros(1, 10), global_orient=torch.zeros(1, 3),
                                                          body_pose=torch.zeros(1, 69)))


def test_oracle_mpjae_matches_reference(goldens):
    from oracle import reference_port as rp

    pred, gt = goldens["mpjae_in_pred"], goldens["mpjae_in_gt"]
    ang = rp.angular_error_deg(pred.reshape(15, 22, 3), gt[:15, :66].reshape(15, 22, 3))
    np.testing.assert_allclose(ang, goldens["mpjae_angles"], atol=1e-4)
    mean, total, count = rp.evaluate_pose_pair(pred, gt)
    ref_mean, ref_total, ref_count = goldens["mpjae_summary"]
    assert count == int(ref_count) == 15 * 22
    assert abs(total - ref_total) < 1e-2 and abs(mean - ref_mean) < 1e-4
