import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def goldens():
    path = os.path.join(ROOT, "tests", "golden", "ref_goldens.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def generic_goldens():
    """Dict-block / explicit-index observations through the unmodified reference (make_goldens_generic.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_goldens_generic.npz")
    return dict(np.load(path))


@pytest.fixture(scope="session")
def weights():
    from keypoints2body_b200 import synthetic as syn

    cache = {}

    def get(model_type):
        if model_type not in cache:
            cache[model_type] = syn.make_body_model(model_type, seed=0)
        return cache[model_type]

    return get


@pytest.fixture(scope="session")
def shims(weights):
    from oracle.smplx_shim import BodyModelShim

    cache = {}

    def get(model_type):
        if model_type not in cache:
            cache[model_type] = BodyModelShim(weights(model_type))
        return cache[model_type]

    return get


@pytest.fixture(scope="session")
def gmm():
    from keypoints2body_b200 import synthetic as syn

    return syn.make_gmm(seed=0)


@pytest.fixture(scope="session")
def oracle_prior(gmm):
    from oracle.reference_port import GMMPrior

    return GMMPrior(gmm)
