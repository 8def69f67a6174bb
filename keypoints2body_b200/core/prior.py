"""Max-mixture pose prior: asset loading and the constants the kernel consumes.

Restates ``MaxMixturePrior.__init__`` of the reference
(/root/reference/keypoints2body/core/prior.py:101-176) -- which constants are
derived from the pickled GMM and in which precision -- without its nn.Module /
``sys.exit`` plumbing.  The evaluation itself (prior.py:182-195) runs inside the
CUDA kernel (csrc/fit_core.cuh ``gmm_prior``).

Kernel form.  The reference evaluates ``0.5 d^T P_m d - log w_m`` with
``P_m = float32(inv(float32(cov_m)))`` and autograd returns ``0.5 (P + P^T) d``.
Only the symmetric part of ``P`` matters for both, so we factor
``sym(P_m) = L_m L_m^T`` (float64 Cholesky, stored as float32) and the kernel
computes ``||L^T d||^2`` and ``L (L^T d)`` -- half the multiply-adds of the dense
product and a sum of squares instead of a signed sum.
"""

from __future__ import annotations

import os
import pickle
from dataclasses import dataclass

import numpy as np

MAX_COMPONENTS = 8  # K2B_GMM_COMPONENTS in include/k2b_b200.h
POSE_DIM = 69


def load_gmm(prior_folder: str, num_gaussians: int = 8) -> dict:
    """Read ``gmm_{num_gaussians:02d}.pkl`` (dict with means / covars / weights).

    The reference calls ``sys.exit(-1)`` on a missing or unknown file
    (prior.py:117-131,145-146); here that is an exception.
    """
    path = os.path.join(prior_folder, f"gmm_{num_gaussians:02d}.pkl")
    if not os.path.exists(path):
        raise FileNotFoundError(f'The path to the mixture prior "{path}" does not exist')
    with open(path, "rb") as f:
        gmm = pickle.load(f, encoding="latin1")
    if not isinstance(gmm, dict):
        # the reference's sklearn branch indexes gmm[...] unconditionally afterwards
        # (prior.py:158-165), i.e. it only ever worked for dicts
        gmm = {"means": gmm.means_, "covars": gmm.covars_, "weights": gmm.weights_}
    for key in ("means", "covars", "weights"):
        if key not in gmm:
            raise ValueError(f"Unknown type for the prior: missing '{key}'")
    return gmm


@dataclass
class GMMConstants:
    means: np.ndarray       # (8, 69) float32
    chol: np.ndarray        # (8, 69, 69) float32 lower-triangular, sym(P) = L L^T
    neg_log_w: np.ndarray   # (8,) float32, -log(nll_weights); +inf for padding components
    precisions: np.ndarray  # (M, 69, 69) float32 exactly as the reference builds them
    nll_weights: np.ndarray # (M,) float32


def prepare_gmm(gmm: dict) -> GMMConstants:
    """Derive kernel constants with the reference's precision choices (prior.py:134-163)."""
    means = np.asarray(gmm["means"]).astype(np.float32)
    covs32 = np.asarray(gmm["covars"]).astype(np.float32)
    M, D = means.shape
    if D != POSE_DIM:
        raise ValueError(f"pose prior must be {POSE_DIM}-dimensional, got {D}")
    if M > MAX_COMPONENTS:
        raise ValueError(f"at most {MAX_COMPONENTS} mixture components are supported, got {M}")
    precisions = np.stack([np.linalg.inv(c) for c in covs32]).astype(np.float32)
    sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm["covars"]])
    const = (2 * np.pi) ** (69 / 2.0)
    nll_w = np.asarray(gmm["weights"] / (const * (sqrdets / sqrdets.min()))).astype(np.float32)

    chol = np.zeros((MAX_COMPONENTS, D, D), np.float32)
    mu = np.zeros((MAX_COMPONENTS, D), np.float32)
    nlw = np.full((MAX_COMPONENTS,), np.inf, np.float32)
    for m in range(M):
        p64 = precisions[m].astype(np.float64)
        chol[m] = np.linalg.cholesky(0.5 * (p64 + p64.T)).astype(np.float32)
        mu[m] = means[m]
    nlw[:M] = -np.log(nll_w)  # float32 log, like torch.log on the float32 buffer
    return GMMConstants(mu, chol, nlw, precisions, nll_w)
