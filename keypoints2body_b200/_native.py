"""ctypes binding of ``libk2b_b200.so`` (C ABI: include/k2b_b200.h).

There is no CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.  Build with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C keypoints2body_b200/csrc``.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .body_model import BodyModelWeights
from .core.prior import GMMConstants

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("K2B_LIB", os.path.join(_HERE, "libk2b_b200.so"))

OPT_ADAM, OPT_LBFGS = 0, 1
FREEZE_BETAS, FREEZE_EXPR = 1, 2      # k2b_fit_args.freeze_betas bits

_c_float_p = C.POINTER(C.c_float)
_c_int_p = C.POINTER(C.c_int32)
_c_u8_p = C.POINTER(C.c_uint8)


class ModelDesc(C.Structure):
    _fields_ = [
        ("num_joints", C.c_int32), ("num_vertices", C.c_int32), ("num_shape", C.c_int32),
        ("num_extra", C.c_int32), ("parents", _c_int_p), ("v_template", _c_float_p),
        ("shapedirs", _c_float_p), ("posedirs", _c_float_p), ("J_regressor", _c_float_p),
        ("lbs_weights", _c_float_p), ("extra_vertex_ids", _c_int_p), ("gmm_means", _c_float_p),
        ("gmm_chol", _c_float_p), ("gmm_neg_log_w", _c_float_p),
    ]


class FitArgs(C.Structure):
    _fields_ = [
        ("num_frames", C.c_int64), ("num_obs", C.c_int32), ("optimizer", C.c_int32),
        ("num_iters", C.c_int32), ("freeze_betas", C.c_int32), ("conf_per_frame", C.c_int32),
        ("lr", C.c_float), ("joint_loss_weight", C.c_float), ("pose_preserve_weight", C.c_float),
        ("targets", C.c_void_p), ("conf", C.c_void_p), ("init_pose", C.c_void_p),
        ("init_betas", C.c_void_p), ("init_transl", C.c_void_p), ("init_expr", C.c_void_p),
        ("preserve_pose", C.c_void_p), ("frame_iters", C.c_void_p), ("frame_preserve", C.c_void_p),
        ("preserve_all", C.c_int32),
        ("out_pose", C.c_void_p), ("out_betas", C.c_void_p), ("out_transl", C.c_void_p),
        ("out_expr", C.c_void_p), ("out_loss", C.c_void_p), ("out_joints", C.c_void_p),
        ("out_evals", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("loss_kind", C.c_int32), ("final_loss_mode", C.c_int32), ("depth_weight", C.c_float),
        ("depth_ref", C.c_void_p),
    ]


class ChainArgs(C.Structure):
    _fields_ = [
        ("num_sequences", C.c_int64), ("frames_per_sequence", C.c_int32), ("num_obs", C.c_int32),
        ("optimizer", C.c_int32), ("num_iters_first", C.c_int32), ("num_iters_followup", C.c_int32),
        ("first_seq_ind", C.c_int64), ("chain_init", C.c_int32), ("freeze_betas", C.c_int32),
        ("conf_mode", C.c_int32), ("out_time_major", C.c_int32), ("in_sequence_stride", C.c_int64),
        ("lr", C.c_float), ("joint_loss_weight", C.c_float), ("pose_preserve_weight", C.c_float),
        ("targets", C.c_void_p), ("conf", C.c_void_p), ("init_pose", C.c_void_p), ("init_betas", C.c_void_p),
        ("init_transl", C.c_void_p), ("init_expr", C.c_void_p), ("preserve_pose", C.c_void_p),
        ("seq_first_ind", C.c_void_p),
        ("out_pose", C.c_void_p), ("out_betas", C.c_void_p), ("out_transl", C.c_void_p), ("out_expr", C.c_void_p),
        ("out_loss", C.c_void_p), ("out_joints", C.c_void_p), ("out_evals", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("loss_kind", C.c_int32), ("final_loss_mode", C.c_int32), ("depth_weight", C.c_float),
        ("depth_ref", C.c_void_p), ("camera_sequence", C.c_int32),
    ]


class EvalArgs(C.Structure):
    _fields_ = [
        ("num_frames", C.c_int64), ("num_obs", C.c_int32), ("conf_per_frame", C.c_int32),
        ("preserve_all", C.c_int32), ("joint_loss_weight", C.c_float),
        ("pose_preserve_weight", C.c_float),
        ("targets", C.c_void_p), ("conf", C.c_void_p), ("pose", C.c_void_p), ("betas", C.c_void_p),
        ("transl", C.c_void_p), ("expr", C.c_void_p), ("preserve_pose", C.c_void_p),
        ("out_loss", C.c_void_p), ("out_grad_pose", C.c_void_p), ("out_grad_betas", C.c_void_p),
        ("out_grad_transl", C.c_void_p), ("out_grad_expr", C.c_void_p), ("out_joints", C.c_void_p),
        ("out_gmm_component", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("warp_evaluator", C.c_int32),
    ]


class ReplayArgs(C.Structure):
    _fields_ = [
        ("num_searches", C.c_int32), ("max_resp", C.c_int32), ("warp_policy", C.c_int32),
        ("t0", C.c_void_p), ("f0", C.c_void_p), ("gtd0", C.c_void_p), ("d_norm", C.c_void_p), ("max_ls", C.c_void_p),
        ("t_is_f32", C.c_void_p), ("n_resp", C.c_void_p), ("resp_f", C.c_void_p), ("resp_gtd", C.c_void_p),
        ("out_t", C.c_void_p), ("out_final", C.c_void_p), ("out_k", C.c_void_p),
    ]


class ArticDesc(C.Structure):
    _fields_ = [
        ("num_joints", C.c_int32), ("num_shape", C.c_int32), ("num_params", C.c_int32), ("num_picked", C.c_int32),
        ("parents", _c_int_p), ("J0", _c_float_p), ("JS", _c_float_p), ("pose_src", _c_int_p), ("shape_src", _c_int_p),
        ("transl_src", C.c_int32),
        ("pv_template", _c_float_p), ("pv_shapedirs", _c_float_p), ("pv_posedirs", _c_float_p),
        ("pv_skin_idx", _c_int_p), ("pv_skin_w", _c_float_p), ("reg_w", _c_float_p), ("keep_w", _c_float_p),
        ("body_off", C.c_int32), ("prior_model", C.c_void_p),
    ]


class ArticFitArgs(C.Structure):
    _fields_ = [
        ("num_frames", C.c_int64), ("num_obs", C.c_int32), ("mode", C.c_int32), ("num_iters", C.c_int32),
        ("conf_per_frame", C.c_int32), ("lr", C.c_float), ("joint_loss_weight", C.c_float), ("keep_scale", C.c_float),
        ("obs_idx", C.c_void_p), ("targets", C.c_void_p), ("conf", C.c_void_p), ("init_x", C.c_void_p),
        ("keep_x", C.c_void_p), ("frozen", C.c_void_p), ("out_x", C.c_void_p), ("out_loss", C.c_void_p),
        ("out_grad", C.c_void_p), ("out_points", C.c_void_p), ("out_evals", C.c_void_p),
        ("out_gmm_component", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


ARTIC_EVAL, ARTIC_ADAM, ARTIC_LBFGS = 0, 1, 2


class MeshArgs(C.Structure):
    _fields_ = [
        ("num_frames", C.c_int64), ("full_pose", C.c_void_p), ("shape", C.c_void_p),
        ("transl", C.c_void_p), ("out_vertices", C.c_void_p), ("out_joints", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("max_ctas", C.c_int32),
    ]


class ShapeArgs(C.Structure):
    _fields_ = [
        ("num_sequences", C.c_int32), ("frames_per_sequence", C.c_int32), ("sequence_stride", C.c_int64),
        ("num_obs", C.c_int32), ("pose_per_frame", C.c_int32), ("conf_per_sequence", C.c_int32),
        ("num_iters", C.c_int32), ("lr", C.c_float), ("shape_prior_weight", C.c_float),
        ("targets", C.c_void_p), ("poses", C.c_void_p), ("conf", C.c_void_p), ("init_betas", C.c_void_p),
        ("out_betas", C.c_void_p), ("out_loss", C.c_void_p), ("out_evals", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


EXPORTS = (
    "k2b_model_create", "k2b_model_destroy", "k2b_fit_workspace_bytes", "k2b_fit_batch",
    "k2b_fit_batch_host", "k2b_chain_workspace_bytes", "k2b_chain_geometry", "k2b_fit_chain", "k2b_evaluate_batch", "k2b_linesearch_replay", "k2b_mesh_workspace_bytes", "k2b_mesh_batch",
    "k2b_shape_workspace_bytes", "k2b_shape_pass", "k2b_artic_create", "k2b_artic_destroy", "k2b_artic_workspace_bytes",
    "k2b_artic_fit", "k2b_mpjae", "k2b_fma_peak", "k2b_launch_count", "k2b_last_error", "k2b_version",
)

_lib = None


def load_library():
    """dlopen the CUDA library; raise (never fall back) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run __graft_entry__.build()).  keypoints2body_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.k2b_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
    lib.k2b_model_create.restype = C.c_int
    lib.k2b_model_destroy.argtypes = [C.c_void_p]
    lib.k2b_model_destroy.restype = None
    lib.k2b_fit_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    lib.k2b_fit_workspace_bytes.restype = C.c_size_t
    lib.k2b_chain_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    lib.k2b_chain_workspace_bytes.restype = C.c_size_t
    lib.k2b_chain_geometry.argtypes = [C.c_void_p, C.c_int64, _c_int_p, _c_int_p]
    lib.k2b_chain_geometry.restype = C.c_int
    for name, st in (("k2b_fit_batch", FitArgs), ("k2b_fit_batch_host", FitArgs), ("k2b_fit_chain", ChainArgs),
                     ("k2b_evaluate_batch", EvalArgs), ("k2b_mesh_batch", MeshArgs),
                     ("k2b_shape_pass", ShapeArgs)):
        fn = getattr(lib, name)
        fn.argtypes = [C.c_void_p, C.POINTER(st), C.c_void_p]
        fn.restype = C.c_int
    lib.k2b_artic_create.argtypes = [C.POINTER(ArticDesc), C.POINTER(C.c_void_p)]
    lib.k2b_artic_create.restype = C.c_int
    lib.k2b_artic_destroy.argtypes = [C.c_void_p]
    lib.k2b_artic_destroy.restype = None
    lib.k2b_artic_workspace_bytes.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32]
    lib.k2b_artic_workspace_bytes.restype = C.c_size_t
    lib.k2b_artic_fit.argtypes = [C.c_void_p, C.POINTER(ArticFitArgs), C.c_void_p]
    lib.k2b_artic_fit.restype = C.c_int
    lib.k2b_linesearch_replay.argtypes = [C.POINTER(ReplayArgs), C.c_void_p]
    lib.k2b_linesearch_replay.restype = C.c_int
    lib.k2b_shape_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
    lib.k2b_shape_workspace_bytes.restype = C.c_size_t
    lib.k2b_mesh_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.k2b_mesh_workspace_bytes.restype = C.c_size_t
    lib.k2b_mpjae.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.k2b_mpjae.restype = C.c_int
    lib.k2b_fma_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p]
    lib.k2b_fma_peak.restype = C.c_int
    lib.k2b_launch_count.restype = C.c_int64
    lib.k2b_last_error.restype = C.c_char_p
    lib.k2b_version.restype = C.c_char_p
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load_library().k2b_last_error().decode()
        exc = {-1: ValueError, -3: MemoryError, -4: NotImplementedError}.get(rc, RuntimeError)
        raise exc(f"k2b_b200 error {rc}: {msg}")


def _fp(a: np.ndarray):
    return a.ctypes.data_as(_c_float_p)


def _ip(a: np.ndarray):
    return a.ctypes.data_as(_c_int_p)


def ptr(t):
    """Device / host address of a contiguous tensor (or None)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return t.data_ptr()


def current_stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class NativeModel:
    """Owns one ``k2b_model`` on the current CUDA device."""

    def __init__(self, weights: BodyModelWeights, gmm: GMMConstants, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("keypoints2body_b200 needs a CUDA device (no CPU fallback)")
        self.lib = load_library()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.weights = weights
        self.num_shape = weights.num_shape
        self.num_joints = weights.num_joints
        self.num_vertices = weights.num_vertices
        self.num_extra = weights.num_extra
        keep = [weights.parents, weights.v_template, weights.shapedirs, weights.posedirs,
                weights.J_regressor, weights.lbs_weights, weights.extra_vertex_ids]
        if gmm is not None:       # None: a mesh-only model (MANO, FLAME -- fitted through k2b_artic_fit)
            keep += [np.ascontiguousarray(gmm.means), np.ascontiguousarray(gmm.chol), np.ascontiguousarray(gmm.neg_log_w)]
        desc = ModelDesc(
            weights.num_joints, weights.num_vertices, weights.num_shape, weights.num_extra,
            _ip(keep[0]), _fp(keep[1]), _fp(keep[2]), _fp(keep[3]), _fp(keep[4]), _fp(keep[5]),
            _ip(keep[6]), *( [_fp(keep[7]), _fp(keep[8]), _fp(keep[9])] if gmm is not None else [None, None, None]),
        )
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.k2b_model_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.k2b_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def workspace(self, key: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws
