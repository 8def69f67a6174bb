"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the UNMODIFIED reference package from ``/root/reference`` in the
authoring container.  Two of its imports are unavailable offline, so tiny stub
modules are planted in ``sys.modules`` first (recipe: SURVEY.md section 8c):

* ``smplx``  -- only ``smplx.create`` is referenced
  (/root/reference/keypoints2body/api/model_factory.py:34-40) and only when the
  caller passes ``model=None``; the stub raises there.
* ``h5py``   -- ``h5py.File(path, "r")`` used as a context manager exposing
  ``["pose"]`` and ``["shape"]`` (/root/reference/keypoints2body/core/engine.py:83-85);
  the stub reads the same keys from an ``.npz`` with the same stem.

The reference hard-codes CWD-relative asset paths (``./data/models/``), so
``reference_cwd`` chdir's into a scratch folder holding synthetic assets.
``/root/reference`` does not exist on the GPU box; nothing that runs there may
import this module.
"""

from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
VENDORED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
"""``oracle/_ref``: an unmodified copy of the reference's Python package made by ``oracle/make_ref.py`` (git-ignored,
travels to the GPU box with the snapshot) so that ``bench.py --impl reference`` times the reference itself there."""


def reference_root():
    """/root/reference in the authoring container, else the vendored copy, else None."""
    for root in (REFERENCE_ROOT, VENDORED_ROOT):
        if os.path.isdir(os.path.join(root, "keypoints2body")):
            return root
    return None


def available() -> bool:
    return reference_root() is not None


class _NpzAsH5:
    def __init__(self, path, mode="r"):
        stem = os.path.splitext(path)[0]
        self._d = np.load(stem + ".npz")

    def __enter__(self):
        return self._d

    def __exit__(self, *exc):
        return False


def _install_stubs():
    if "smplx" not in sys.modules:
        smplx = types.ModuleType("smplx")

        def create(*a, **k):
            raise RuntimeError("smplx is not installed; pass model= explicitly")

        smplx.create = create
        sys.modules["smplx"] = smplx
    if "h5py" not in sys.modules:
        h5py = types.ModuleType("h5py")
        h5py.File = _NpzAsH5
        sys.modules["h5py"] = h5py
    # torch_geometric is optional in the reference's ikgat package; if its import
    # fails the reference handles it lazily -- nothing to stub.


def load_reference():
    """Return the reference's top-level ``keypoints2body`` module (unmodified)."""
    root = reference_root()
    if root is None:
        raise RuntimeError(f"neither {REFERENCE_ROOT} nor {VENDORED_ROOT} holds the reference package")
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import keypoints2body  # noqa: WPS433  (the reference, not this repo's package)

    return keypoints2body


@contextlib.contextmanager
def reference_cwd(asset_root: str):
    """chdir into ``asset_root`` (must contain ``data/models/``) for the reference's relative paths."""
    old = os.getcwd()
    os.chdir(asset_root)
    try:
        yield
    finally:
        os.chdir(old)
