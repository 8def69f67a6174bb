"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Recipe for ``oracle/_ref``: an UNMODIFIED copy of the reference's Python package, so that the GPU box -- where
/root/reference does not exist -- can time the reference itself (``bench.py --impl reference``,
``cpu_baseline.kind = "reference"``).

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference is present)

The reference is pure Python, so "building" it is copying ``/root/reference/keypoints2body/**/*.py`` where it lies
into ``oracle/_ref/keypoints2body/``; nothing is edited.  ``oracle/_ref/`` is git-ignored (no reference source enters
the history) and not gpurun-ignored (it travels with the snapshot).  The two imports the reference cannot satisfy
offline (``smplx``, ``h5py``) are stubbed at import time by ``oracle/ref_loader.py``, the body model is
``oracle/smplx_shim.py`` -- the same recipe that generated tests/golden/*.npz.
"""

from __future__ import annotations

import hashlib
import os
import shutil

SRC = "/root/reference/keypoints2body"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "keypoints2body")


def main() -> bool:
    if not os.path.isdir(SRC):
        print("make_ref: /root/reference is not present; keeping", DST if os.path.isdir(DST) else "(nothing)")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n, digest = 0, hashlib.sha256()
    for root, _dirs, files in os.walk(SRC):
        for name in sorted(files):
            if not name.endswith(".py"):
                continue
            src = os.path.join(root, name)
            dst = os.path.join(DST, os.path.relpath(src, SRC))
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            digest.update(open(src, "rb").read())
            n += 1
    with open(os.path.join(os.path.dirname(DST), "MANIFEST"), "w") as fh:
        fh.write(f"unmodified copy of {SRC}: {n} files, sha256 of their concatenation {digest.hexdigest()}\n")
    print(f"make_ref: copied {n} files to {DST}")
    return True


if __name__ == "__main__":
    main()
