"""The ctypes mirrors in keypoints2body_b200/_native.py must have exactly the layout of the C structs in
include/k2b_b200.h: same field names, offsets and sizes.  A tiny C program compiled with gcc prints
offsetof / sizeof for every field the Python side declares (so a field missing in the header fails to compile)."""

import ctypes as C
import os
import shutil
import subprocess

import pytest

from keypoints2body_b200 import _native as nat

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
STRUCTS = {"k2b_model_desc": nat.ModelDesc, "k2b_fit_args": nat.FitArgs, "k2b_chain_args": nat.ChainArgs,
           "k2b_eval_args": nat.EvalArgs, "k2b_mesh_args": nat.MeshArgs, "k2b_shape_args": nat.ShapeArgs,
           "k2b_replay_args": nat.ReplayArgs, "k2b_artic_desc": nat.ArticDesc, "k2b_artic_fit_args": nat.ArticFitArgs}

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc needed")


def test_ctypes_structs_match_the_header(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "k2b_b200.h"', "int main(void) {"]
    for cname, st in STRUCTS.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu %zu\\n", offsetof({cname}, {fname}), '
                         f'sizeof((({cname}*)0)->{fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = {}
    for ln in out:
        parts = ln.split()
        if len(parts) == 2:
            seen[parts[0]] = (int(parts[1]),)
        elif len(parts) == 3:
            seen[parts[0]] = (int(parts[1]), int(parts[2]))
    for cname, st in STRUCTS.items():
        assert seen[cname] == (C.sizeof(st),), (cname, seen[cname], C.sizeof(st))
        for fname, _ in st._fields_:
            f = getattr(st, fname)
            assert seen[f"{cname}.{fname}"] == (f.offset, f.size), (cname, fname, seen[f"{cname}.{fname}"], f.offset, f.size)


def test_header_has_no_fields_the_python_side_lacks():
    """Count the members of every struct in the header: a field added to the header only would shift nothing the first
    test sees if it sits at the end."""
    import re

    text = open(os.path.join(ROOT, "include", "k2b_b200.h")).read()
    for cname, st in STRUCTS.items():
        body = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", text, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        members = [m for m in body.split(";") if m.strip()]
        assert len(members) == len(st._fields_), (cname, len(members), len(st._fields_))
