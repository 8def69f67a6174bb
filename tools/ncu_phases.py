"""Stall samples of one kernel grouped by the CALL SITE inside a chosen function (authoring-container helper).

    python tools/ncu_phases.py report.ncu-rep build/csrc/chain_inst_10_22.o KERNEL_SUBSTRING FILE:FIRST-LAST CALLER_FILE [top]

e.g.  ... r02_chain_bench.ncu-rep build/csrc/chain_inst_10_22.o chain_kernelILi10ELi22ELb1ELb0E \
          chain_core.cuh:1083-1492 chain_kernel.cuh

tools/ncu_lines.py attributes a sample to the innermost source line, which for a kernel that is one big inlined function
says "a shuffle" or "a shared-memory load".  Here `nvdisasm -gi` supplies the inline chain of every instruction, and a
sample is attributed to the line of FILE:FIRST-LAST (the body of run_evaluator) whose caller frame is in CALLER_FILE
(the kernel): the phase of the round the warp was in.  Joined with ncu's SASS page by instruction offset (opcodes are
checked).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern, span, caller = sys.argv[1:6]
    top = int(sys.argv[6]) if len(sys.argv) > 6 else 30
    fname, rng = span.split(":")
    lo, hi = (int(v) for v in rng.split("-"))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", cubin], cwd=tmp, capture_output=True, text=True).stdout
    frames_at, ops, blk, fresh, inside = {}, {}, [], False, False
    for ln in dis.split("\n"):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if not fresh:
                blk, fresh = [], True
            blk.append((os.path.basename(m.group(1)), int(m.group(2)),
                        os.path.basename(m.group(3)) if m.group(3) else None, int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            fresh = False
            fr = [(b[0], b[1]) for b in blk]
            if blk and blk[-1][2]:
                fr.append((blk[-1][2], blk[-1][3]))
            off = int(m.group(1), 16)
            frames_at[off] = fr
            t = m.group(2).split()
            ops[off] = t[1] if t[0].startswith("@") else t[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ix["Address"]], 16)
    agg, ins = collections.Counter(), collections.Counter()
    why = collections.defaultdict(collections.Counter)
    sub = collections.defaultdict(collections.Counter)
    bad = 0
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        bad += ops.get(off) != op
        fr = frames_at.get(off, [])
        key, below = None, None
        for i, f in enumerate(fr):
            if f[0] == fname and lo <= f[1] <= hi and i + 1 < len(fr) and fr[i + 1][0] == caller:
                key, below = f[1], (fr[i - 1] if i > 0 else None)
        if key is None:
            key = "outside"
        s = int(r[ix["# Samples"]])
        agg[key] += s
        ins[key] += int(r[ix["Instructions Executed"]])
        for h in stalls:
            why[key][h[6:]] += int(r[ix[h]])
        if below:
            sub[key][below] += s
    tot = sum(agg.values())
    print(f"samples {tot}, instructions joined by offset: {len(data)} ({bad} opcode mismatches)")
    for k, v in agg.most_common(top):
        reasons = ", ".join(f"{n} {100 * c / max(v, 1):.0f}%" for n, c in why[k].most_common(3))
        print(f"{fname}:{k}  {100 * v / tot:5.1f}%  {ins[k]:12d} warp instr   [{reasons}]")
        for b, c in sub[k].most_common(3):
            if c > 0.15 * v:
                print(f"        in {b[0]}:{b[1]}  {100 * c / tot:5.1f}%")


if __name__ == "__main__":
    main()
