"""Per-source-line stall samples of one kernel from an ncu report (authoring-container helper).

    python tools/ncu_lines.py report.ncu-rep build/csrc/chain_inst_10_22.o [top]

ncu's CSV source page lists SASS instructions without line numbers; nvdisasm -g lists the same instructions in the
same order with `//## File ..., line N` markers.  The two are joined by position.
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
    rep, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout
    seq, cur = [], None
    for ln in dis.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", ln):
            seq.append(cur)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    if len(data) != len(seq):
        print("warning: instruction counts differ", len(data), len(seq))

    def f(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    agg, ins = collections.Counter(), collections.Counter()
    for r, s in zip(data, seq):
        agg[s] += f(r, "# Samples")
        ins[s] += f(r, "Instructions Executed")
    tot, tin = sum(agg.values()), sum(ins.values())
    print(f"samples {tot:.0f}  warp instructions {tin:.0f}")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    st = {h: sum(f(r, h) for r in data) for h in stalls}
    print("stalls:", ", ".join(f"{h[6:]} {100 * v / tot:.1f}%" for h, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
    for k, v in agg.most_common(top):
        print(f"{k[0]}:{k[1]:<5d} {100 * v / tot:5.1f}%  {ins[k]:12.0f} instr")


if __name__ == "__main__":
    main()
