"""Selected raw metrics of every kernel in an ncu report (authoring-container helper).

    python tools/ncu_raw.py report.ncu-rep [substring ...]     # extra substrings select more metric names
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def main():
    rep, extra = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## launch {n}:", d.get("Kernel Name", "")[:90])
        for k in KEYS:
            if k in d:
                print(f"{k} = {d[k]} {u[k]}")
        for k in hdr:
            if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued") and float(d[k] or 0) > 0:
                print(f"stall samples {k.replace('smsp__pcsamp_warps_issue_stalled_', '')} = {d[k]}")
            elif any(e in k for e in extra) and k not in KEYS:
                print(f"{k} = {d[k]} {u[k]}")


if __name__ == "__main__":
    main()
