#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) as a small CSV/markdown: per profiled launch the duration, DRAM bytes and
throughput, occupancy and its limiters, pipe utilisation, shared-memory bank conflicts and the top stall reasons.
usage: tools/ncu_summary.py <file.ncu-rep> [out.csv]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for k in KEYS:
            if k in hdr:
                d[k + " [" + units[hdr.index(k)] + "]"] = r[hdr.index(k)]
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h[len("smsp__pcsamp_warps_issue_stalled_"):]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        d["stalls (share of samples)"] = " ".join("%s=%.0f%%" % (n, 100 * v / tot) for v, n in stalls[:6])
        out.append(d)
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    keys = list(out[0].keys())
    w.writerow(["metric"] + [o["kernel"] for o in out])
    for k in keys[1:]:
        w.writerow([k] + [o.get(k, "") for o in out])


if __name__ == "__main__":
    main()
