#!/usr/bin/env python
"""One-screen summary of an .ncu-rep (raw page): duration, DRAM/L2/L1 traffic and hit rates,
occupancy, issue utilisation, top stall reasons.  Usage: ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__inst_executed_pipe_lsu.sum",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "")[:110])
        for w in WANT:
            if w in d:
                print("  %-62s %16s %s" % (w, d[w], units[hdr.index(w)]))
        stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")),
                        reverse=True)
        # fall back to pcsamp-based names
        if not stalls:
            stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                             if "warp_issue_stalled" in k and k.endswith(".pct") and v not in ("", "n/a")), reverse=True)
        for v, k in stalls[:8]:
            print("  stall %-56s %16.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__", "")[:56], v))


if __name__ == "__main__":
    main(sys.argv[1])
