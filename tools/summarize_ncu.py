#!/usr/bin/env python
"""Short text summary of an `ncu --set full` report (raw page): time, DRAM traffic, occupancy,
instruction counts, top stall reasons.   python tools/summarize_ncu.py gpurun_out/x.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("== %s" % d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print("  %-70s %14s %s" % (k, d[k][1], d[k][0]))
        def mbytes(key):
            if key not in d:
                return 0.0
            unit, val = d[key]
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
            return float(val) * scale
        print("  %-70s %14.3f %s" % ("dram traffic (read+write)", mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum"), "Mbyte"))
        stalls = sorted(((float(v[1] or 0), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")), reverse=True)
        print("  top stalls (warps stalled per issue): " + ", ".join(
            "%s=%.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in stalls[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
