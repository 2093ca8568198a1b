#!/usr/bin/env python
"""scripts/ncu_summary.py <file.ncu-rep> [--stalls] -- prints the handful of ncu metrics DESIGN.md / profiles/ quote."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__inst_executed_op_shfl.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        if "--stalls" in sys.argv:
            st = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct") and r[i]]
            for v, h in sorted(st, reverse=True)[:10]:
                print(f"  {h:70s} {v:16.2f} %")


if __name__ == "__main__":
    main()
