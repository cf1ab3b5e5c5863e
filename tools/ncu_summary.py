#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py cite.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.per_cycle_active", "sm__cycles_active.avg",
    "smsp__inst_executed.sum", "sass__thread_inst_executed_true_per_opcode", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in KEEP:
            if k in d:
                print("  %-78s %s %s" % (k, d[k], u[k]))
        stalls = sorted(((float(v), k[len(STALL):-len("_per_issue_active.ratio")]) for k, v in d.items()
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        print("  warp stall reasons (warps per issue-active cycle):", ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
