#!/usr/bin/env python
"""Summarise an `ncu --set full` report (run here, no GPU needed):
    python scripts/ncu_summary.py gpurun_out/<tag>/prof.ncu-rep > profiles/<tag>_ncu_summary.md
Prints, per profiled launch, the metrics DESIGN.md and bench.py's roofline refer to."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_static", "static smem/CTA"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/CTA"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__inst_executed_op_branch.sum", "branch warp instructions"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "LDS warp instructions"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local-memory loads (warp inst)"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local-memory stores (warp inst)"),
    ("sass__inst_executed_register_spilling", "spill instructions executed"),
    ("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", "FADD thread-inst / cycle (all SMs)"),
    ("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", "FMUL thread-inst / cycle (all SMs)"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "FFMA thread-inst / cycle (all SMs)"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of `{rep}`\n")
    print("Times under ncu are cold-cache and serialised (replayed ~40x): compare shares, not absolutes.\n")
    for r in data:
        print(f"## {r[col['Kernel Name']]}\n")
        print("| metric | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in col:
                print(f"| {label} (`{k}`) | {r[col[k]]} {units[col[k]]} |")
        try:
            fa = float(r[col["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed"]])
            fm = float(r[col["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed"]])
            ff = float(r[col["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"]])
            cyc = float(r[col["sm__cycles_elapsed.max"]])
            dur = float(r[col["gpu__time_duration.sum"]])
            dur_s = dur * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(units[col["gpu__time_duration.sum"]].strip(), 1e-9)
            flop = (fa + fm + 2 * ff) * cyc
            print(f"| EXECUTED fp32 flop (FADD+FMUL+2*FFMA) | {flop:.4e} |")
            print(f"| EXECUTED fp32 TFLOP/s | {flop / dur_s / 1e12:.2f} |")
            print(f"| fp32 lanes busy (FADD+FMUL+FFMA thread-inst / (148*128) / cycle) | {(fa + fm + ff) / 18944 * 100:.1f} % |")
        except (KeyError, ValueError):
            pass
        print()


if __name__ == "__main__":
    main()
