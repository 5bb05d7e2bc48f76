#!/usr/bin/env python
"""Extract the per-kernel numbers bench.py quotes from an `ncu --set full` report into a small tracked JSON:
    python scripts/ncu_metrics_json.py gpurun_out/<tag>/prof.ncu-rep profiles/<tag>_ncu_metrics.json
(run here; the .ncu-rep itself is scratch).  bench.py reads profiles/ncu_metrics.json for `roofline.traffic`."""
import csv, io, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k, scale=True):
        v = float(r[col[k]].replace(",", ""))
        return v * UNIT.get(units[col[k]].strip(), 1.0) if scale else v

    res = {"source": rep, "note": "one ncu --set full capture per kernel (cold cache, replayed): shares, not absolutes",
           "kernels": {}}
    for r in data:
        name = r[col["Kernel Name"]]
        import re
        short = re.search(r"(\w+)\s*<", name.replace("d2d::", "")) or re.search(r"(\w+)\s*\(", name)
        short = short.group(1)
        fa = val(r, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed", False)
        fm = val(r, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed", False)
        ff = val(r, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", False)
        cyc = val(r, "sm__cycles_elapsed.max", False)
        dur = val(r, "gpu__time_duration.sum")
        res["kernels"][short] = {
            "name": name,
            "duration_ms_under_ncu": dur * 1e3,
            "dram_bytes_read": val(r, "dram__bytes_read.sum"),
            "dram_bytes_write": val(r, "dram__bytes_write.sum"),
            "warp_instructions": val(r, "smsp__inst_executed.sum", False),
            "issue_slots_busy_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "achieved_occupancy_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "registers_per_thread": val(r, "launch__registers_per_thread", False),
            "executed_fp32_flop": (fa + fm + 2 * ff) * cyc,
            "fp32_lanes_busy_pct": (fa + fm + ff) / 18944 * 100,
            "stall_no_instruction_per_issue": val(r, "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", False),
        }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
