#!/usr/bin/env python
"""Summarise `ncu --set full` captures (raw CSV pages) of the step's kernels into profiles/r02_ncu_kernels.json, keyed by
the number of objects per GPU: per kernel the launch duration, DRAM bytes per launch, FP64 pipe utilisation, issue-slot
utilisation, active lanes per instruction, registers, achieved occupancy and the main stall reasons.
    ncu -i X.ncu-rep --page raw --csv > X.csv ;  python tools/ncu_summary.py N_OBJECTS X.csv [N2 Y.csv ...] > out.json"""
import csv
import json
import re
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_inst",
    "launch__registers_per_thread": "registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    # hardware-counted FP64 arithmetic: thread instructions per elapsed cycle, summed over the sub-partitions
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed": "dadd_per_cycle",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed": "dmul_per_cycle",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed": "dfma_per_cycle",
    "smsp__cycles_elapsed.avg": "sm_cycles",
    "smsp__cycles_elapsed.avg.per_second": "sm_ghz",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}


def summarise(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        m = re.search(r"(k_\w+|ssa_\w+)", name)
        short = m.group(1) if m else name
        d = out.setdefault(short, {"launches": 0})
        d["launches"] += 1
        for k, nk in KEYS.items():
            if k in ix:
                v = float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1.0)
                d[nk] = d.get(nk, 0.0) + v
    for d in out.values():
        n = d.pop("launches")
        for k in list(d):
            d[k] = d[k] / n
        d["launches_averaged"] = n
        d["dram_bytes_per_launch"] = d.get("dram_read", 0.0) + d.get("dram_write", 0.0)
        if "dfma_per_cycle" in d and "sm_cycles" in d:  # DADD + DMUL + 2 DFMA executed by the threads of one launch
            d["fp64_flop_per_launch_hw"] = (d["dadd_per_cycle"] + d["dmul_per_cycle"] + 2.0 * d["dfma_per_cycle"]) * d["sm_cycles"]
            d["fp64_tflops_hw"] = d["fp64_flop_per_launch_hw"] / (d["duration_us"] * 1e-6) / 1e12
    return out


def main():
    res = {}
    args = sys.argv[1:]
    for n, path in zip(args[0::2], args[1::2]):
        s = summarise(path)
        tot = sum(v["dram_bytes_per_launch"] for v in s.values())
        dur = sum(v["duration_us"] for v in s.values())
        s["_step"] = {"dram_bytes": tot, "dram_bytes_per_object": tot / float(n), "kernel_time_us": dur,
                      "fp64_pipe_pct_time_weighted": sum(v["fp64_pipe_pct"] * v["duration_us"] for k, v in s.items() if k != "_step") / dur}
        res[str(int(n))] = s
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
