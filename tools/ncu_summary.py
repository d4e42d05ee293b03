#!/usr/bin/env python3
"""Summarise an ncu report (--set full [+ pipe instruction counts]) for profiles/: one JSON entry per
kernel launch with the figures bench.py's roofline uses.

    ncu -i X.ncu-rep --page raw --csv > X_raw.csv
    python tools/ncu_summary.py X_raw.csv name=<entry name> cells=<DP cells per launch> [algo_bytes=..] [launch="..."]

Prints the entry (add it to profiles/traffic_r2.json).  alu_thread_inst_per_cell = warp instructions
issued on the ALU pipe x 32 / DP cells of the launch: the numerator of the pipe-aware roofline."""
import csv
import json
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "inst": 1, "%": 1, "register/thread": 1, "": 1, "warp": 1,
        "cycle": 1}


def main():
    path = sys.argv[1]
    kw = dict(a.split("=", 1) for a in sys.argv[2:])
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    pick = int(kw.get("row", 0))
    r = data[pick]

    def val(name, default=None):
        if name not in col or r[col[name]] in ("", "no data"):
            return default
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1)

    cells = float(kw["cells"]) if "cells" in kw else None
    alu = val("sm__inst_executed_pipe_alu.sum")
    fma = val("sm__inst_executed_pipe_fma.sum")
    lsu = val("sm__inst_executed_pipe_lsu.sum")
    tot = val("smsp__inst_executed.sum")
    e = {
        "kernel": r[col["Kernel Name"]][:90],
        "launch": kw.get("launch", ""),
        "duration_ms_under_ncu": val("gpu__time_duration.sum"),
        "grid": val("launch__grid_size"), "registers": val("launch__registers_per_thread"),
        "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": val("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "warp_inst": tot, "alu_warp_inst": alu, "fma_warp_inst": fma, "lsu_warp_inst": lsu,
        "stall_per_issue": {h.split("issue_stalled_")[1].split("_per_issue")[0]: round(float(r[i]), 3)
                            for h, i in col.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio")
                            and r[i] not in ("", "no data") and float(r[i]) >= 0.05},
    }
    if e["dram_read_bytes"] is not None:
        e["dram_bytes"] = e["dram_read_bytes"] + e["dram_write_bytes"]
    if cells:
        e["cells"] = cells
        if alu:
            e["alu_thread_inst_per_cell"] = 32 * alu / cells
        if tot:
            e["thread_inst_per_cell"] = 32 * tot / cells
    if "algo_bytes" in kw:
        e["algorithmic_bytes"] = float(kw["algo_bytes"])
    print(json.dumps({kw.get("name", "kernel"): e}, indent=1))


if __name__ == "__main__":
    main()
