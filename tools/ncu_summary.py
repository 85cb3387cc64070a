#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into a per-launch summary (JSON) for profiles/.
usage: ncu -i rep --page raw --csv | python tools/ncu_summary.py > profiles/NAME.json"""
import csv, json, sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"id": int(r[hdr.index("ID")]), "kernel": r[hdr.index("Kernel Name")]}
    for k, name in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            v *= UNIT_SCALE.get(units[i], 1.0)
            d[name] = round(v, 3)
    if "dram_read" in d and "dram_write" in d:
        d["dram_bytes"] = d["dram_read"] + d["dram_write"]
    out.append(d)
json.dump(out, sys.stdout, indent=1)
