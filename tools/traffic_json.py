#!/usr/bin/env python
"""profiles/rNN_<kernel>.txt digests (tools/ncu_summary.py) -> profiles/rNN_traffic.json, the per-launch
DRAM bytes and pipe utilisations bench.py quotes in `roofline.traffic` / `roofline.ncu`.
usage: python tools/traffic_json.py profiles/r01_*_kernel.txt > profiles/r01_traffic.json"""
import json
import re
import sys

WANT = {
    "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "dur",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct_of_peak",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed_pipe_alu.sum": "alu_inst", "smsp__inst_executed.sum": "inst",
    "launch__grid_size": "grid", "launch__block_size": "block",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
        "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
out = {}
for path in sys.argv[1:]:
    launches = []
    cur = None
    for line in open(path):
        m = re.match(r"kernel: (?:void )?(?:mtsv::)?(\w+)", line)
        if m:
            cur = {"name": m.group(1)}
            launches.append(cur)
            continue
        f = line.split()
        if cur is not None and len(f) >= 2 and f[0] in WANT:
            v = float(f[1].replace(",", ""))
            unit = f[2] if len(f) > 2 else ""
            if WANT[f[0]] in ("rd", "wr", "dur"):
                v *= UNIT.get(unit, 1.0)
            cur[WANT[f[0]]] = v
    if not launches:
        continue
    name = launches[0]["name"]
    n = len(launches)
    avg = lambda k: sum(l.get(k, 0.0) for l in launches) / n
    out[name] = {"dram_bytes_per_launch": avg("rd") + avg("wr"), "duration_s_under_ncu": avg("dur"),
                 "dram_throughput_pct_of_peak": avg("dram_throughput_pct_of_peak"),
                 "alu_pipe_pct_of_peak": avg("alu_pipe_pct_of_peak"), "fma_pipe_pct_of_peak": avg("fma_pipe_pct_of_peak"),
                 "issue_active_pct": avg("issue_active_pct"), "warps_active_pct": avg("warps_active_pct"),
                 "launches_captured": n, "source": path}
    threads = sum(l.get("grid", 0.0) * l.get("block", 0.0) for l in launches)
    if threads and any("alu_inst" in l for l in launches):
        # thread-per-work-item kernels: ALU-pipe warp instructions per work item (grid x block threads, the last
        # block padded), summed over the captured launches (one step)
        out[name]["alu_warp_inst_per_launch"] = sum(l.get("alu_inst", 0.0) for l in launches) / n
        out[name]["warp_inst_per_launch"] = sum(l.get("inst", 0.0) for l in launches) / n
        out[name]["alu_warp_inst_per_candidate"] = sum(l.get("alu_inst", 0.0) for l in launches) / threads
print(json.dumps(out, indent=1))
