#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [...] > profiles/rNN_X.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__sectors_read.sum", "dram__sectors_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print("# %s: no kernels" % path)
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print("# %s\nkernel: %s" % (path, name[:160]))
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print("  %-82s %16s %s" % (k, r[i], units[i]))
            stalls = []
            for i, h in enumerate(hdr):
                pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
                if h.startswith(pre) and h.endswith(suf):
                    try:
                        stalls.append((float(r[i]), h[len(pre):-len(suf)]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("  stalled warps per issue-active cycle, top reasons: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:6]))
            print()


if __name__ == "__main__":
    main()
