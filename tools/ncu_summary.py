#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers the roofline needs.

    python tools/ncu_summary.py gpurun_out/prof_cfg2.ncu-rep [...] > profiles/rNN_summary.md

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and prints,
per captured launch: duration, DRAM bytes read / written, DRAM throughput %, issue
rate, pipe utilisation, occupancy limits and the warp stall reasons (pc sampling).
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    # the store path (rotated frames: runs of 24 bytes in 32 different rows per store instruction)
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_global_op_st.sum",   # (not in every ncu version)
    "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_l1tex2xbar_write_bytes.sum.per_second",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "lts__d_sectors_fill_sysmem.sum",
    "lts__t_sectors_lookup_miss.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def raw_rows(path: str):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main() -> None:
    for path in sys.argv[1:]:
        hdr, units, rows = raw_rows(path)
        col = {k: i for i, k in enumerate(hdr)}
        for r in rows:
            print(f"## {path.split('/')[-1]} -- launch {r[col['ID']]}: {r[col['Kernel Name']][:110]}")
            print()
            print("| metric | value | unit |")
            print("|---|---|---|")
            for k in KEYS:
                if k in col:
                    print(f"| {k} | {r[col[k]]} | {units[col[k]]} |")
            stalls = []
            for k, i in col.items():
                if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                    try:
                        stalls.append((float(r[i]), k.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                    except ValueError:
                        pass
            tot = sum(v for v, _ in stalls) or 1.0
            print()
            print("warp states (pc samples): " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(stalls, reverse=True)[:9]))
            print()


if __name__ == "__main__":
    main()
