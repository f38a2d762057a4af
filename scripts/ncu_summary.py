#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small tracked text table for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: ncu --set full --clock-control none (one block per profiled launch)")
    for r in rows[2:]:
        print(f"\nkernel: {r[idx['Kernel Name']][:110]}")
        for m, short in METRICS:
            if m in idx:
                print(f"  {short:22s} {r[idx[m]]:>18s} {units[idx[m]]}")
        if "dram__bytes_read.sum" in idx and "gpu__time_duration.sum" in idx:
            def val(m):
                v, u = float(r[idx[m]].replace(",", "")), units[idx[m]]
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
                return v * scale.get(u, 1.0)
            t = val("gpu__time_duration.sum")
            b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            print(f"  {'dram_GBps(derived)':22s} {b / t / 1e9:18.1f} GB/s")


if __name__ == "__main__":
    main()
