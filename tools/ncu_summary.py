#!/usr/bin/env python
"""Condenses an `ncu --set full` report into the handful of numbers DESIGN.md quotes per kernel: duration, DRAM bytes and
GB/s, tensor-pipe %, XU / FMA / ALU pipe %, issue-active %, registers, achieved warps, the top warp-stall reasons.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt     (runs where ncu is installed; no GPU needed)"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"), ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "XU (SFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "ALU pipe active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "LSU pipe %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of max"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"== {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print(f"   {label:34s} {d[k]} {u.get(k, '')}")
        stalls = []
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v:
                try:
                    stalls.append((float(v), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        for v, name in sorted(stalls, reverse=True)[:6]:
            print(f"   stall {name:28s} {v:.2f} warps per issue-active cycle")
        print()


if __name__ == "__main__":
    main()
