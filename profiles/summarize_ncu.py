#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the handful of counters DESIGN.md argues from.

    python profiles/summarize_ncu.py gpurun_out/prof_raw.csv > profiles/rNN_xxx_summary.md
"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs, blocks)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem, blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("### %s" % d.get("Kernel Name", "?"))
        print()
        print("| counter | value |")
        print("|---|---|")
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("| %s (`%s`) | %s %s |" % (label, k, d[k], u.get(k, "")))
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            print("| DRAM traffic per launch (read+write) | %.3f %s |" % (float(rd) + float(wr), u["dram__bytes_read.sum"]))
        stalls = [(k[len("smsp__pcsamp_warps_issue_stalled_"):], float(d[k])) for k in hdr
                  if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and d[k] not in ("", "0")]
        tot = sum(v for _, v in stalls) or 1.0
        stalls.sort(key=lambda kv: -kv[1])
        print()
        print("PC-sample stall mix: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in stalls[:8]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
