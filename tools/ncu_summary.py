#!/usr/bin/env python3
"""Summarise ncu captures into the markdown kept under profiles/.

    tools/ncu_summary.py full  gpurun_out/prof.ncu-rep      one block per profiled launch (from `ncu --set full`)
    tools/ncu_summary.py list  gpurun_out/launches.csv      per-kernel totals and shares (from `--metrics gpu__time_duration.sum --csv`)
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occ. limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ. limit smem (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-inst"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-inst"),
    ("smsp__inst_executed.sum", "warp inst executed"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sass__inst_executed_local_loads", "local loads"),
    ("sass__inst_executed_local_stores", "local stores"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full summary of `%s`\n" % rep)
    for r in rows[2:]:
        print("## %s\n" % r[col["Kernel Name"]].split("(")[0].replace("void ", ""))
        print("| metric | value |\n|---|---|")
        for key, label in FULL:
            if key in col:
                print("| %s (`%s`) | %s %s |" % (label, key, r[col[key]], units[col[key]]))
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[col[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("| top stall reasons (warps stalled per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:7]))
        print()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    tot = OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        t = float(r[-1])
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    s = sum(v[1] for v in tot.values())
    print("# ncu launch list `%s`: %d launches, %.3f ms of kernel time (cold-cache, serialised: compare shares)\n" % (path, len(rows), s / 1e6))
    print("| kernel | launches | total ms | mean ms | share % |\n|---|---|---|---|---|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.3f | %.1f |" % (name, n, t / 1e6, t / 1e6 / n, 100.0 * t / s))


if __name__ == "__main__":
    {"full": full, "list": launches}[sys.argv[1]](sys.argv[2])
