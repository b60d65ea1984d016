#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` output: one line per kernel with registers / spills / stack / smem."""
import re, sys, subprocess
txt = sys.stdin.read()
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        cur["stack"], cur["sst"], cur["sld"] = m.groups()
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = m.group(1)
for r in rows:
    print("%-70s regs=%-4s stack=%-5s spill_st=%-5s spill_ld=%s" % (r["name"][-70:], r.get("regs"), r.get("stack"), r.get("sst"), r.get("sld")))
