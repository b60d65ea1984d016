#!/usr/bin/env python3
"""Static SASS opcode histogram of the hot kernels of libxfluids_b200.so (cuobjdump -sass), written as markdown:

    tools/sass_hist.py > profiles/r02_sass_histogram.md

Static counts (instructions in the binary, not executed): they show WHICH instructions a kernel is made of -- FP64 arithmetic by
opcode, the TMA / mbarrier instructions of the staged pencils, the division sequences (MUFU.RCP64H) -- next to the executed counts
that ncu reports (profiles/traffic.json)."""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "xfluids_b200", "libxfluids_b200.so")
# (label, regex on the mangled name): strict flavour, SBI configuration (5 species, Emax = 9)
KERNELS = OrderedDict([
    ("k_sweep x, WENO5", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi0ELi5ELb0ELb0ELb0E"),
    ("k_sweep y, WENO5 (TMA-staged pencil)", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi1ELi5ELb0ELb0ELb0E"),
    ("k_sweep z, WENO5 (TMA-staged pencil)", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi2ELi5ELb0ELb0ELb0E"),
    ("k_sweep z, WENO5 + viscous tail", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi2ELi5ELb0ELb0ELb1E"),
    ("k_sweep z, WENO7", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi2ELi7ELb0ELb0ELb0E"),
    ("k_sweep z, WENO-CU6 + limiter", r"xf_strict7k_sweepI5XfCfgILi5ELb1EELi2ELi6ELb1ELb0ELb0E"),
    ("k_prim", r"xf_strict6k_primI5XfCfgILi5ELb1EEE"),
    ("k_prim_hard", r"xf_strict11k_prim_hardI5XfCfgILi5ELb1EEE"),
    ("k_rk<9, fused divergence>", r"xf_strict4k_rkILi9ELb1EE"),
    ("k_transport", r"xf_strict11k_transportI5XfCfgILi5ELb1EEE"),
    ("k_march z, WENO5 (opt-in)", r"xf_strict7k_marchI5XfCfgILi5ELb1EELi2ELi5ELb0E"),
])
GROUPS = OrderedDict([
    ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("DFMA", r"^DFMA"), ("DSETP/DMNMX", r"^(DSETP|DMNMX)"), ("MUFU.RCP64H", r"^MUFU\.RCP64H"), ("MUFU.RSQ64H", r"^MUFU\.RSQ64H"),
    ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDL/STL", r"^(LDL|STL)"), ("LDC/ULDC", r"^(LDC|ULDC)"),
    ("UTMALDG", r"^UTMALDG"), ("SYNCS", r"^SYNCS"), ("LDGSTS (cp.async)", r"^LDGSTS"), ("BAR", r"^BAR"), ("CALL/RET", r"^(CALL|RET)"), ("BRA", r"^BRA"),
    ("ATOM/RED", r"^(ATOM|RED|ATOMG)"), ("SHFL", r"^SHFL"),
])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    print("# Static SASS opcode histogram of the hot kernels (cuobjdump -sass xfluids_b200/libxfluids_b200.so; sm_100a, strict flavour, 5 species / Emax = 9)\n")
    print("Instructions in the binary, not executed counts (those: profiles/traffic.json, from ncu).  `UTMALDG` = `cp.async.bulk.tensor` (TMA), `SYNCS` = mbarrier arrive /")
    print("try_wait, `LDGSTS` = `cp.async`; `MUFU.RCP64H` starts every FP64 division / reciprocal sequence; `CALL` includes the out-of-line WENO / exp / pow bodies")
    print("(whose instructions are counted inside the calling kernel's function by cuobjdump).\n")
    cols = list(GROUPS)
    print("| kernel | total | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    for label, rx in KERNELS.items():
        names = [n for n in funcs if re.search(rx, n)]
        if not names:
            print("| %s | (not in this build) |" % label + " |" * len(cols))
            continue
        c = funcs[names[0]]
        row = []
        for g, grx in GROUPS.items():
            row.append(sum(v for k, v in c.items() if re.match(grx, k)))
        print("| %s | %d | " % (label, sum(c.values())) + " | ".join(str(x) for x in row) + " |")
    print("\nFull opcode list of the z sweep (WENO5), most frequent first:\n")
    names = [n for n in funcs if re.search(KERNELS["k_sweep z, WENO5 (TMA-staged pencil)"], n)]
    if names:
        print(", ".join("%s %d" % kv for kv in funcs[names[0]].most_common(40)))


if __name__ == "__main__":
    main()
