#!/usr/bin/env python3
"""profiles/traffic.json from the committed ncu --set full summaries (tools/ncu_summary.py full) of one round:

    tools/make_traffic.py r02

For every workload with a summary: DRAM read + write bytes per launch of each hot kernel (mean over the captured launches), its FP64-pipe
utilisation, and -- for the slowest sweep -- the executed DADD / DMUL / DFMA thread instructions per face (bench.py turns those into
roofline.frac_issue).  bench.py reads the file; nothing else does."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PROF = os.path.join(os.path.dirname(HERE), "profiles")
OPS = {  # tag of profiles/<round>_fp64ops_<tag>.csv (ncu --metrics ...op_d{add,mul,fma}... --csv of the sweeps) -> traffic.json key, faces
    "sbi512": "sbi:512x512x512:weno5", "w7": "sbi:512x512x512:weno7", "cu6pp": "sbi:512x512x512:weno6", "jet": "jet:1024x512x512:weno5",
    "riemann": "riemann:4096x4096x1:weno5", "vortex": "vortex:1024x1024x1:weno5", "visc": "sbi:512x512x512:weno5:visc"}
WORKLOADS = {  # summary file tag -> (traffic.json key, faces per sweep launch by direction)
    "sbi512": ("sbi:512x512x512:weno5", {"x": 513 * 512 * 512, "y": 512 * 513 * 512, "z": 512 * 512 * 513}),
    "w7": ("sbi:512x512x512:weno7", {"x": 513 * 512 * 512, "y": 512 * 513 * 512, "z": 512 * 512 * 513}),
    "riemann": ("riemann:4096x4096x1:weno5", {"x": 4097 * 4096, "y": 4096 * 4097}),
    "vortex": ("vortex:1024x1024x1:weno5", {"x": 1025 * 1024, "y": 1024 * 1025}),
    "visc": ("sbi:512x512x512:weno5:visc", {"x": 513 * 512 * 512, "y": 512 * 513 * 512, "z": 512 * 512 * 513}),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def num(v):
    v = v.strip().replace(",", "")
    m = re.match(r"^([-+0-9.eE]+)\s*(\S*)", v)
    x = float(m.group(1))
    return x * UNIT.get(m.group(2), 1.0)


def blocks(path):
    out, cur = [], None
    for line in open(path):
        if line.startswith("## "):
            cur = {"name": line[3:].strip()}
            out.append(cur)
        elif cur is not None and line.startswith("| ") and "`" in line:
            m = re.match(r"\| .*\(`([^`]+)`\) \| (.*) \|\s*$", line)
            if m:
                try:
                    cur[m.group(1)] = num(m.group(2))
                except (AttributeError, ValueError):
                    pass
    return out


def kind(name):
    m = re.search(r"k_sweep<XfCfg<\d+, \d+>, (\d),", name)
    if "k_sweep" in name:
        d = int(m.group(1)) if m else None
        return "sweep_" + "xyz"[d] if d is not None else "sweep"
    if "k_prim_hard" in name:
        return "prim_hard"
    if "k_prim" in name:
        return "prim"
    if "k_rk" in name:
        return "lu_rk"
    return None


def main(rnd):
    res = {"source": "profiles/%s_ncu_full_*.md: dram__bytes_read.sum + dram__bytes_write.sum per launch, FP64 pipe utilisation and executed FP64 thread instructions "
                     "(ncu --set full --clock-control none, tools/gpu_round2.sh; written by tools/make_traffic.py)" % rnd}
    for tag, (key, faces) in WORKLOADS.items():
        p = os.path.join(PROF, "%s_ncu_full_%s.md" % (rnd, tag))
        if not os.path.exists(p):
            continue
        acc = {}
        for b in blocks(p):
            k = kind(b["name"])
            if k is None or "dram__bytes_read.sum" not in b:
                continue
            a = acc.setdefault(k, {"n": 0, "bytes": 0.0, "pipe": 0.0, "ns": 0.0, "dadd": 0.0, "dmul": 0.0, "dfma": 0.0})
            a["n"] += 1
            a["bytes"] += b["dram__bytes_read.sum"] + b.get("dram__bytes_write.sum", 0.0)
            a["pipe"] += b.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0)
            a["ns"] += b.get("gpu__time_duration.sum", 0.0)
            a["dadd"] += b.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
            a["dmul"] += b.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
            a["dfma"] += b.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
        if not acc:
            continue
        w = {k: int(round(a["bytes"] / a["n"])) for k, a in acc.items()}
        sweeps = [k for k in acc if k.startswith("sweep_")]
        ex = {"fp64_pipe_pct_of_peak": {k: round(a["pipe"] / a["n"], 1) for k, a in acc.items()}, "per_sweep": {}}
        for k in sweeps:
            a, f = acc[k], faces[k[-1]]
            if a["dadd"] + a["dmul"] + a["dfma"] > 0:
                ex["per_sweep"][k] = {"DADD": round(a["dadd"] / a["n"] / f, 1), "DMUL": round(a["dmul"] / a["n"] / f, 1), "DFMA": round(a["dfma"] / a["n"] / f, 1)}
        if ex["per_sweep"]:
            slow = max(ex["per_sweep"], key=lambda k: acc[k]["ns"] / acc[k]["n"])
            ops = ex["per_sweep"][slow]
            ex["kernel"] = "k_sweep<%s>" % slow[-1]
            ex["fp64_thread_inst_per_face"] = {k: int(round(v)) for k, v in ops.items()}
            ex["flop_per_face_fma2"] = int(round(ops["DADD"] + ops["DMUL"] + 2 * ops["DFMA"]))
            ex["note"] = "smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum of the capture / faces per launch"
        w["executed"] = ex
        res[key] = w
    import csv
    for tag, key in OPS.items():
        p = os.path.join(PROF, "%s_fp64ops_%s.csv" % (rnd, tag))
        if not os.path.exists(p):
            continue
        dims = [int(x) for x in key.split(":")[1].split("x")]
        per = {}
        for r in csv.reader(open(p)):
            if len(r) < 8 or not r[0].isdigit():
                continue
            name, metric, val = r[4], r[-3], r[-1]
            k = kind(name)
            if k is None or not k.startswith("sweep_"):
                continue
            try:
                per.setdefault(k, {}).setdefault(metric, []).append(float(val.replace(",", "")))
            except ValueError:
                pass
        if not per:
            continue
        w = res.setdefault(key, {})
        ex = w.setdefault("executed", {"fp64_pipe_pct_of_peak": {}, "per_sweep": {}})
        mean = lambda v: sum(v) / len(v)
        for k, m in per.items():
            ax = "xyz".index(k[-1])
            faces = 1
            for a in range(3):
                faces *= (dims[a] + 1) if a == ax else dims[a]
            ex["per_sweep"][k] = {op.upper(): round(mean(m["smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op]) / faces, 1) for op in ("dadd", "dmul", "dfma")}
            ex["per_sweep"][k]["thread_inst_all"] = round(mean(m["smsp__inst_executed.sum"]) * 32 / faces, 1) if "smsp__inst_executed.sum" in m else None
            ex["per_sweep"][k]["ms"] = round(mean(m["gpu__time_duration.sum"]) / 1e6, 3)
            ex["fp64_pipe_pct_of_peak"][k] = round(mean(m["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]), 1)
        slow = max(ex["per_sweep"], key=lambda k: ex["per_sweep"][k]["ms"])
        ops = ex["per_sweep"][slow]
        ex["kernel"] = "k_sweep<%s>" % slow[-1]
        ex["fp64_thread_inst_per_face"] = {k: int(round(ops[k])) for k in ("DADD", "DMUL", "DFMA")}
        ex["flop_per_face_fma2"] = int(round(ops["DADD"] + ops["DMUL"] + 2 * ops["DFMA"]))
        ex["note"] = "smsp__sass_thread_inst_executed_op_{dadd,dmul,dfma}_pred_on.sum per launch / faces per launch (profiles/%s_fp64ops_%s.csv)" % (rnd, tag)
    json.dump(res, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
