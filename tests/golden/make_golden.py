#!/usr/bin/env python3
"""Generate the committed golden fixtures from the UNMODIFIED reference compiled by oracle/build_ref.sh.

Run in the build container (needs /root/reference to have built oracle/_ref):
    python tests/golden/make_golden.py
Each tests/golden/<case>_w<weno>.npz holds, for a small grid of one BASELINE config: the raw initial condition
(ic_U, ic_T), the reference's conserved state after 1 and 10 steps (AoS incl. ghosts), T after 10 steps and the dt
sequence.  These pin the CPU restatement (oracle/xf_oracle.cpp) and, on the GPU box where /root/reference does not
exist, the CUDA path directly."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xfref

GRID = {"shock-tube": (400, 0, 0), "vortex": (32, 32, 0), "riemann": (32, 32, 0), "sbi": (24, 12, 12), "jet": (24, 12, 12)}
VARIANTS = [("shock-tube", 5), ("shock-tube", 7), ("vortex", 5), ("riemann", 5), ("sbi", 5), ("sbi", 7), ("jet", 5)]

# flux-splitting variants other than LLF (Artificial_type 1 = ROE, 3 = GLF; the GLF running maximum is never reset, so 10 steps
# exercise that driver-state semantic): tests/golden/<case>_w5_<alpha>.npz
ALPHA_VARIANTS = [("vortex", 5, 3), ("vortex", 5, 1), ("sbi", 5, 3), ("sbi", 5, 1)]

# (f) rows: WENO-CU6 (SCHEME_ORDER 6) and the positivity-preserving flux limiter (oracle/cases/<case>_pp.json: PP on, CFL 0.9 --
# at that CFL the limiter acts from the first step on): tests/golden/<case>_w<weno>[_pp].npz
NEXT_VARIANTS = [("sbi", 6, 0), ("shock-tube", 6, 0), ("jet", 6, 0), ("sbi", 5, 1), ("sbi", 6, 1), ("shock-tube", 5, 1)]

if __name__ == "__main__":
    for case, weno, pp in NEXT_VARIANTS:
        res = GRID[case]
        A, meta, out = xfref.run_ref(case, res, 10, dump_steps=(1, 10), weno=weno, stage_dump=True, pp=pp)
        assert "ORACLE_TIMING" in out and "error=0" in out, out[-2000:]
        np.savez_compressed(os.path.join(xfref.GOLDEN, "%s_w%d%s.npz" % (case, weno, "_pp" if pp else "")), res=np.array(res), weno=weno, pp=pp,
                            cfl=xfref.PP_CFL if pp else xfref.CASES[case]["cfl"],
                            ic_U=A["ic_U"], ic_T=A["ic_T"], U_step1=A["U_step1"], U_step10=A["U_step10"], T_step10=A["T_step10"],
                            s1_LU=A["s1_LU"], s1_Fwx=A["s1_Fwx"], dt=np.array(meta["dt"]))
        print(case, weno, "pp" if pp else "", "ok", meta["dt"][:2])
    if "--next-only" in sys.argv:
        sys.exit(0)
    for case, weno, alpha in ALPHA_VARIANTS:
        res = GRID[case]
        A, meta, out = xfref.run_ref(case, res, 10, dump_steps=(1, 10), weno=weno, stage_dump=True, alpha=alpha)
        assert "ORACLE_TIMING" in out and "error=0" in out, out[-2000:]
        np.savez_compressed(os.path.join(xfref.GOLDEN, "%s_w%d_%s.npz" % (case, weno, xfref.ALPHA_NAME[alpha].lower())), res=np.array(res), weno=weno, alpha=alpha,
                            ic_U=A["ic_U"], ic_T=A["ic_T"], U_step1=A["U_step1"], U_step10=A["U_step10"], T_step10=A["T_step10"],
                            s1_LU=A["s1_LU"], dt=np.array(meta["dt"]))
        print(case, weno, xfref.ALPHA_NAME[alpha], "ok", meta["dt"][:2])
    if "--alpha-only" in sys.argv:
        sys.exit(0)
    for case, weno in VARIANTS:
        res = GRID[case]
        A, meta, out = xfref.run_ref(case, res, 10, dump_steps=(1, 10), weno=weno, stage_dump=True)
        assert "ORACLE_TIMING" in out and "error=0" in out, out[-2000:]
        np.savez_compressed(os.path.join(xfref.GOLDEN, "%s_w%d.npz" % (case, weno)), res=np.array(res), weno=weno,
                            ic_U=A["ic_U"], ic_T=A["ic_T"], U_step1=A["U_step1"], U_step10=A["U_step10"], T_step10=A["T_step10"],
                            s1_LU=A["s1_LU"], dt=np.array(meta["dt"]))
        print(case, weno, "ok", meta["dt"][:2])
