#!/usr/bin/env python3
"""Golden fixtures of the viscous / heat-conduction / species-diffusion terms from the UNMODIFIED reference built with
-DVisc=1 -DVisc_Heat=1 -DVisc_Diffu=1 and the fourth-order viscous discretisation (oracle/build_ref.sh <case> <weno> parity <alpha> <pp> 1):

    python tests/golden/make_golden_visc.py      ->  tests/golden/<case>_w<weno>[_glf][_pp]_visc.npz

Per fixture: the raw initial condition, the transport fits the reference's Setup computed (viscfit.cpp), the intermediates of the
viscous block in stage 1 of step 1 (mixture viscosity / conductivity / diffusion coefficients, species enthalpies, the nine velocity
derivatives after their ghost fill, the wall fluxes AFTER the viscous part was subtracted, LU), U after 1 and 10 steps and the dt sequence."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xfref

GRID = {"sbi": (24, 12, 12), "jet": (24, 12, 12)}
# (case, order, splitting, limiter[, grid, tag suffix]); the third is the shipped shock-bubble preset's scheme set: WENO-CU6 + GLF + limiter + viscous;
# the fourth a 2-D block (z inactive: three of the nine derivatives vanish, the z wall flux does not exist)
VARIANTS = [("sbi", 5, 2, 0), ("jet", 5, 2, 0), ("sbi", 6, 3, 1), ("sbi", 5, 2, 0, (32, 16, 0), "2d")]

if __name__ == "__main__":
    for v in VARIANTS:
        case, weno, alpha, pp = v[:4]
        res, sfx = (v[4], v[5]) if len(v) > 4 else (GRID[case], "")
        A, meta, out = xfref.run_ref(case, res, 10, dump_steps=(1, 10), weno=weno, stage_dump=True, alpha=alpha, pp=pp, visc=1)
        assert "ORACLE_TIMING" in out and "error=0" in out, out[-2000:]
        tag = "%s_w%d%s%s_visc%s" % (case, weno, "" if alpha == 2 else "_" + xfref.ALPHA_NAME[alpha].lower(), "_pp" if pp else "", sfx)
        keep = {k: A[k] for k in ("ic_U", "ic_T", "U_step1", "U_step10", "T_step10", "fit_visc", "fit_therm", "fit_Dkj", "s1_visc", "s1_therm", "s1_Dkm", "s1_hi",
                                  "s1_LU", "s1_Fwx", "s1_Fwy", "s1_Fwz") + tuple("s1_Vde%d" % m for m in range(9)) if k in A}
        np.savez_compressed(os.path.join(xfref.GOLDEN, tag + ".npz"), res=np.array(res), weno=weno, alpha=alpha, pp=pp,
                            cfl=xfref.PP_CFL if pp else xfref.CASES[case]["cfl"], dt=np.array(meta["dt"]), **keep)
        print(tag, "ok", meta["dt"][:2])
