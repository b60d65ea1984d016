#!/usr/bin/env python3
"""Pins at (or near) BASELINE scale, written from the UNMODIFIED reference compiled by oracle/build_ref.sh (build container only):

    python tests/golden/make_scale_pins.py        ->  tests/golden/scale_pins.json

For each entry the reference (oracle/_ref/<case>_w5_parity, serial, -ffp-contract=off) runs 3 time steps from its own initial
condition on a grid large enough to exercise what the 24x12x12 / 32^2 fixtures cannot: the 32-bit in-plane index arithmetic of
the kernels, several x chunks / y-z tiles per pencil, the 8x8 block order of the update, the hard-cell list of the Newton solver.
Stored: SHA-256 of the raw bytes of ic_U, ic_T and of U after 3 steps (reference checkpoint payload layout, AoS incl. ghosts,
XFLUIDS.cpp:658-687), the dt sequence (hex floats) and, for eyeballing a mismatch, per-z-plane sums of rho after 3 steps.
tests/test_scale_pins.py asserts them: the host initial-condition hooks on the CPU, the CUDA path on the GPU."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xfref

PINS = [("riemann", (1024, 1024, 0), 5), ("vortex", (1024, 1024, 0), 5), ("sbi", (128, 64, 64), 5), ("jet", (128, 64, 64), 5), ("shock-tube", (4000, 0, 0), 5)]
NSTEPS = 3


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


if __name__ == "__main__":
    out = {"nsteps": NSTEPS, "source": "oracle/_ref/<case>_w5_parity (unmodified reference, g++ -O2 -ffp-contract=off, serial)", "pins": {}}
    for case, res, weno in PINS:
        A, meta, log = xfref.run_ref(case, res, NSTEPS, dump_steps=(NSTEPS,), weno=weno, stage_dump=False, dump_T=True)
        assert "ORACLE_TIMING" in log and "error=0" in log, log[-2000:]
        # the raw initial condition needs XF_DUMP_IC
        import subprocess, tempfile
        d = xfref.ref_dir(case, weno)
        tmp = tempfile.mkdtemp(prefix="xfpin_")
        env = dict(os.environ, XF_NSTEPS="0", XF_DUMP_DIR=tmp, XF_DUMP_IC="1", XF_DUMP_STEPS="")
        subprocess.run(["./XFLUIDS", "-run=%d,%d,%d,1" % res], cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        icU, icT = np.fromfile(os.path.join(tmp, "ic_U.bin")), np.fromfile(os.path.join(tmp, "ic_T.bin"))
        U = A["U_step%d" % NSTEPS]
        E = int(meta["Emax"])
        Zmax, Ymax, Xmax = int(meta["Zmax"]), int(meta["Ymax"]), int(meta["Xmax"])
        rho = U.reshape(Zmax, Ymax, Xmax, E)[..., 0]
        out["pins"]["%s_w%d" % (case, weno)] = {
            "case": case, "res": list(res), "weno": weno, "emax": E, "ncells": int(Zmax * Ymax * Xmax),
            "ic_U_sha256": sha(icU), "ic_T_sha256": sha(icT), "U_sha256": sha(U), "T_sha256": sha(A["T_step%d" % NSTEPS]),
            "dt_hex": [float(x).hex() for x in meta["dt"][:NSTEPS]],
            "rho_plane_sums_hex": [float(x).hex() for x in rho.reshape(Zmax, -1).sum(axis=1)][:8],
            "rho_sum_hex": float(rho.sum()).hex()}
        print(case, res, "ok", out["pins"]["%s_w%d" % (case, weno)]["U_sha256"][:16], meta["dt"][:NSTEPS])
    json.dump(out, open(os.path.join(xfref.GOLDEN, "scale_pins.json"), "w"), indent=1)
