#!/usr/bin/env python3
"""Field-output fixtures: the files the UNMODIFIED reference writes (XFLUIDS::Output, src/XFLUIDS.cpp:1024-1793) for the three kinds of
output stamp of oracle/cases/sbi_out.json -- common domain, compressed dimensions (-C=X,Y,0.0), partial domain (-P=yi[Xe] > 0.01) --
after 2 steps of the 16 x 8 x 8 shock-bubble block (oracle/build_ref.sh sbi 5 parity LLF 0 0 1; oracle/ref_driver.cpp XF_OUTPUT=<stamp>).

    python tests/golden/make_golden_out.py      ->  tests/golden/out/<stamp>/*.{vti,pvti,dat}

tests/test_gpu_output.py compares what the `xfluids` executable writes with these, byte for byte."""
import os
import shutil
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xfref

GRID, NSTEPS = (16, 8, 8), 2

if __name__ == "__main__":
    d = os.path.join(xfref.REF_DIR, "sbi_w5_parity_out")
    assert os.path.exists(os.path.join(d, "XFLUIDS")), "build it: bash oracle/build_ref.sh sbi 5 parity LLF 0 0 1"
    out = os.path.join(d, "output")
    for stamp in (0, 1, 2):
        for f in os.listdir(out):
            if f.endswith((".vti", ".pvti", ".dat")) or "CheckingPoint" in f:
                os.remove(os.path.join(out, f))
        env = dict(os.environ, XF_NSTEPS=str(NSTEPS), XF_OUTPUT=str(stamp))
        r = subprocess.run(["./XFLUIDS", "-run=%d,%d,%d,%d" % (GRID + (NSTEPS,))], cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert "ORACLE_TIMING" in r.stdout and "error=0" in r.stdout, r.stdout[-2000:]
        dst = os.path.join(xfref.GOLDEN, "out", str(stamp))
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(dst)
        for f in sorted(os.listdir(out)):
            if f.endswith((".vti", ".pvti", ".dat")):
                shutil.copy(os.path.join(out, f), os.path.join(dst, f))
                print(stamp, f, os.path.getsize(os.path.join(dst, f)))
