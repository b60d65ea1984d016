"""GPU (-m gpu): field output of the `xfluids` executable (xfluids_b200/host/xfh_output.cpp) in the reference's file formats -- VTK ImageData
with raw appended Float32 blocks (.vti + .pvti), the compressed-dimension variant (CVTI) and Tecplot ASCII (CPLT .dat) -- compared BYTE FOR
BYTE with the files the unmodified reference wrote for the same case, grid and step count (tests/golden/out/, tests/golden/make_golden_out.py;
reference: XFLUIDS::Output / Output_svti / Output_cvti / Output_cplt, src/XFLUIDS.cpp:1024-1793)."""
import json
import os
import subprocess

import pytest

import xfref

pytestmark = pytest.mark.gpu
EXE = os.path.join(xfref.REPO, "xfluids_b200", "xfluids")
GOLD = os.path.join(xfref.GOLDEN, "out")


def strip_comments(text):
    # the reference's JSON files carry //-comments (json.hpp ignore_comments); the case files under oracle/cases have none
    return text


@pytest.mark.parametrize("stamp", [0, 1, 2])
def test_field_files_equal_reference_bytes(tmp_path, stamp):
    import xfgpu  # noqa: F401  (fails loudly if the CUDA library is missing)
    case = json.load(open(os.path.join(xfref.REPO, "oracle", "cases", "sbi_out.json")))
    stamps = case["run"]["OutTimeStamps"]
    case["run"]["OutTimeStamps"] = [stamps[stamp]]          # the final output uses the LAST stamp's format (XFLUIDS.cpp:309)
    js = tmp_path / "case.json"
    js.write_text(json.dumps(case))
    out = tmp_path / "output"
    out.mkdir()
    r = subprocess.run([EXE, str(js), "-run=16,8,8,2", "-sample=shock-bubble", "-mixture=Inert-SBI", "-weno=5", "-alpha=LLF", "-dv=host", "-out", "-outdir=" + str(out), "-quiet"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    want = sorted(os.listdir(os.path.join(GOLD, str(stamp))))
    assert want, "no golden files"
    for f in want:
        mine = out / f
        assert mine.exists(), "%s not written (have: %s)" % (f, sorted(os.listdir(out)))
        a, b = mine.read_bytes(), open(os.path.join(GOLD, str(stamp), f), "rb").read()
        if a != b:
            n = next(i for i in range(min(len(a), len(b))) if a[i] != b[i]) if len(a) == len(b) or a[:min(len(a), len(b))] != b[:min(len(a), len(b))] else min(len(a), len(b))
            raise AssertionError("%s differs from the reference's file at byte %d of %d / %d:\n mine %r\n ref  %r" % (f, n, len(a), len(b), a[max(0, n - 60):n + 60], b[max(0, n - 60):n + 60]))
