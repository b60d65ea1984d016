"""The device logarithm (xfluids_b200/csrc/xf_log.cuh) against the libm log() of the reference CPU path, bit for bit.

log(T) in the NASA-9 enthalpy (reference src/solver_Ini/Thermo_device.h:62-80) is the one operation of the inviscid path that is not an
IEEE-754 basic operation; everything multi-species parity rests on is this equality.  CPU: the host restatement of the same operation
sequence over 1.5e8 arguments (tools/check_xf_log.cpp) + the generated table is the one of this machine's libm.  GPU: the device
function over 1e8 arguments of [200, 6000] and 2e7 random normal doubles."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_log_table_is_this_libms_table():
    tmp = tempfile.mkdtemp(prefix="xflog_")
    hdr = os.path.join(REPO, "xfluids_b200", "csrc", "xf_log_data.h")
    cur = open(hdr).read()
    try:
        subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "gen_log_table.py")], stdout=subprocess.DEVNULL)
        assert open(hdr).read() == cur, "xf_log_data.h differs from the table of this machine's libm"
    finally:
        open(hdr, "w").write(cur)
    del tmp


def test_host_restatement_bitwise_equal_to_libm_over_1e8_arguments():
    exe = os.path.join(tempfile.mkdtemp(prefix="xflog_"), "check_xf_log")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-mfma", "-ffp-contract=off", os.path.join(REPO, "tools", "check_xf_log.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "120000000"], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.startswith("mismatches=0 of "), out.stdout
    assert int(out.stdout.split()[-1]) >= 100000000


def _libm_log(x):
    import xfref
    if not os.path.exists(xfref.ORACLE_SO):
        subprocess.check_call([os.path.join(REPO, "oracle", "build_oracle.sh")])
    L = C.CDLL(xfref.ORACLE_SO)
    L.xo_libm_log.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    y = np.empty_like(x)
    L.xo_libm_log(x.ctypes.data, y.ctypes.data, x.size)
    return y


@pytest.mark.gpu
def test_device_log_bitwise_equal_to_libm_over_1e8_arguments():
    from xfluids_b200 import capi
    n, chunks, bad = 20_000_000, 5, 0
    for c in range(chunks):                      # 1e8 equidistant arguments of [200, 6000]
        i = np.arange(c * n, (c + 1) * n, dtype=np.float64)
        x = 200.0 + 5800.0 * (i + 0.5) / float(n * chunks)
        bad += int(np.count_nonzero(capi.log_eval(x).view(np.uint64) != _libm_log(x).view(np.uint64)))
    assert bad == 0, "%d of 1e8 arguments in [200, 6000] differ" % bad
    rng = np.random.default_rng(7)               # positive normal doubles of every magnitude, and the neighbourhood of 1 (fallback path)
    bits = (rng.integers(1, 2046, 20_000_000, dtype=np.uint64) << np.uint64(52)) | rng.integers(0, 1 << 52, 20_000_000, dtype=np.uint64)
    x = np.concatenate([bits.view(np.float64), np.linspace(0.9, 1.1, 1_000_001), [200.0, 1000.0, 6000.0]])
    g, h = capi.log_eval(x), _libm_log(x)
    main = (x < 0.9375) | (x >= 1.064697265625)  # the table-driven path; |x - 1| small takes the platform log (never reached: T >= 200)
    assert np.array_equal(g[main].view(np.uint64), h[main].view(np.uint64))
    assert np.max(np.abs(g[~main] - h[~main])) < 1e-15


# ---- exp and pow of the viscous block's transport fits (xfluids_b200/csrc/xf_exp.cuh; reference Visc_device.h:10-41) ----
def test_exp_pow_tables_are_this_libms_tables():
    hdr = os.path.join(REPO, "xfluids_b200", "csrc", "xf_exp_data.h")
    cur = open(hdr).read()
    try:
        subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "gen_exp_pow_table.py")], stdout=subprocess.DEVNULL)
        assert open(hdr).read() == cur, "xf_exp_data.h differs from the tables of this machine's libm"
    finally:
        open(hdr, "w").write(cur)


def test_host_restatement_of_exp_pow_bitwise_equal_to_libm():
    exe = os.path.join(tempfile.mkdtemp(prefix="xfexp_"), "check_xf_exp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fopenmp", "-mfma", "-ffp-contract=off", os.path.join(REPO, "tools", "check_xf_exp.cpp"), "-o", exe, "-lm"])
    out = subprocess.run([exe, "40000000"], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and out.stdout.startswith("mismatches=0 of "), out.stdout
    assert int(out.stdout.split()[2]) >= 100000000


def _libm_eval(which, x, y2=None):
    import xfref
    L = C.CDLL(xfref.ORACLE_SO)
    L.xo_libm_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    y = np.empty_like(x)
    L.xo_libm_eval(which, x.ctypes.data, y2.ctypes.data if y2 is not None else None, y.ctypes.data, x.size)
    return y


@pytest.mark.gpu
def test_device_exp_pow_bitwise_equal_to_libm():
    from xfluids_b200 import capi
    n = 40_000_000
    x = -40.0 + 50.0 * (np.arange(n, dtype=np.float64) + 0.5) / n          # the fits' range: ln(mu) ~ -12 ... ln(p D) ~ 3
    assert np.array_equal(capi.math_eval(1, x).view(np.uint64), _libm_eval(1, x).view(np.uint64))
    rng = np.random.default_rng(11)
    x = rng.uniform(-500.0, 500.0, 10_000_000)                             # the whole table-driven path (|x| < 512)
    assert np.array_equal(capi.math_eval(1, x).view(np.uint64), _libm_eval(1, x).view(np.uint64))
    x = np.concatenate([rng.uniform(-700.0, -512.0, 100_000), rng.uniform(512.0, 709.0, 100_000)])   # beyond it: the platform's exp (never reached by the fits)
    g, h = capi.math_eval(1, x), _libm_eval(1, x)
    assert np.max(np.abs(g - h) / h) < 1e-15
    x = 0.02 * np.exp(np.log(2500.0) * (np.arange(n, dtype=np.float64) + 0.5) / n)   # viscosity ratios: 0.02 .. 50
    y = np.full_like(x, 0.5)
    assert np.array_equal(capi.math_eval(2, x, y).view(np.uint64), _libm_eval(2, x, y).view(np.uint64))
    x = np.exp(rng.uniform(-6.9, 6.9, 10_000_000))
    y = rng.uniform(-3.0, 3.0, 10_000_000)
    assert np.array_equal(capi.math_eval(2, x, y).view(np.uint64), _libm_eval(2, x, y).view(np.uint64))
