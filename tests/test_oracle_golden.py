"""CPU: the oracle restatement (oracle/xf_oracle.cpp) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py) -- bit-exact on U, T and dt for all five BASELINE configs."""
import os

import numpy as np
import pytest

import xfref

VARIANTS = [("shock-tube", 5), ("shock-tube", 7), ("vortex", 5), ("riemann", 5), ("sbi", 5), ("sbi", 7), ("jet", 5)]


@pytest.mark.parametrize("case,weno", VARIANTS)
def test_oracle_matches_reference_golden(case, weno):
    g = np.load(os.path.join(xfref.GOLDEN, "%s_w%d.npz" % (case, weno)))
    res = tuple(int(x) for x in g["res"])
    o = xfref.Oracle(case, res, weno=weno)
    o.set_state(g["ic_U"], g["ic_T"])
    assert o.startup() == 0
    # stage-1 right-hand side of step 1
    dt = o.get_dt()
    assert dt == g["dt"][0]
    o.boundary(0); o.update_states(0); o.get_lu(0)
    assert np.array_equal(o.arr("LU"), g["s1_LU"])
    # restart and run the full steps
    o.set_state(g["ic_U"], g["ic_T"])
    o.startup()
    n, dts, t = o.run(1)
    assert n == 1 and dts[0] == g["dt"][0]
    assert np.array_equal(o.arr("U"), g["U_step1"])
    n, dts, t = o.run(9)
    assert n == 9 and np.array_equal(np.array(dts), g["dt"][1:10])
    assert np.array_equal(o.arr("U"), g["U_step10"])
    assert np.array_equal(o.arr("T"), g["T_step10"])
    assert not o.flags().any()


@pytest.mark.parametrize("case,alpha", [("vortex", 3), ("vortex", 1), ("sbi", 3), ("sbi", 1)])
def test_oracle_matches_reference_golden_other_splittings(case, alpha):
    """ROE (Artificial_type 1) and GLF (3; running maximum never reset between steps, ConVenction_block.hpp:115-170)."""
    g = np.load(os.path.join(xfref.GOLDEN, "%s_w5_%s.npz" % (case, xfref.ALPHA_NAME[alpha].lower())))
    res = tuple(int(x) for x in g["res"])
    o = xfref.Oracle(case, res, weno=5, alpha=alpha)
    o.set_state(g["ic_U"], g["ic_T"])
    assert o.startup() == 0
    n, dts, t = o.run(10)
    assert n == 10 and np.array_equal(np.array(dts), g["dt"][:10])
    assert np.array_equal(o.arr("U"), g["U_step10"])
    assert np.array_equal(o.arr("T"), g["T_step10"])


NEXT_VARIANTS = [("sbi", 6, 0), ("shock-tube", 6, 0), ("jet", 6, 0), ("sbi", 5, 1), ("sbi", 6, 1), ("shock-tube", 5, 1)]


@pytest.mark.parametrize("case,weno,pp", NEXT_VARIANTS)
def test_oracle_matches_reference_golden_cu6_and_positivity(case, weno, pp):
    """WENO-CU6 (SCHEME_ORDER 6, WENO6s_schemes.hpp:5-78) and the positivity-preserving flux limiter
    (PositivityPreserving_kernels.hpp:5-76; PP goldens run at CFL 0.9, where the limiter acts from step 1 on)."""
    g = np.load(os.path.join(xfref.GOLDEN, "%s_w%d%s.npz" % (case, weno, "_pp" if pp else "")))
    res = tuple(int(x) for x in g["res"])
    o = xfref.Oracle(case, res, weno=weno, pp=pp, cfl=float(g["cfl"]))
    o.set_state(g["ic_U"], g["ic_T"])
    assert o.startup() == 0
    dt = o.get_dt()
    assert dt == g["dt"][0]
    o.boundary(0); o.update_states(0); o.get_lu(0)
    assert np.array_equal(o.arr("FluxFw"), g["s1_Fwx"])
    assert np.array_equal(o.arr("LU"), g["s1_LU"])
    if pp:
        # the limiter must actually have acted in this fixture: the same 10 steps without it end elsewhere
        o2 = xfref.Oracle(case, res, weno=weno, pp=0, cfl=float(g["cfl"]))
        o2.set_state(g["ic_U"], g["ic_T"]); o2.startup(); o2.run(10)
        assert not np.array_equal(o2.arr("U"), g["U_step10"])
    o.set_state(g["ic_U"], g["ic_T"])
    o.startup()
    n, dts, t = o.run(10)
    assert n == 10 and np.array_equal(np.array(dts), g["dt"][:10])
    assert np.array_equal(o.arr("U"), g["U_step10"])
    assert np.array_equal(o.arr("T"), g["T_step10"])
    assert not o.flags().any()


VISC_FIXTURES = ["sbi_w5_visc", "jet_w5_visc", "sbi_w6_glf_pp_visc", "sbi_w5_visc2d"]


@pytest.mark.parametrize("name", VISC_FIXTURES)
def test_oracle_viscous_block_matches_reference_golden(name):
    """the viscous / heat-conduction / species-diffusion block of GetLU (ConVenction_block.hpp:424-575) restated in the
    oracle against the UNMODIFIED reference built with Visc, Visc_Heat and Visc_Diffu (tests/golden/make_golden_visc.py).  The oracle
    runs on this host's glibc like the reference did, so every intermediate and the state after 10 steps are bit-exact."""
    from xfluids_b200 import host
    g = np.load(os.path.join(xfref.GOLDEN, name + ".npz"))
    case = name.split("_")[0]
    res = tuple(int(x) for x in g["res"])
    weno, alpha, pp, cfl = int(g["weno"]), int(g["alpha"]), int(g["pp"]), float(g["cfl"])
    js = {"sbi": "shock-bubble.json", "jet": "expanded-jet.json"}[case]
    s = host.Setup(os.path.join(xfref.REPO, "settings", js), ["-run=%d,%d,%d" % res, "-visc=1"])
    assert np.array_equal(np.ctypeslib.as_array(s.transport.fit_visc, shape=(s.num_species * 4,)), g["fit_visc"])
    o = xfref.Oracle(case, res, weno=weno, alpha=alpha, pp=pp, cfl=cfl, transport=s.transport)
    o.set_state(g["ic_U"], g["ic_T"])
    assert o.startup() == 0
    assert o.get_dt() == g["dt"][0]
    o.boundary(0); o.update_states(0); o.get_lu(0)
    c = o.cfg
    wide = np.zeros((c.Zmax, c.Ymax, c.Xmax), bool)
    wide[c.Bw_Z - 2:c.Zmax - c.Bw_Z + 2, c.Bw_Y - 2:c.Ymax - c.Bw_Y + 2, c.Bw_X - 2:c.Xmax - c.Bw_X + 2] = True
    wide = wide.ravel()
    for m in range(9):
        assert np.array_equal(o.arr("Vde%d" % m)[wide], g["s1_Vde%d" % m][wide]), "Vde%d" % m
    for k in ("visc", "therm", "Dkm", "hi"):
        assert np.array_equal(o.arr(k), g["s1_" + k]), k
    for k, a in (("Fwx", "FluxFw"), ("Fwy", "FluxGw"), ("Fwz", "FluxHw")):
        assert np.array_equal(o.arr(a), g["s1_" + k]), k
    assert np.array_equal(o.arr("LU"), g["s1_LU"])
    # the viscous terms are not a rounding-level effect in these fixtures: the inviscid oracle gives another LU
    o2 = xfref.Oracle(case, res, weno=weno, alpha=alpha, pp=pp, cfl=cfl)
    o2.set_state(g["ic_U"], g["ic_T"]); o2.startup(); o2.boundary(0); o2.update_states(0); o2.get_lu(0)
    assert not np.array_equal(o2.arr("LU"), g["s1_LU"])
    o.set_state(g["ic_U"], g["ic_T"])
    o.startup()
    n, dts, t = o.run(10)
    assert n == 10 and np.array_equal(np.array(dts), g["dt"][:10])
    assert np.array_equal(o.arr("U"), g["U_step10"])
    assert np.array_equal(o.arr("T"), g["T_step10"])
    assert not o.flags().any()
