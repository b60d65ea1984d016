"""Viscous / heat-conduction / species-diffusion terms against the UNMODIFIED reference built with Visc, Visc_Heat and
Visc_Diffu on (tests/golden/*_visc.npz, written by tests/golden/make_golden_visc.py from oracle/_ref).

CPU (not gpu): the host transport fits (xfluids_b200/host/xfh_transport.cpp: Lennard-Jones data, collision integrals, least squares) equal
the reference's arrays BIT FOR BIT -- including the reference's out-of-range table row (see that file).
GPU: the intermediates of the viscous block of stage 1 and the state after 1 and 10 steps, BIT FOR BIT: log, exp and pow replay glibc's
algorithms on the device (csrc/xf_log.cuh, csrc/xf_exp.cuh), everything else is IEEE basic operations in the reference's order."""
import ctypes as C
import os

import numpy as np
import pytest

import xfref

FIXTURES = ["sbi_w5_visc", "jet_w5_visc", "sbi_w6_glf_pp_visc", "sbi_w5_visc2d"]
SETTINGS = {"sbi": "shock-bubble.json", "jet": "expanded-jet.json"}


def load(name):
    g = np.load(os.path.join(xfref.GOLDEN, name + ".npz"))
    case = name.split("_")[0]
    res = tuple(int(x) for x in g["res"])
    cli = ["-run=%d,%d,%d" % res, "-weno=%d" % int(g["weno"]), "-alpha=" + xfref.ALPHA_NAME[int(g["alpha"])], "-pp=%d" % int(g["pp"]), "-cfl=%.17g" % float(g["cfl"]), "-visc=1"]
    from xfluids_b200 import host
    s = host.Setup(os.path.join(xfref.REPO, "settings", SETTINGS[case]), cli)
    return g, s


@pytest.mark.parametrize("name", FIXTURES)
def test_host_transport_fits_equal_reference_bitwise(name):
    g, s = load(name)
    ns, tr = s.num_species, s.transport
    assert (tr.visc, tr.visc_heat, tr.visc_diffu) == (1, 1, 1)
    assert np.array_equal(np.ctypeslib.as_array(tr.fit_visc, shape=(ns * 4,)), g["fit_visc"])
    assert np.array_equal(np.ctypeslib.as_array(tr.fit_therm, shape=(ns * 4,)), g["fit_therm"])
    assert np.array_equal(np.ctypeslib.as_array(tr.fit_Dkj, shape=(ns * ns * 4,)), g["fit_Dkj"])


def engine(s):
    from xfluids_b200 import capi
    eng = capi.Engine(s.block, s.thermal, s.scheme, device=0, keepalive=(s,))
    eng.set_transport(s.transport, keepalive=(s,))
    return eng


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_viscous_block_stage1_vs_reference(name):
    g, s = load(name)
    eng = engine(s)
    E, ns, b = eng.E, s.num_species, s.block
    eng.set_state(g["ic_U"], g["ic_T"])
    eng.boundary(eng.U, s.bc)
    assert eng.update_states(eng.U) == 0
    dt, _ = eng.get_dt()
    assert dt == float(g["dt"][0])
    eng.boundary(eng.U, s.bc)
    assert eng.update_states(eng.U) == 0
    eng.get_lu(eng.U)
    mask = np.zeros((b.Zmax, b.Ymax, b.Xmax), bool)
    mask[b.Bwidth_Z:b.Zmax - b.Bwidth_Z, b.Bwidth_Y:b.Ymax - b.Bwidth_Y, b.Bwidth_X:b.Xmax - b.Bwidth_X] = True
    mask = mask.ravel()
    # cells the viscous fluxes read: inner + 2 in every direction
    wide = np.zeros((b.Zmax, b.Ymax, b.Xmax), bool)
    wide[b.Bwidth_Z - 2:b.Zmax - b.Bwidth_Z + 2, b.Bwidth_Y - 2:b.Ymax - b.Bwidth_Y + 2, b.Bwidth_X - 2:b.Xmax - b.Bwidth_X + 2] = True
    wide = wide.ravel()
    for m in range(9):   # velocity derivatives after their ghost fill: IEEE basic operations only
        assert np.array_equal(eng.get_scalar("Vde%d" % m)[wide], g["s1_Vde%d" % m][wide]), "Vde%d" % m
    hi, Dkm = g["s1_hi"].reshape(-1, ns), g["s1_Dkm"].reshape(-1, ns)
    for k in range(ns):
        assert np.array_equal(eng.get_scalar("hi%d" % k), hi[:, k]), "hi%d" % k          # log() is bit-exact
        assert np.array_equal(eng.get_scalar("Dkm%d" % k), Dkm[:, k]), "Dkm%d" % k        # exp(): glibc's sequence
    assert np.array_equal(eng.get_scalar("visc"), g["s1_visc"])                            # exp(), pow()
    assert np.array_equal(eng.get_scalar("therm"), g["s1_therm"])
    for d_, nm in enumerate(("s1_Fwx", "s1_Fwy", "s1_Fwz")):
        if not (b.DimX, b.DimY, b.DimZ)[d_]:
            continue                                                                       # 2-D fixture: no z wall flux
        a, r = eng.wallflux(d_).reshape(-1, E), g[nm].reshape(-1, E)
        w = np.abs(r).sum(axis=1) > 0
        assert np.array_equal(a[w], r[w]), (nm, rel(a[w], r[w]))
    assert np.array_equal(eng.download(eng.LU).reshape(-1, E)[mask], g["s1_LU"].reshape(-1, E)[mask])
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_viscous_steps_vs_reference_golden(name):
    import xfgpu
    g, s = load(name)
    eng = engine(s)
    E, b = eng.E, s.block
    mask = np.zeros((b.Zmax, b.Ymax, b.Xmax), bool)
    mask[b.Bwidth_Z:b.Zmax - b.Bwidth_Z, b.Bwidth_Y:b.Ymax - b.Bwidth_Y, b.Bwidth_X:b.Xmax - b.Bwidth_X] = True
    mask = mask.ravel()
    eng.set_state(g["ic_U"], g["ic_T"])
    eng.boundary(eng.U, s.bc)
    assert eng.update_states(eng.U) == 0
    # (the goldens were written with a single far-away output stamp, oracle/cases/*.json: no dt clipping)
    done, t, err = eng.run(s.bc, 1)
    assert (done, err) == (1, 0)
    e1 = xfgpu.rel_linf(eng.download(eng.U).reshape(-1, E)[mask], g["U_step1"].reshape(-1, E)[mask], E)
    done, t, err = eng.run(s.bc, 9)
    assert (done, err) == (9, 0)
    e10 = xfgpu.rel_linf(eng.download(eng.U).reshape(-1, E)[mask], g["U_step10"].reshape(-1, E)[mask], E)
    print("\n%s: rel Linf step1 %.3e step10 %.3e, t %.9e (ref %.9e)" % (name, e1, e10, t, float(np.sum(g["dt"][:10]))))
    assert e1 == 0.0 and e10 == 0.0
    assert np.array_equal(eng.download(eng.U).reshape(-1, E)[mask], g["U_step10"].reshape(-1, E)[mask])
    tref = 0.0
    for d in g["dt"][:10]:
        tref += float(d)
    assert t == tref
    eng.close()


@pytest.mark.gpu
def test_species_diffusion_with_the_mpi_builds_limiter_equals_oracle():
    """With Dim_max = 0 (the reference's single-process build, which the goldens come from) the diffusion limiter clamps every species-diffusion
    flux to 0, so the fixtures above never see a non-zero one.  The reference's MPI build starts Dim_max at 1 (`-diffu-mpi=1`): that build cannot be
    made here (no MPI), so this case is pinned against the oracle restatement only (same block of ConVenction_block.hpp:458-575, other branch):
    5 steps, bit for bit, and the species-diffusion fluxes must have acted."""
    import xfgpu
    g = np.load(os.path.join(xfref.GOLDEN, "jet_w5_visc.npz"))   # (the jet: the shock-bubble preset sets Yil_limiter = 0, which switches the diffusion off by itself)
    res = tuple(int(x) for x in g["res"])
    from xfluids_b200 import host
    out = {}
    for mpi in (0, 1):
        s = host.Setup(os.path.join(xfref.REPO, "settings", SETTINGS["jet"]), ["-run=%d,%d,%d" % res, "-weno=5", "-alpha=LLF", "-visc=1", "-diffu-mpi=%d" % mpi])
        assert (s.transport.dim_max0 != 0.0) == bool(mpi)
        eng = engine(s)
        eng.set_state(g["ic_U"], g["ic_T"])
        eng.boundary(eng.U, s.bc)
        assert eng.update_states(eng.U) == 0
        done, t, err = eng.run(s.bc, 5)
        assert (done, err) == (5, 0)
        out[mpi] = eng.download(eng.U)
        eng.close()
        if mpi:
            o = xfref.Oracle("jet", res, weno=5, alpha=2, transport=s.transport)
            o.set_state(g["ic_U"], g["ic_T"])
            assert o.startup() == 0
            n, dts, t_o = o.run(5)
            assert n == 5 and t_o == t
            assert np.array_equal(out[1], o.arr("U"))
    assert not np.array_equal(out[0], out[1]), "the species-diffusion fluxes did not act"
