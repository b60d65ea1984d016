"""Test-side glue between the oracle's configuration (xfref) and the CUDA engine (xfluids_b200.Engine)."""
import ctypes as C

import numpy as np

import xfref
from xfluids_b200 import Engine, XfBlock, XfScheme, XfThermal


def make_engine(case, res, weno=5, alpha=2, fp_mode=0, device=0, pp=0, cfl=None):
    cfg = xfref.make_cfg(case, res, weno, alpha)
    if cfl is not None:
        cfg.CFL = cfl
    cs = xfref.CASES[case]
    names, _, _ = xfref.read_species(cs["mix"])
    Hia, Hib, Wi, _Wi, Ri = xfref.read_thermal(names)
    b = XfBlock()
    b.X_inner, b.Y_inner, b.Z_inner = cfg.X_inner, cfg.Y_inner, cfg.Z_inner
    b.Bwidth_X, b.Bwidth_Y, b.Bwidth_Z = cfg.Bw_X, cfg.Bw_Y, cfg.Bw_Z
    b.Xmax, b.Ymax, b.Zmax = cfg.Xmax, cfg.Ymax, cfg.Zmax
    b.DimX, b.DimY, b.DimZ = cfg.DimX, cfg.DimY, cfg.DimZ
    b.dx, b.dy, b.dz, b._dx, b._dy, b._dz = cfg.dx, cfg.dy, cfg.dz, cfg._dx, cfg._dy, cfg._dz
    b.CFLnumber = cfg.CFL
    t = XfThermal()
    t.num_species, t.cop, t.ghost_species, t.ncop_gamma = cfg.NS, cfg.cop, cfg.ghost_species, 1.4
    keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (Hia, Hib, Ri, _Wi)]
    t.Hia, t.Hib, t.Ri, t._Wi = [a.ctypes.data_as(C.POINTER(C.c_double)) for a in keep]
    s = XfScheme(weno, alpha, fp_mode, int(pp))
    eng = Engine(b, t, s, device=device, keepalive=keep)
    eng.cfg = cfg
    eng.bc = list(cs["bc"])
    return eng


def rel_linf(a, b, E):
    """max over conserved variables of max|a-b| / max|b|."""
    a = np.asarray(a).reshape(-1, E)
    b = np.asarray(b).reshape(-1, E)
    den = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return float((np.abs(a - b).max(axis=0) / den).max())


def reference_step_unfused(eng, nsteps, t_end=1e300):
    """The reference's stage sequence through the individual block entry points (XFLUIDS.cpp:441-525)."""
    t = 0.0
    dts = []
    for _ in range(nsteps):
        dt, _m = eng.get_dt()
        if t + dt > t_end:
            dt = t_end - t
        t += dt
        dts.append(dt)
        for flag in (1, 2, 3):
            UI = eng.U if flag == 1 else eng.U1
            eng.boundary(UI, eng.bc)
            assert eng.update_states(UI) == 0
            eng.get_lu(UI)
            assert eng.estimate_nan(eng.U if flag == 3 else eng.U1) == 0
            eng.update_u(dt, flag)
    return dts


def rel_linf_components(a, b, E, joint_momentum=True):
    """Per-variable max|a-b| / scale_n.  scale_n = max|b_n|, except that the three momentum components share one scale
    (max over the three): momentum is a vector, and a component that is zero by symmetry (rho*w in the planar jet, rho*v and
    rho*w in the 1-D tube) has max|b_n| ~ 1e-18 of rounding noise, against which any difference is O(1)."""
    a = np.asarray(a).reshape(-1, E)
    b = np.asarray(b).reshape(-1, E)
    den = np.abs(b).max(axis=0)
    if joint_momentum:
        den[1:4] = den[1:4].max()
    den = np.maximum(den, 1e-300)
    return np.abs(a - b).max(axis=0) / den
