"""Test-side helpers (TEST INFRASTRUCTURE): ctypes binding of the CPU oracle (oracle/xf_oracle.cpp),
run/dump access to the compiled reference (oracle/_ref) and the per-case configuration table.

Nothing in xfluids_b200/ imports this module."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(REPO, "oracle", "_build", "liboracle.so")
REF_DIR = os.path.join(REPO, "oracle", "_ref")
GOLDEN = os.path.join(REPO, "tests", "golden")

# reference constant chain, include/global_setup.h:49-54
Ru = (6.02214076e26 * 1.380649e-23) * 1.0e-3


class XoCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("Xmax", "Ymax", "Zmax", "X_inner", "Y_inner", "Z_inner", "Bw_X", "Bw_Y", "Bw_Z",
                                       "DimX", "DimY", "DimZ", "NS", "Emax", "NCOP", "cop", "ghost_species", "weno", "alpha", "positivity")] + \
               [(n, C.c_double) for n in ("dx", "dy", "dz", "_dx", "_dy", "_dz", "CFL", "ncop_gamma")] + \
               [("bc", C.c_int * 6)] + [(n, C.POINTER(C.c_double)) for n in ("Hia", "Hib", "Ri", "_Wi")] + \
               [(n, C.c_int) for n in ("visc", "visc_heat", "visc_diffu")] + [(n, C.c_double) for n in ("Yil_limiter", "Dim_limiter", "dim_max0")] + \
               [(n, C.POINTER(C.c_double)) for n in ("fit_visc", "fit_therm", "fit_Dkj", "Wi_kg")]


# ---- case table: what cmake/init_sample.cmake + the north_star overrides select
CASES = {
    #  name        mixture dir                 species                         ghost  cop  bc                  cfl  domain (size, origin)
    "shock-tube": dict(mix="1d-mc-insert-shock-tube", cop=1, ghost=1, bc=[2, 2, 2, 2, 2, 2], cfl=0.4, size=(0.1, 0.1, 0.1), org=(0.0, 0.0, 0.0)),
    "vortex": dict(mix="NO-COP", cop=0, ghost=0, bc=[3] * 6, cfl=0.6, size=(10.0, 10.0, 0.0), org=(0.0, 0.0, 0.0)),
    "riemann": dict(mix="NO-COP", cop=0, ghost=0, bc=[1] * 6, cfl=0.4, size=(1.0, 1.0, 1.0), org=(0.0, 0.0, 0.0)),
    "sbi": dict(mix="Inert-SBI", cop=1, ghost=1, bc=[0, 1, 2, 1, 2, 1], cfl=0.4, size=(0.1, 0.05, 0.05), org=(-0.03, 0.0, 0.0)),
    "jet": dict(mix="2d-under-expanded-jet", cop=1, ghost=0, bc=[0, 1, 1, 1, 1, 1], cfl=0.4, size=(3.0, 1.5, 1.5), org=(0.0, -0.75, -0.75)),
}


def read_species(mix):
    """species_list.dat: names / mole ratios outside / inside (reference thermal.cpp:6-32)."""
    toks = open(os.path.join(REPO, "runtime.dat", mix, "species_list.dat")).read().split()
    if mix == "NO-COP":
        return [toks[0]], None, None
    ns = len(toks) // 3
    return toks[:ns], [float(x) for x in toks[ns:2 * ns]], [float(x) for x in toks[2 * ns:3 * ns]]


def read_thermal(names):
    """thermal_dynamics.dat -> Hia[n*21+m*3+r], Hib[n*6+m*3+r], Wi, _Wi, Ri (reference thermal.cpp:36-161)."""
    toks = open(os.path.join(REPO, "runtime.dat", "thermal_dynamics.dat")).read().split()
    ns = len(names)
    Hia, Hib, W = np.zeros(ns * 21), np.zeros(ns * 6), np.zeros(ns)
    for n, nm in enumerate(names):
        p = toks.index("*" + nm) + 1
        for r in range(3):
            for m in range(7):
                Hia[n * 21 + m * 3 + r] = float(toks[p]); p += 1
            for m in range(2):
                Hib[n * 6 + m * 3 + r] = float(toks[p]); p += 1
        W[n] = float(toks[p])
    Wi = W * 1e-3
    _Wi = 1.0 / Wi if ns and Wi.all() else np.zeros(ns)
    with np.errstate(divide="ignore"):
        _Wi = np.where(Wi != 0, 1.0 / np.where(Wi != 0, Wi, 1), np.inf)
        Ri = np.where(Wi != 0, Ru / np.where(Wi != 0, Wi, 1), np.inf)
    return Hia, Hib, Wi, _Wi, Ri


class Oracle:
    """One oracle state (all reference arrays, AoS) for one case/grid."""

    def __init__(self, case, res, weno=5, alpha=2, so=ORACLE_SO, pp=0, cfl=None, transport=None):
        """transport: a host.Setup.transport (XfTransport) to switch the viscous block on (fits and limiters are the host Setup's)."""
        if not os.path.exists(so):
            subprocess.check_call([os.path.join(REPO, "oracle", "build_oracle.sh")])
        self.lib = L = C.CDLL(so)
        L.xo_state_create.restype = C.c_void_p
        L.xo_state_create.argtypes = [C.POINTER(XoCfg)]
        L.xo_state_destroy.argtypes = [C.c_void_p]
        L.xo_array.restype = C.POINTER(C.c_double)
        L.xo_array.argtypes = [C.POINTER(XoCfg), C.c_void_p, C.c_char_p, C.POINTER(C.c_size_t)]
        L.xo_error_flags.restype = C.POINTER(C.c_int)
        L.xo_error_flags.argtypes = [C.c_void_p]
        for f in ("xo_boundary", "xo_update_states", "xo_get_lu", "xo_startup"):
            getattr(L, f).argtypes = [C.POINTER(XoCfg), C.c_void_p] + ([C.c_int] if f != "xo_startup" else [])
        L.xo_update_u.argtypes = [C.POINTER(XoCfg), C.c_void_p, C.c_double, C.c_int]
        L.xo_get_dt.restype = C.c_double
        L.xo_get_dt.argtypes = [C.POINTER(XoCfg), C.c_void_p]
        L.xo_rk_stage.argtypes = [C.POINTER(XoCfg), C.c_void_p, C.c_double, C.c_int]
        L.xo_run.argtypes = [C.POINTER(XoCfg), C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.xo_newton_stats.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]

        self.case = case
        cs = CASES[case]
        self.names, _, _ = read_species(cs["mix"])
        self.Hia, self.Hib, self.Wi, self._Wi, self.Ri = read_thermal(self.names)
        self.cfg = cfg = make_cfg(case, res, weno, alpha)
        cfg.positivity = int(pp)
        if cfl is not None:
            cfg.CFL = cfl
        for n in ("Hia", "Hib", "Ri", "_Wi"):
            setattr(cfg, n, getattr(self, n).ctypes.data_as(C.POINTER(C.c_double)))
        if transport is not None and transport.visc:
            self._tr = transport
            cfg.visc, cfg.visc_heat, cfg.visc_diffu = transport.visc, transport.visc_heat, transport.visc_diffu
            cfg.Yil_limiter, cfg.Dim_limiter, cfg.dim_max0 = transport.Yil_limiter, transport.Dim_limiter, transport.dim_max0
            cfg.fit_visc, cfg.fit_therm, cfg.fit_Dkj, cfg.Wi_kg = transport.fit_visc, transport.fit_therm, transport.fit_Dkj, transport.Wi
        self.st = L.xo_state_create(C.byref(cfg))
        self.N = cfg.Xmax * cfg.Ymax * cfg.Zmax

    def arr(self, name):
        ln = C.c_size_t()
        p = self.lib.xo_array(C.byref(self.cfg), self.st, name.encode(), C.byref(ln))
        assert p, name
        return np.ctypeslib.as_array(p, shape=(ln.value,))

    def flags(self):
        return np.ctypeslib.as_array(self.lib.xo_error_flags(self.st), shape=(4,))

    def set_state(self, U, T=None):
        self.arr("U")[:] = np.asarray(U).ravel()
        self.arr("U1")[:] = np.asarray(U).ravel()
        if T is not None:
            self.arr("T")[:] = np.asarray(T).ravel()

    def startup(self):
        return self.lib.xo_startup(C.byref(self.cfg), self.st)

    def boundary(self, which=0):
        self.lib.xo_boundary(C.byref(self.cfg), self.st, which)

    def update_states(self, which=0):
        return self.lib.xo_update_states(C.byref(self.cfg), self.st, which)

    def get_lu(self, which=0):
        self.lib.xo_get_lu(C.byref(self.cfg), self.st, which)

    def update_u(self, dt, flag):
        self.lib.xo_update_u(C.byref(self.cfg), self.st, dt, flag)

    def get_dt(self):
        return self.lib.xo_get_dt(C.byref(self.cfg), self.st)

    def rk_stage(self, dt, flag):
        return self.lib.xo_rk_stage(C.byref(self.cfg), self.st, dt, flag)

    def run(self, nsteps, t_start=0.0, t_end=1e10):
        dts = (C.c_double * max(nsteps, 1))()
        t = C.c_double()
        n = self.lib.xo_run(C.byref(self.cfg), self.st, nsteps, t_start, t_end, dts, C.byref(t))
        return n, list(dts)[:abs(n)], t.value

    def newton_stats(self, reset=True):
        a, b = C.c_int(), C.c_int()
        self.lib.xo_newton_stats(C.byref(a), C.byref(b), int(reset))
        return a.value, b.value

    def __del__(self):
        try:
            self.lib.xo_state_destroy(self.st)
        except Exception:
            pass


def make_cfg(case, res, weno=5, alpha=2, mz=1):
    """Block metrics exactly as reference iniset.cpp:290-336 (Setup::init)."""
    cs = CASES[case]
    names, _, _ = read_species(cs["mix"])
    cfg = XoCfg()
    dims = [int(bool(r)) for r in res]
    inner = [r if d else 1 for r, d in zip(res, dims)]
    bw = [4 if d else 0 for d in dims]
    size = [s if d else 1.0 for s, d in zip(cs["size"], dims)]
    cfg.X_inner, cfg.Y_inner, cfg.Z_inner = inner
    cfg.Bw_X, cfg.Bw_Y, cfg.Bw_Z = bw
    cfg.DimX, cfg.DimY, cfg.DimZ = dims
    cfg.Xmax, cfg.Ymax, cfg.Zmax = [(i + 2 * b) if d else 1 for i, b, d in zip(inner, bw, dims)]
    d = [(s / float(i)) if dd else 1.0 for s, i, dd in zip(size, inner, dims)]
    cfg.dx, cfg.dy, cfg.dz = d
    cfg._dx, cfg._dy, cfg._dz = [1.0 / x for x in d]
    cfg.CFL = cs["cfl"]
    cfg.NS = len(names)
    cfg.Emax = cfg.NS + 4 if cs["cop"] else 5
    cfg.NCOP = cfg.NS - 1 if cs["cop"] else 0
    cfg.cop, cfg.ghost_species = cs["cop"], cs["ghost"]
    cfg.weno, cfg.alpha = weno, alpha
    cfg.ncop_gamma = 1.4
    for i, b in enumerate(cs["bc"]):
        cfg.bc[i] = b
    return cfg


# ---- compiled reference (oracle/_ref) ---------------------------------------------------------------
ALPHA_NAME = {1: "ROE", 2: "LLF", 3: "GLF"}
PP_CFL = 0.9   # CFL of the positivity-preserving variants (oracle/cases/<case>_pp.json)


def ref_dir(case, weno=5, mode="parity", alpha=2, pp=0, visc=0):
    tag = "%s_w%d_%s" % (case, weno, mode)
    if alpha != 2:
        tag += "_" + ALPHA_NAME[alpha]
    if pp:
        tag += "_pp"
    if visc:
        tag += "_visc"
    return os.path.join(REF_DIR, tag)


def ref_available(case, weno=5, mode="parity", alpha=2, pp=0, visc=0):
    return os.path.exists(os.path.join(ref_dir(case, weno, mode, alpha, pp, visc), "XFLUIDS"))


def run_ref(case, res, nsteps, dump_steps=(), weno=5, mode="parity", stage_dump=False, dump_T=True, outdir=None, threads=None, alpha=2, pp=0, visc=0):
    """Run the compiled reference; returns (dict of arrays, meta dict, stdout)."""
    d = ref_dir(case, weno, mode, alpha, pp, visc)
    out = outdir or tempfile.mkdtemp(prefix="xfref_")
    env = dict(os.environ, XF_NSTEPS=str(nsteps), XF_DUMP_DIR=out, XF_DUMP_STEPS=",".join(map(str, dump_steps)),
               XF_DUMP_STAGE="1" if stage_dump else "0", XF_DUMP_T="1" if dump_T else "0")
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    if not dump_steps and not stage_dump:
        env.pop("XF_DUMP_DIR")
    for f in os.listdir(os.path.join(d, "output")):
        if "CheckingPoint" in f or "AdaptiveRange" in f:
            os.remove(os.path.join(d, "output", f))
    r = subprocess.run(["./XFLUIDS", "-run=%d,%d,%d,%d" % (res[0], res[1], res[2], nsteps)], cwd=d, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    arrays, meta = {}, {"dt": []}
    if "XF_DUMP_DIR" in env:
        for f in os.listdir(out):
            if f.endswith(".bin"):
                arrays[f[:-4]] = np.fromfile(os.path.join(out, f))
        for line in open(os.path.join(out, "meta.txt")):
            t = line.split()
            if t[0] == "dt":
                meta["dt"].append(float(t[2]))
            else:
                meta[t[0]] = float(t[1]) if "." in t[1] or "e" in t[1] else int(t[1])
    return arrays, meta, r.stdout


def rel_linf(a, b, E, inner_mask=None):
    """relative L-inf per conserved variable, normalised by max|b_n|."""
    a = np.asarray(a).reshape(-1, E)
    b = np.asarray(b).reshape(-1, E)
    if inner_mask is not None:
        a, b = a[inner_mask], b[inner_mask]
    den = np.maximum(np.abs(b).max(axis=0), 1e-300)
    return (np.abs(a - b).max(axis=0) / den).max()


def inner_mask(cfg):
    m = np.zeros((cfg.Zmax, cfg.Ymax, cfg.Xmax), bool)
    m[cfg.Bw_Z:cfg.Zmax - cfg.Bw_Z, cfg.Bw_Y:cfg.Ymax - cfg.Bw_Y, cfg.Bw_X:cfg.Xmax - cfg.Bw_X] = True
    return m.ravel()
