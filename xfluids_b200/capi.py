"""ctypes binding of include/xfluids_b200.h.  Fails loudly when the CUDA library is not built."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    # XF_LIB: an alternative build of the same library (kernel tuning experiments); default: the in-tree build
    return os.environ.get("XF_LIB") or os.path.join(_HERE, "libxfluids_b200.so")


class XfError(RuntimeError):
    pass


class XfBlock(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("X_inner", "Y_inner", "Z_inner", "Bwidth_X", "Bwidth_Y", "Bwidth_Z", "Xmax", "Ymax", "Zmax",
                                       "DimX", "DimY", "DimZ")] + \
               [(n, C.c_double) for n in ("dx", "dy", "dz", "_dx", "_dy", "_dz", "CFLnumber")]


class XfThermal(C.Structure):
    _fields_ = [("num_species", C.c_int), ("cop", C.c_int), ("ghost_species", C.c_int), ("ncop_gamma", C.c_double),
                ("Hia", C.POINTER(C.c_double)), ("Hib", C.POINTER(C.c_double)), ("Ri", C.POINTER(C.c_double)), ("_Wi", C.POINTER(C.c_double))]


class XfTransport(C.Structure):
    _fields_ = [("visc", C.c_int), ("visc_heat", C.c_int), ("visc_diffu", C.c_int), ("fit_visc", C.POINTER(C.c_double)), ("fit_therm", C.POINTER(C.c_double)),
                ("fit_Dkj", C.POINTER(C.c_double)), ("Wi", C.POINTER(C.c_double)), ("Yil_limiter", C.c_double), ("Dim_limiter", C.c_double), ("dim_max0", C.c_double)]


class XfScheme(C.Structure):
    _fields_ = [("weno_order", C.c_int), ("artificial_type", C.c_int), ("fp_mode", C.c_int), ("positivity", C.c_int)]


_P = C.c_void_p
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)
_BC = C.c_int * 6

_PROTOS = {
    "xf_create": (C.c_int, [C.POINTER(XfBlock), C.POINTER(XfThermal), C.POINTER(XfScheme), C.c_int, C.POINTER(_P)]),
    "xf_destroy": (C.c_int, [_P]),
    "xf_last_error": (C.c_char_p, []),
    "xf_set_transport": (C.c_int, [_P, C.POINTER(XfTransport)]),
    "xf_transport_needs_global_extrema": (C.c_int, [_P]),
    "xf_set_stream": (C.c_int, [_P, _P]),
    "xf_synchronize": (C.c_int, [_P]),
    "xf_pitch": (C.c_size_t, [_P]),
    "xf_field_stride": (C.c_size_t, [_P]),
    "xf_field_doubles": (C.c_size_t, [_P]),
    "xf_emax": (C.c_int, [_P]),
    "xf_field_alloc": (C.c_int, [_P, C.POINTER(_P)]),
    "xf_field_free": (C.c_int, [_P, _P]),
    "xf_upload_aos": (C.c_int, [_P, _P, _P]),
    "xf_download_aos": (C.c_int, [_P, _P, _P]),
    "xf_set_scalar": (C.c_int, [_P, C.c_char_p, _P]),
    "xf_get_scalar": (C.c_int, [_P, C.c_char_p, _P]),
    "xf_get_wallflux_aos": (C.c_int, [_P, C.c_int, _P]),
    "xf_boundary": (C.c_int, [_P, _P, _BC]),
    "xf_update_states": (C.c_int, [_P, _P, _IP]),
    "xf_get_lu": (C.c_int, [_P, _P, _P]),
    "xf_estimate_nan": (C.c_int, [_P, _P, _P, _IP]),
    "xf_update_u_rk3": (C.c_int, [_P, _P, _P, _P, C.c_double, C.c_int]),
    "xf_get_dt": (C.c_int, [_P, _DP, _DP]),
    "xf_rk_stage": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_int), C.c_int]),
    "xf_dt_device": (C.c_int, [_P, C.c_double]),
    "xf_run": (C.c_int, [_P, _P, _P, _P, _BC, C.c_int, C.c_double, _IP, _DP, _IP]),
    "xf_get_time": (C.c_int, [_P, _DP, _DP]),
    "xf_set_time": (C.c_int, [_P, C.c_double]),
    "xf_error_flags": (C.c_int, [_P, C.c_int * 4]),
    "xf_clear_errors": (C.c_int, [_P]),
    "xf_device_dtmax": (_P, [_P]),
    "xf_device_errors": (_P, [_P]),
    "xf_device_glfmax": (_P, [_P]),
    "xf_stage_states": (C.c_int, [_P, _P, _P, C.c_int]),
    "xf_stage_fluxes": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "xf_halo_doubles": (C.c_size_t, [_P]),
    "xf_halo_pack": (C.c_int, [_P, _P, C.c_int, _P]),
    "xf_halo_unpack": (C.c_int, [_P, _P, C.c_int, _P]),
    "xf_halo_pack_on": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "xf_halo_unpack_on": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "xf_stage_interior": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "xf_stage_finish": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "xf_step_host": (C.c_int, [_P, _P, _BC, C.c_int, C.c_double, _P, _P, _P, _IP, _IP]),
    "xf_set_host_overlap": (C.c_int, [_P, C.c_int]),
    "xf_host_begin": (C.c_int, [_P, _P, _BC, C.c_double, _P, _P, _P]),
    "xf_host_stage1_finish": (C.c_int, [_P, _BC, _P, _P, _P]),
    "xf_host_stage3": (C.c_int, [_P, _P, _P, _P, _P, _IP]),
    "xf_host_alloc_pinned": (_P, [C.c_size_t]),
    "xf_host_free_pinned": (None, [_P]),
    "xf_launch_count": (C.c_longlong, [_P]),
    "xf_profile_step": (C.c_int, [_P, _P, _P, _P, _BC, C.c_double, C.c_float * 8]),
    "xf_measure_peaks": (C.c_int, [C.c_int, _DP, _DP]),
    "xf_log_eval": (C.c_int, [C.c_int, _P, _P, C.c_size_t]),
    "xf_math_eval": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, C.c_size_t]),
    "xf_measure_pcie": (C.c_int, [_P, _P, C.c_size_t, _DP, _DP]),
    "xf_measure_fp64_issue": (C.c_int, [C.c_int, C.c_double * 3]),
    "xf_slab_last_error": (C.c_char_p, []),
    "xf_comm_unique_id": (C.c_int, [C.c_char * 128]),
    "xf_comm_create": (C.c_int, [C.c_char * 128, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "xf_comm_destroy": (C.c_int, [_P]),
    "xf_comm_rank": (C.c_int, [_P]),
    "xf_comm_world": (C.c_int, [_P]),
    "xf_comm_allreduce_max": (C.c_int, [_P, _P, C.c_int, _P]),
    "xf_comm_allreduce_max_int": (C.c_int, [_P, _P, C.c_int, _P]),
    "xf_slab_create": (C.c_int, [_P, _P, _BC, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "xf_slab_destroy": (C.c_int, [_P]),
    "xf_slab_set_overlap": (C.c_int, [_P, C.c_int]),
    "xf_slab_neighbours": (C.c_int, [_P, C.c_int * 2]),
    "xf_slab_halo": (C.c_int, [_P, _P]),
    "xf_slab_startup": (C.c_int, [_P, _P, _IP]),
    "xf_slab_stage": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "xf_slab_step": (C.c_int, [_P, _P, _P, _P, C.c_double]),
    "xf_slab_run": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_double, _IP, _DP, _IP]),
    "xf_slab_any_error": (C.c_int, [_P, _IP]),
    "xf_slab_step_host": (C.c_int, [_P, _P, C.c_double, _P, _P, _P, _IP]),
    "xf_slab_allreduce_max_host": (C.c_int, [_P, _DP, C.c_int]),
}

EXPORTED_SYMBOLS = sorted(_PROTOS)


class Lib:
    """The loaded shared library with typed prototypes."""
    _inst = None

    def __init__(self, path=None):
        path = path or lib_path()
        if not os.path.exists(path):
            raise XfError("CUDA library %s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or xfluids_b200/build.sh); xfluids_b200 has no CPU fallback" % path)
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in _PROTOS.items():
            f = getattr(self.dll, name)
            f.restype, f.argtypes = res, args

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def check(self, rc, allow_numeric=False):
        if rc == 0 or (allow_numeric and rc == -3):
            return rc
        msg = (self.dll.xf_slab_last_error() if rc == -4 else self.dll.xf_last_error()) or b""
        raise XfError("xfluids_b200 error %d: %s" % (rc, msg.decode() or (self.dll.xf_slab_last_error() or b"").decode()))


def _dptr(a):
    return a.ctypes.data_as(_P)


def measure_peaks(device=0):
    """(FP64 FMA TFLOP/s, copy GB/s) measured on the device by the library's micro-benchmarks."""
    L = Lib.get()
    a, b = C.c_double(), C.c_double()
    L.check(L.dll.xf_measure_peaks(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def measure_fp64_issue(device=0):
    """(DADD, DMUL, DFMA) issue rates in 1e12 thread-instructions per second."""
    L = Lib.get()
    v = (C.c_double * 3)()
    L.check(L.dll.xf_measure_fp64_issue(device, v))
    return tuple(v)


def log_eval(x, device=0):
    """The device logarithm (csrc/xf_log.cuh) of a host array, evaluated on `device`."""
    L = Lib.get()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    L.check(L.dll.xf_log_eval(device, _dptr(x), _dptr(y), x.size))
    return y


def math_eval(which, x, y2=None, device=0):
    """Device log (0) / exp (1) / pow (2; exponents y2) of host arrays (csrc/xf_log.cuh, csrc/xf_exp.cuh), evaluated on `device`."""
    L = Lib.get()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    if y2 is not None:
        y2 = np.ascontiguousarray(np.broadcast_to(np.asarray(y2, dtype=np.float64), x.shape))
    L.check(L.dll.xf_math_eval(device, which, _dptr(x), _dptr(y2) if y2 is not None else None, _dptr(out), x.size))
    return out


class Engine:
    """One fluid block on one GPU: owns a xf_ctx plus the three conserved fields (reference Fluid::d_U, d_U1, d_LU).

    Method names follow the reference's Fluid / block functions."""

    def __init__(self, block, thermal, scheme, device=0, keepalive=()):
        self.L = Lib.get()
        self._keep = (block, thermal, scheme) + tuple(keepalive)
        self.block, self.scheme = block, scheme
        self.ctx = _P()
        self.L.check(self.L.dll.xf_create(C.byref(block), C.byref(thermal), C.byref(scheme), device, C.byref(self.ctx)))
        self.E = self.L.dll.xf_emax(self.ctx)
        self.ncells = block.Xmax * block.Ymax * block.Zmax
        self.U, self.U1, self.LU = _P(), _P(), _P()
        for f in (self.U, self.U1, self.LU):
            self.L.check(self.L.dll.xf_field_alloc(self.ctx, C.byref(f)))

    def set_transport(self, tr, keepalive=()):
        """Viscous / heat-conduction / species-diffusion terms on (xf_set_transport); tr: XfTransport (host.Setup.transport)."""
        self._keep = self._keep + (tr,) + tuple(keepalive)
        self.L.check(self.L.dll.xf_set_transport(self.ctx, C.byref(tr)))

    # ---- state I/O in the reference's AoS layout -------------------------------------------------
    def upload(self, field, aos):
        aos = np.ascontiguousarray(aos, dtype=np.float64).ravel()
        assert aos.size == self.ncells * self.E
        self.L.check(self.L.dll.xf_upload_aos(self.ctx, field, _dptr(aos)))

    def download(self, field):
        out = np.empty(self.ncells * self.E)
        self.L.check(self.L.dll.xf_download_aos(self.ctx, field, _dptr(out)))
        return out

    def set_state(self, U_aos, T=None):
        """InitialCondition: U and U1 = U (every sample's InitialUFKernel does that), T = Newton warm start."""
        self.upload(self.U, U_aos)
        self.upload(self.U1, U_aos)
        if T is not None:
            self.set_scalar("T", T)

    def set_scalar(self, name, h):
        h = np.ascontiguousarray(h, dtype=np.float64).ravel()
        assert h.size == self.ncells
        self.L.check(self.L.dll.xf_set_scalar(self.ctx, name.encode(), _dptr(h)))

    def get_scalar(self, name):
        out = np.empty(self.ncells)
        self.L.check(self.L.dll.xf_get_scalar(self.ctx, name.encode(), _dptr(out)))
        return out

    def wallflux(self, d):
        out = np.empty(self.ncells * self.E)
        self.L.check(self.L.dll.xf_get_wallflux_aos(self.ctx, d, _dptr(out)))
        return out

    # ---- block functions ------------------------------------------------------------------------------
    def boundary(self, field, bc):
        self.L.check(self.L.dll.xf_boundary(self.ctx, field, _BC(*bc)))

    def update_states(self, field):
        err = C.c_int()
        self.L.check(self.L.dll.xf_update_states(self.ctx, field, C.byref(err)))
        return err.value

    def get_lu(self, field):
        self.L.check(self.L.dll.xf_get_lu(self.ctx, field, self.LU))

    def estimate_nan(self, field):
        err = C.c_int()
        self.L.check(self.L.dll.xf_estimate_nan(self.ctx, field, self.LU, C.byref(err)))
        return err.value

    def update_u(self, dt, flag):
        self.L.check(self.L.dll.xf_update_u_rk3(self.ctx, self.U, self.U1, self.LU, dt, flag))

    def get_dt(self):
        dt = C.c_double()
        m = (C.c_double * 3)()
        self.L.check(self.L.dll.xf_get_dt(self.ctx, C.byref(dt), m))
        return dt.value, list(m)

    # ---- fused path -------------------------------------------------------------------------------------
    def rk_stage(self, bc, flag):
        b = _BC(*bc) if bc is not None else None
        self.L.check(self.L.dll.xf_rk_stage(self.ctx, self.U, self.U1, self.LU, b, flag))

    def stage_interior(self, flag):
        self.L.check(self.L.dll.xf_stage_interior(self.ctx, self.U, self.U1, self.LU, flag))

    def stage_finish(self, flag):
        self.L.check(self.L.dll.xf_stage_finish(self.ctx, self.U, self.U1, self.LU, flag))

    def dt_device(self, t_end=1e300):
        self.L.check(self.L.dll.xf_dt_device(self.ctx, t_end))

    def run(self, bc, nsteps, t_end=1e300):
        done, err, t = C.c_int(), C.c_int(), C.c_double()
        rc = self.L.dll.xf_run(self.ctx, self.U, self.U1, self.LU, _BC(*bc), nsteps, t_end, C.byref(done), C.byref(t), C.byref(err))
        self.L.check(rc, allow_numeric=True)
        return done.value, t.value, err.value

    def time(self):
        t, dt = C.c_double(), C.c_double()
        self.L.check(self.L.dll.xf_get_time(self.ctx, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def set_time(self, t):
        self.L.check(self.L.dll.xf_set_time(self.ctx, t))

    def error_flags(self):
        f = (C.c_int * 4)()
        self.L.check(self.L.dll.xf_error_flags(self.ctx, f))
        return list(f)

    def profile_step(self, bc, t_end=1e300):
        """One eager step with per-kernel CUDA-event timing: dict of ms (summed over the 3 stages)."""
        ms = (C.c_float * 8)()
        self.L.check(self.L.dll.xf_profile_step(self.ctx, self.U, self.U1, self.LU, _BC(*bc), t_end, ms))
        return dict(zip(("dt", "bc", "prim", "sweep_x", "sweep_y", "sweep_z", "lu_rk", "step"), [float(x) for x in ms]))

    def set_stream(self, cuda_stream):
        self.L.check(self.L.dll.xf_set_stream(self.ctx, _P(cuda_stream)))

    def step_host(self, h_ptr, bc, nsteps, t_end=1e300):
        """xf_step_host: upload AoS U from (pinned) host memory, run nsteps, download AoS U into the same buffer."""
        done, err = C.c_int(), C.c_int()
        rc = self.L.dll.xf_step_host(self.ctx, _P(h_ptr), _BC(*bc), nsteps, t_end, self.U, self.U1, self.LU, C.byref(done), C.byref(err))
        self.L.check(rc, allow_numeric=True)
        return done.value, err.value

    def launches(self):
        return self.L.dll.xf_launch_count(self.ctx)

    def sync(self):
        self.L.check(self.L.dll.xf_synchronize(self.ctx))

    def close(self):
        if self.ctx:
            for f in (self.U, self.U1, self.LU):
                self.L.dll.xf_field_free(self.ctx, f)
            self.L.dll.xf_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
