"""ctypes binding of libxfluids_host.so: the C++ host layer (Setup in XFluids' JSON / runtime.dat formats, initial
conditions, the XFLUIDS driver mirror).  The library links libxfluids_b200.so; Setup and InitialCondition are pure host
code and work without a GPU, the solver entry points fail loudly without one."""
import ctypes as C
import os

import numpy as np

from .capi import XfBlock, XfScheme, XfThermal, XfTransport, XfError, Lib

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
_P = C.c_void_p

HOST_SYMBOLS = ["xfh_last_error", "xfh_setup_create", "xfh_setup_destroy", "xfh_setup_block", "xfh_setup_thermal", "xfh_setup_scheme", "xfh_setup_transport",
                "xfh_setup_bc", "xfh_setup_info", "xfh_setup_stamps", "xfh_setup_ini", "xfh_initial_condition", "xfh_solver_create",
                "xfh_solver_destroy", "xfh_solver_init", "xfh_solver_evolve", "xfh_solver_download", "xfh_solver_checkpoint",
                "xfh_solver_ctx", "xfh_solver_fields"]


def host_lib_path():
    return os.path.join(_HERE, "libxfluids_host.so")


class HostLib:
    _inst = None

    def __init__(self):
        Lib.get()  # libxfluids_b200.so first (fails loudly when missing)
        p = host_lib_path()
        if not os.path.exists(p):
            raise XfError("host library %s is not built (xfluids_b200/build.sh)" % p)
        self.dll = d = C.CDLL(p)
        d.xfh_last_error.restype = C.c_char_p
        d.xfh_setup_create.restype = _P
        d.xfh_setup_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
        d.xfh_setup_destroy.argtypes = [_P]
        d.xfh_setup_block.argtypes = [_P, C.POINTER(XfBlock)]
        d.xfh_setup_thermal.argtypes = [_P, C.POINTER(XfThermal)]
        d.xfh_setup_scheme.argtypes = [_P, C.POINTER(XfScheme)]
        d.xfh_setup_transport.argtypes = [_P, C.POINTER(XfTransport)]
        d.xfh_setup_bc.argtypes = [_P, C.c_int * 6]
        d.xfh_setup_info.argtypes = [_P, C.c_int * 8]
        d.xfh_setup_stamps.argtypes = [_P, _P]
        d.xfh_setup_ini.argtypes = [_P, C.c_double * 10]
        d.xfh_initial_condition.argtypes = [_P, _P, _P]
        d.xfh_solver_create.restype = _P
        d.xfh_solver_create.argtypes = [_P, C.c_int]
        d.xfh_solver_destroy.argtypes = [_P]
        d.xfh_solver_init.argtypes = [_P]
        d.xfh_solver_evolve.argtypes = [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        d.xfh_solver_download.argtypes = [_P, _P]
        d.xfh_solver_checkpoint.argtypes = [_P, C.c_char_p]
        d.xfh_solver_ctx.restype = _P
        d.xfh_solver_ctx.argtypes = [_P]
        d.xfh_solver_fields.argtypes = [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def err(self):
        return (self.dll.xfh_last_error() or b"").decode()


class Setup:
    """Reference struct Setup: settings JSON + CLI overrides (-run=, -mpi=, -mpi-s=, -sample= ...) + runtime.dat tables."""

    def __init__(self, json_path, cli=(), workdir=REPO, rank=0, nranks=1):
        self.H = HostLib.get()
        argv = (C.c_char_p * max(len(cli), 1))(*[a.encode() for a in cli])
        self.h = self.H.dll.xfh_setup_create(str(json_path).encode(), str(workdir).encode(), len(cli), argv, rank, nranks)
        if not self.h:
            raise XfError("Setup: " + self.H.err())
        self.block, self.thermal, self.scheme = XfBlock(), XfThermal(), XfScheme()
        self.H.dll.xfh_setup_block(self.h, C.byref(self.block))
        self.H.dll.xfh_setup_thermal(self.h, C.byref(self.thermal))
        self.H.dll.xfh_setup_scheme(self.h, C.byref(self.scheme))
        self.transport = XfTransport()
        self.H.dll.xfh_setup_transport(self.h, C.byref(self.transport))
        bc, info = (C.c_int * 6)(), (C.c_int * 8)()
        self.H.dll.xfh_setup_bc(self.h, bc)
        self.H.dll.xfh_setup_info(self.h, info)
        self.bc = list(bc)
        self.Emax, self.num_species, self.cop, self.ghost_species, self.nStepmax, self.mz, self.myMpiPos_z, ns = list(info)
        st = np.zeros(max(ns, 1))
        self.H.dll.xfh_setup_stamps(self.h, st.ctypes.data_as(_P))
        self.stamps = st[:ns].tolist()
        self.ncells = self.block.Xmax * self.block.Ymax * self.block.Zmax

    def ini(self):
        v = (C.c_double * 10)()
        self.H.dll.xfh_setup_ini(self.h, v)
        return list(v)

    def initial_condition(self, U=None, T=None):
        """InitializeFluidStates on the host: returns (U_aos [ncells*Emax], T [ncells]); U / T may be caller-provided
        (e.g. a view of pinned memory) so that large blocks are written in place."""
        U = np.empty(self.ncells * self.Emax) if U is None else U
        T = np.empty(self.ncells) if T is None else T
        assert U.size == self.ncells * self.Emax and T.size == self.ncells and U.dtype == np.float64 and T.dtype == np.float64
        rc = self.H.dll.xfh_initial_condition(self.h, U.ctypes.data_as(_P), T.ctypes.data_as(_P))
        if rc:
            raise XfError("unknown or inconsistent sample/mixture (rc %d)" % rc)
        return U, T

    def __del__(self):
        try:
            self.H.dll.xfh_setup_destroy(self.h)
        except Exception:
            pass


class Solver:
    """Reference class XFLUIDS on one GPU: create -> init (IC, BC, UpdateStates) -> evolve."""

    def __init__(self, setup, device=0):
        self.S, self.H = setup, setup.H
        self.h = self.H.dll.xfh_solver_create(setup.h, device)
        if not self.h:
            raise XfError("Solver: " + self.H.err())

    def init(self):
        rc = self.H.dll.xfh_solver_init(self.h)
        if rc < 0:
            raise XfError(self.H.err())
        return rc

    def evolve(self, fused=True):
        it, t, s = C.c_int(), C.c_double(), C.c_double()
        rc = self.H.dll.xfh_solver_evolve(self.h, int(fused), C.byref(it), C.byref(t), C.byref(s))
        if rc < 0:
            raise XfError(self.H.err())
        return dict(error=rc, iteration=it.value, time=t.value, seconds=s.value)

    def download(self):
        U = np.empty(self.S.ncells * self.S.Emax)
        if self.H.dll.xfh_solver_download(self.h, U.ctypes.data_as(_P)):
            raise XfError(self.H.err())
        return U

    def checkpoint(self, path):
        if self.H.dll.xfh_solver_checkpoint(self.h, str(path).encode()):
            raise XfError(self.H.err())

    def close(self):
        if self.h:
            self.H.dll.xfh_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
