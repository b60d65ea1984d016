"""xfluids_b200 -- B200-native (sm_100a CUDA, FP64) implementation of the XFluids inviscid per-RK-stage RHS path.

The compute lives in libxfluids_b200.so (hand-written CUDA kernels behind the C ABI of include/xfluids_b200.h);
this package is the thin Python binding used by the tests and bench.py, plus the host-side mirror of the reference's
Setup / Fluid / XFLUIDS driver flow (C++ in xfluids_b200/host, bound through libxfluids_host.so).
There is no CPU fallback: importing the binding without the built CUDA library raises."""
from .capi import Lib, Engine, XfBlock, XfThermal, XfScheme, XfError, lib_path  # noqa: F401

__all__ = ["Lib", "Engine", "XfBlock", "XfThermal", "XfScheme", "XfError", "lib_path"]
