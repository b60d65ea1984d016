"""Correctness pinned at (or near) BASELINE scale against the UNMODIFIED reference (tests/golden/scale_pins.json, written by
tests/golden/make_scale_pins.py from oracle/_ref): SHA-256 of the reference's conserved state after 3 steps on Riemann 1024^2,
vortex 1024^2, SBI 128x64x64, jet 128x64x64 and a 4000-cell shock tube -- sizes where the kernels' 32-bit in-plane index arithmetic,
several x chunks and y/z tiles per pencil, the 8x8 block order of the update and the Newton hard-cell list are all exercised
(the 24x12x12 / 32^2 fixtures touch one tile each).  Bit-exact or fail.

CPU (not gpu): the C++ initial-condition hooks reproduce the reference's ic_U / ic_T at these sizes.
GPU: 3 steps through the fused CUDA path reproduce U, T and the dt sequence."""
import hashlib
import json
import os

import numpy as np
import pytest

import xfref

PINS = json.load(open(os.path.join(xfref.GOLDEN, "scale_pins.json")))
SETTINGS = {"shock-tube": "1d-shock-tube", "vortex": "2d-euler-vortex", "riemann": "2d-riemann", "sbi": "shock-bubble", "jet": "expanded-jet"}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def setup_of(pin):
    from xfluids_b200 import host
    return host.Setup(os.path.join(xfref.REPO, "settings", SETTINGS[pin["case"]] + ".json"), ["-run=%d,%d,%d" % tuple(pin["res"]), "-weno=%d" % pin["weno"], "-alpha=LLF", "-pp=0"])


@pytest.mark.parametrize("key", sorted(PINS["pins"]))
def test_host_initial_condition_matches_reference_at_scale(key):
    pin = PINS["pins"][key]
    s = setup_of(pin)
    assert s.ncells == pin["ncells"] and s.Emax == pin["emax"]
    U, T = s.initial_condition()
    assert sha(U) == pin["ic_U_sha256"]
    assert sha(T) == pin["ic_T_sha256"]


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(PINS["pins"]))
def test_three_steps_hash_equals_reference_at_scale(key):
    from xfluids_b200 import capi
    pin = PINS["pins"][key]
    s = setup_of(pin)
    U0, T0 = s.initial_condition()
    eng = capi.Engine(s.block, s.thermal, s.scheme, device=0, keepalive=(s,))
    eng.set_state(U0, T0)
    eng.boundary(eng.U, s.bc)
    assert eng.update_states(eng.U) == 0
    dts, t_prev = [], 0.0
    for _ in range(PINS["nsteps"]):
        done, t, err = eng.run(s.bc, 1)
        assert (done, err) == (1, 0)
        dts.append(eng.time()[1])
        t_prev = t
    assert [float(x).hex() for x in dts] == pin["dt_hex"]
    U = eng.download(eng.U)
    E = eng.E
    b = s.block
    rho = U.reshape(b.Zmax, b.Ymax, b.Xmax, E)[..., 0]
    sums = [float(x).hex() for x in rho.reshape(b.Zmax, -1).sum(axis=1)][:8]
    assert sums == pin["rho_plane_sums_hex"], "per-plane sums of rho differ (first planes)"
    assert sha(U) == pin["U_sha256"]
    if s.cop:
        assert sha(eng.get_scalar("T")) == pin["T_sha256"]
    eng.close()
