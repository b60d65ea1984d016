"""CPU: the C++ host layer (JSON / runtime.dat readers, Block metrics, Mach_Shock, the five initial-condition hooks)
against the golden vectors written by the unmodified reference -- bit-exact -- plus the library/ABI surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import xfref
from xfluids_b200 import capi, host

REPO = xfref.REPO
SETTINGS = {"shock-tube": "1d-shock-tube", "vortex": "2d-euler-vortex", "riemann": "2d-riemann", "sbi": "shock-bubble", "jet": "expanded-jet"}


def setup_for(case, res, weno=5, extra=()):
    return host.Setup(os.path.join(REPO, "settings", SETTINGS[case] + ".json"), ["-run=%d,%d,%d" % tuple(res), "-weno=%d" % weno] + list(extra))


@pytest.mark.parametrize("case", ["shock-tube", "vortex", "riemann", "sbi", "jet"])
def test_initial_condition_bit_exact_vs_reference(case):
    g = np.load(os.path.join(xfref.GOLDEN, "%s_w5.npz" % case))
    res = tuple(int(x) for x in g["res"])
    s = setup_for(case, res)
    cfg = xfref.make_cfg(case, res)
    b = s.block
    for a, c in [(b.Xmax, cfg.Xmax), (b.Ymax, cfg.Ymax), (b.Zmax, cfg.Zmax), (b.dx, cfg.dx), (b._dx, cfg._dx), (b._dy, cfg._dy), (b._dz, cfg._dz),
                 (b.CFLnumber, cfg.CFL), (s.Emax, cfg.Emax), (s.num_species, cfg.NS), (s.cop, cfg.cop), (s.ghost_species, cfg.ghost_species)]:
        assert a == c
    assert s.bc == list(cfg.bc)
    U, T = s.initial_condition()
    assert np.array_equal(U, g["ic_U"])
    assert np.array_equal(T, g["ic_T"])


def test_thermal_tables_match_reference_parser():
    s = setup_for("sbi", (24, 12, 12))
    names, _, _ = xfref.read_species("Inert-SBI")
    Hia, Hib, Wi, _Wi, Ri = xfref.read_thermal(names)
    t = s.thermal
    assert np.array_equal(np.ctypeslib.as_array(t.Hia, (5 * 21,)), Hia)
    assert np.array_equal(np.ctypeslib.as_array(t.Hib, (5 * 6,)), Hib)
    assert np.array_equal(np.ctypeslib.as_array(t.Ri, (5,)), Ri)
    assert np.array_equal(np.ctypeslib.as_array(t._Wi, (5,)), _Wi)


def test_cli_overrides_and_slab_decomposition():
    """-mpi / -mpi-s semantics of iniset.cpp:110-133 and the per-rank boundary list (mpiPacks.cpp:44-72)."""
    one = setup_for("sbi", (16, 8, 8))
    with pytest.raises(capi.XfError):
        setup_for("sbi", (16, 8, 8), extra=["-mpi=1,1,2", "-mpi-s=weak"])   # mz must equal the number of ranks
    lo = host.Setup(os.path.join(REPO, "settings", "shock-bubble.json"), ["-run=16,8,8", "-mpi=1,1,2", "-mpi-s=weak"], rank=0, nranks=2)
    hi = host.Setup(os.path.join(REPO, "settings", "shock-bubble.json"), ["-run=16,8,8", "-mpi=1,1,2", "-mpi-s=weak"], rank=1, nranks=2)
    assert lo.block.Z_inner == 8 and lo.block.dz == one.block.dz          # weak: domain height x mz, same dz
    assert lo.bc == [0, 1, 2, 1, 2, 99] and hi.bc == [0, 1, 2, 1, 99, 1]
    st = host.Setup(os.path.join(REPO, "settings", "expanded-jet.json"), ["-run=16,8,8", "-mpi=1,1,2", "-mpi-s=strong"], rank=1, nranks=2)
    assert st.block.Z_inner == 4 and st.myMpiPos_z == 1
    # the two weak-scaling slabs tile the one-rank domain of doubled height: rank 1's cells continue rank 0's in z
    big = host.Setup(os.path.join(REPO, "settings", "shock-bubble.json"), ["-run=16,8,16", "-domain=0.1,0.05,0.1"])
    Ub, Tb = big.initial_condition()
    U0, _ = lo.initial_condition()
    U1, _ = hi.initial_condition()
    E = big.Emax
    Ub = Ub.reshape(big.block.Zmax, -1, E)
    U0 = U0.reshape(lo.block.Zmax, -1, E)
    U1 = U1.reshape(hi.block.Zmax, -1, E)
    assert np.array_equal(U0[4:12], Ub[4:12]) and np.array_equal(U1[4:12], Ub[12:20])
    assert np.array_equal(U0[12:16], Ub[12:16])   # rank 0's upper ghost planes = rank 1's first inner planes


def test_output_stamps_parse_like_reference():
    s = setup_for("sbi", (16, 8, 8))
    # 5 arrays x 100 stamps + the inserted stamps (iniset.cpp:201-285); the last one (0.0005) lies beyond every array
    # stamp, so the reference's "insert before the first later stamp" loop never places it
    assert len(s.stamps) == 515
    assert abs(s.stamps[0] - 0.5e-6) < 1e-18 and s.stamps == sorted(s.stamps)
    assert setup_for("shock-tube", (400, 0, 0)).stamps == [0.00004]


def test_json_comment_rule_is_first_slash(tmp_path):
    """read_json.cpp:171 cuts every line at the first '/' character."""
    p = tmp_path / "c.json"
    p.write_text('{ "run": {"CFLnumber": 0.25, // trailing comment\n "nStepMax": 7}, / odd single slash\n "mesh": {"Resolution": [32,0,0]},\n'
                 ' "b200": {"sample": "1d-insert-st", "mixture": "1d-mc-insert-shock-tube"}}\n')
    s = host.Setup(str(p))
    assert s.block.CFLnumber == 0.25 and s.nStepmax == 7 and s.block.X_inner == 32


def test_abi_exports_every_declared_symbol():
    """include/xfluids_b200.h <-> libxfluids_b200.so: every declared entry point is exported (no compute call here)."""
    hdr = open(os.path.join(REPO, "include", "xfluids_b200.h")).read()
    declared = set(re.findall(r"\b(xf_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    dll = C.CDLL(capi.lib_path())
    for sym in declared:
        assert hasattr(dll, sym), sym
    hdll = C.CDLL(host.host_lib_path())
    for sym in host.HOST_SYMBOLS:
        assert hasattr(hdll, sym), sym


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the engine must refuse to construct (XF_ERR_CUDA), never compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    s = setup_for("vortex", (16, 16, 0))
    with pytest.raises(capi.XfError):
        capi.Engine(s.block, s.thermal, s.scheme)


def test_scheme_switches_from_cli_and_json(tmp_path):
    """-weno=6 (WENO-CU6), -pp= / equations.PositivityPreserving, -cfl= and -alpha= reach the xf_scheme / xf_block the engine is created with."""
    s = setup_for("sbi", (16, 8, 8), weno=6, extra=["-pp=1", "-cfl=0.9", "-alpha=GLF"])
    assert (s.scheme.weno_order, s.scheme.positivity, s.scheme.artificial_type, s.scheme.fp_mode) == (6, 1, 3, 0)
    assert s.block.CFLnumber == 0.9
    s = setup_for("sbi", (16, 8, 8))
    assert (s.scheme.weno_order, s.scheme.positivity, s.scheme.artificial_type) == (5, 0, 2) and s.block.CFLnumber == 0.4
    # the JSON key (read_json.cpp:68) without any command-line override
    src = open(os.path.join(REPO, "settings", "shock-bubble.json")).read()
    assert '"PositivityPreserving": false' in src
    p = tmp_path / "sbi_pp.json"
    p.write_text(src.replace('"PositivityPreserving": false', '"PositivityPreserving": true'))
    s = host.Setup(str(p), ["-run=16,8,8"])
    assert s.scheme.positivity == 1
