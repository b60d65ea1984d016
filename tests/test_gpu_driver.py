"""GPU (-m gpu): the `xfluids` executable (C++ host mirror of main.cpp / XFLUIDS::Evolution on the CUDA engine): checkpoint
written in the reference's CheckingPoint layout (XFLUIDS.cpp:658-687) and restart from it (XFLUIDS.cpp:616-623, 689-724)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import xfref

pytestmark = pytest.mark.gpu
EXE = os.path.join(xfref.REPO, "xfluids_b200", "xfluids")


def read_ckpt(path, n):
    raw = open(path, "rb").read()
    step, = struct.unpack_from("<i", raw, 0)
    time, = struct.unpack_from("<d", raw, 4)
    U = np.frombuffer(raw, dtype="<f8", count=n, offset=16)
    assert len(raw) == 16 + 8 * n
    return step, time, U


def run(args):
    r = subprocess.run([EXE] + args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


@pytest.mark.parametrize("js,grid,E,cfl_extra", [("2d-euler-vortex.json", (64, 64, 0), 5, []), ("shock-bubble.json", (24, 12, 12), 9, ["-weno=6", "-pp=1", "-cfl=0.9"])])
def test_checkpoint_and_restart(tmp_path, js, grid, E, cfl_extra):
    import xfgpu  # noqa: F401  (fails loudly if the CUDA library is missing)
    from xfluids_b200 import capi, host
    settings = os.path.join(xfref.REPO, "settings", js)
    g = "%d,%d,%d" % grid
    b, c = str(tmp_path / "b.ckpt"), str(tmp_path / "c.ckpt")
    run([settings, "-run=%s,6" % g, "-ckpt=" + b, "-quiet"] + cfl_extra)
    out = run([settings, "-run=%s,12" % g, "-restart=" + b, "-ckpt=" + c, "-quiet"] + cfl_extra)
    assert "steps=12" in out
    s = host.Setup(settings, ["-run=%s" % g] + cfl_extra)
    n = s.ncells * E
    sb, tb, Ub = read_ckpt(b, n)
    sc, tc, Uc = read_ckpt(c, n)
    assert (sb, sc) == (6, 12) and tc > tb > 0
    # the same continuation through the C ABI: the reference's restart keeps everything but Step, Time and U as the initial
    # condition left it (the Newton warm-start T in particular), then BC + UpdateStates, then the time loop
    eng = capi.Engine(s.block, s.thermal, s.scheme, device=0, keepalive=(s,))
    U0, T0 = s.initial_condition()
    eng.set_state(U0, T0)                                   # U = U1 = initial condition, T = its warm start (InitialU)
    eng.upload(eng.U, np.ascontiguousarray(Ub))            # Read_Ubak replaces d_U only: U1's never-refilled Inflow ghosts stay at the IC
    eng.set_time(tb)
    eng.boundary(eng.U, s.bc)
    assert eng.update_states(eng.U) == 0
    # XFLUIDS::Evolution: dt is clipped to every output time stamp on the way (XFLUIDS.cpp:167-199)
    left, t = 6, tb
    for stamp in s.stamps:
        if left == 0:
            break
        if stamp <= t:
            continue
        done, t, err = eng.run(s.bc, left, t_end=stamp)
        assert err == 0
        left -= done
    assert left == 0
    assert t == tc
    assert np.array_equal(eng.download(eng.U), Uc)


@pytest.mark.parametrize("case,res,weno,pp,chunks,alpha", [("sbi", (40, 16, 40), 5, 0, 8, "LLF"), ("sbi", (24, 12, 27), 6, 1, 5, "ROE"), ("jet", (24, 12, 24), 7, 0, 3, "LLF"),
                                                          ("sbi", (24, 12, 32), 5, 0, 4, "GLF")])
def test_host_step_overlapped_upload_is_bitwise_identical(case, res, weno, pp, chunks, alpha):
    """xf_step_host with the chunked upload overlapped with the plane-local part of stage 1 (the bench's e2e leg) against the plain
    upload -> step -> download sequence: identical bits in the returned host buffer, over several steps."""
    import ctypes as C
    import xfgpu  # noqa: F401
    from xfluids_b200 import capi, host
    SET = {"sbi": "shock-bubble.json", "jet": "expanded-jet.json"}
    cli = ["-run=%d,%d,%d" % res, "-weno=%d" % weno, "-alpha=" + alpha] + (["-pp=1", "-cfl=0.9"] if pp else [])   # GLF: the library falls back to the plain sequence
    s = host.Setup(os.path.join(xfref.REPO, "settings", SET[case]), cli)
    U0, T0 = s.initial_condition()
    outs = []
    for nch in (0, chunks):
        eng = capi.Engine(s.block, s.thermal, s.scheme, device=0, keepalive=(s,))
        L = eng.L
        L.check(L.dll.xf_set_host_overlap(eng.ctx, nch))
        eng.set_state(U0, T0)
        eng.boundary(eng.U, s.bc)
        assert eng.update_states(eng.U) == 0          # primitives + the dt maxima of the state the host buffer holds
        h = np.ascontiguousarray(eng.download(eng.U))
        hp = h.ctypes.data_as(C.c_void_p).value
        for _ in range(4):
            done, err = eng.step_host(hp, s.bc, 1)
            assert (done, err) == (1, 0)
        outs.append((h.copy(), eng.time()[0], eng.get_scalar("T").copy()))
        eng.close()
    assert outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("js,grid,E,scaling,extra", [("expanded-jet.json", (32, 16, 32), 7, "strong", []), ("shock-bubble.json", (32, 16, 16), 9, "weak", ["-weno=6", "-pp=1", "-cfl=0.9"]),
                                                      ("shock-bubble.json", (24, 12, 24), 9, "strong", ["-alpha=GLF", "-blocks"])])
def test_executable_two_ranks_equal_one_block_bitwise(tmp_path, js, grid, E, scaling, extra):
    """`xfluids ... -mpi=1,1,2`: one process, one host thread per GPU, the C++ NCCL slab stepper (halo exchange overlapped with interior
    work, dt MAX-reduced on the device) -- the reference's N > 1 flow (mpiPacks.cpp:357-505, Fluids.cpp:902-913) entirely in C++.
    The two ranks' checkpoints must tile the single-GPU run's checkpoint BIT FOR BIT (-blocks: the call-by-call loop with the
    host-side reductions; GLF: the nine running maxima are MAX-reduced every stage)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from xfluids_b200 import host
    settings = os.path.join(xfref.REPO, "settings", js)
    nx, ny, nz = grid
    one, two = str(tmp_path / "one.ckpt"), str(tmp_path / "two.ckpt")
    if scaling == "strong":
        run([settings, "-run=%d,%d,%d,6" % grid, "-ckpt=" + one, "-quiet"] + extra)
        out = run([settings, "-run=%d,%d,%d,6" % grid, "-mpi=1,1,2", "-mpi-s=strong", "-ckpt=" + two, "-quiet"] + extra)
        zi = nz // 2
        s1 = host.Setup(settings, ["-run=%d,%d,%d" % grid] + extra)
    else:
        s0 = host.Setup(settings, ["-run=%d,%d,%d" % grid] + extra)
        dom = "-domain=%.17g,%.17g,%.17g" % (0.1, 0.05, 0.05 * 2)      # shock-bubble.json DOMAIN_Size with the z extent doubled
        run([settings, "-run=%d,%d,%d,6" % (nx, ny, 2 * nz), dom, "-ckpt=" + one, "-quiet"] + extra)
        out = run([settings, "-run=%d,%d,%d,6" % grid, "-mpi=1,1,2", "-mpi-s=weak", "-ckpt=" + two, "-quiet"] + extra)
        zi = nz
        s1 = host.Setup(settings, ["-run=%d,%d,%d" % (nx, ny, 2 * nz), dom] + extra)
        del s0
    assert "ranks=2" in out and "steps=6" in out
    b = s1.block
    plane = b.Xmax * b.Ymax * E
    st1, t1, U1 = read_ckpt(one, b.Zmax * plane)
    U1 = U1.reshape(b.Zmax, plane)
    Bz = b.Bwidth_Z
    for r in range(2):
        st, t, U = read_ckpt(two + ".rank%d" % r, (zi + 2 * Bz) * plane)
        assert (st, t) == (st1, t1)
        assert np.array_equal(U.reshape(zi + 2 * Bz, plane)[Bz:Bz + zi], U1[Bz + r * zi:Bz + (r + 1) * zi]), "rank %d" % r
