#!/usr/bin/env python3
"""Multi-rank check of the z-slab path, launched by tests/test_slab.py (and by hand under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/slab_check.py --mode gpu|cpu [--steps 5] [--case sbi|jet] [--strong]

mode gpu (nccl, one GPU per rank): every rank advances its slab of the shock-bubble (weak scaling: the z extent grows with the
          ranks) or jet (strong: a fixed box is cut) block for --steps steps, once with the blocking exchange and once with the
          overlapped one; rank 0 also advances the undecomposed block on its own GPU; the gathered slabs must equal it BIT FOR BIT
          (strict mode: no floating-point reduction crosses ranks except max).
mode cpu (gloo, no GPU): the exchange protocol alone -- numpy stands in for the pack/unpack kernels -- must reproduce the z ghost
          planes of every slab from the undecomposed block's initial condition, and the MAX all-reduce must agree on all ranks.
Prints SLAB_CHECK_OK on rank 0."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from xfluids_b200 import host  # noqa: E402
from xfluids_b200.slab import HaloExchanger  # noqa: E402

CASES = {"sbi": ("shock-bubble.json", (32, 16, 16), (0.1, 0.05, 0.05)), "jet": ("expanded-jet.json", (32, 16, 16), (3.0, 1.5, 1.5))}


def setups(case, rank, world, strong, extra=()):
    js, grid, dom = CASES[case]
    path = os.path.join(REPO, "settings", js)
    extra = list(extra)
    if strong:
        gz = grid[2] * world   # a fixed box whose z extent is cut into `world` slabs
        one = host.Setup(path, ["-run=%d,%d,%d" % (grid[0], grid[1], gz)] + extra)
        mine = host.Setup(path, ["-run=%d,%d,%d" % (grid[0], grid[1], gz), "-mpi=1,1,%d" % world, "-mpi-s=strong"] + extra, rank=rank, nranks=world)
    else:
        one = host.Setup(path, ["-run=%d,%d,%d" % (grid[0], grid[1], grid[2] * world), "-domain=%.17g,%.17g,%.17g" % (dom[0], dom[1], dom[2] * world)] + extra)
        mine = host.Setup(path, ["-run=%d,%d,%d" % grid, "-mpi=1,1,%d" % world, "-mpi-s=weak"] + extra, rank=rank, nranks=world)
    return one, mine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="gpu")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--case", default="sbi")
    ap.add_argument("--strong", action="store_true")
    ap.add_argument("--weno", type=int, default=5)
    ap.add_argument("--pp", type=int, default=0, help="positivity-preserving limiter on, at CFL 0.9 (where it acts)")
    ap.add_argument("--alpha", default="LLF", help="LLF | ROE | GLF (GLF: 9 running maxima MAX-reduced over the ranks every stage)")
    ap.add_argument("--visc", action="store_true", help="viscous / heat-conduction / species-diffusion terms on (the slabs' split stages run the stand-alone wall-flux "
                    "kernel after the z halo has arrived, the undecomposed block the sweeps' viscous tails)")
    ap.add_argument("--periodic-z", action="store_true", help="periodic z boundary: the outer faces of the first / last rank exchange with each other (on 2 ranks both neighbours are the same peer)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    extra = ["-weno=%d" % a.weno, "-alpha=" + a.alpha] + (["-pp=1", "-cfl=0.9"] if a.pp else []) + (["-visc=1"] if a.visc else [])
    if a.periodic_z:
        js, _, _ = CASES[a.case]
        bc0 = host.Setup(os.path.join(REPO, "settings", js), []).bc
        extra.append("-bc=" + ",".join(str(b) for b in bc0[:4] + [3, 3]))
    one, mine = setups(a.case, rank, world, a.strong, extra)
    E, Bz = mine.Emax, mine.block.Bwidth_Z
    zi = mine.block.Z_inner
    plane = mine.block.Xmax * mine.block.Ymax * E
    U1, T1 = one.initial_condition()
    Um, Tm = mine.initial_condition()
    # a slab's inner planes are the undecomposed block's planes [rank*zi, (rank+1)*zi) (+Bz ghost offset)
    U1p = U1.reshape(one.block.Zmax, plane)
    Ump = Um.reshape(mine.block.Zmax, plane)
    assert np.array_equal(Ump[Bz:Bz + zi], U1p[Bz + rank * zi:Bz + (rank + 1) * zi]), "slab initial condition does not tile the block"

    if a.mode == "cpu":
        dist.init_process_group("gloo")
        hx = HaloExchanger(rank, world, mine.bc)
        # scribble over the z ghosts that the exchange must fill, then pack / exchange / unpack with numpy
        work = Ump.copy()
        if hx.lo is not None:
            work[:Bz] = -1.0
        if hx.hi is not None:
            work[-Bz:] = -1.0
        send_lo = torch.from_numpy(work[Bz:2 * Bz].copy())
        send_hi = torch.from_numpy(work[-2 * Bz:-Bz].copy())
        recv_lo, recv_hi = torch.empty_like(send_lo), torch.empty_like(send_hi)
        for r in hx.exchange(send_lo, send_hi, recv_lo, recv_hi):
            r.wait()
        if hx.lo is not None:
            work[:Bz] = recv_lo.numpy()
        if hx.hi is not None:
            work[-Bz:] = recv_hi.numpy()
        lo, hi = rank * zi, (rank + 1) * zi + 2 * Bz
        ref = U1p[lo:hi].copy()
        if a.periodic_z:   # the ghosts of the outer faces are the inner planes at the other end of the undecomposed block
            ztot = zi * world
            if rank == 0:
                ref[:Bz] = U1p[ztot:ztot + Bz]
            if rank == world - 1:
                ref[-Bz:] = U1p[Bz:2 * Bz]
            assert hx.lo == (rank - 1) % world and hx.hi == (rank + 1) % world
        if hx.lo is not None:
            assert np.array_equal(work[:Bz], ref[:Bz]), "zmin ghosts"
        if hx.hi is not None:
            assert np.array_equal(work[-Bz:], ref[-Bz:]), "zmax ghosts"
        assert np.array_equal(work[Bz:-Bz], ref[Bz:-Bz])
        m = torch.tensor([float(rank + 1), 10.0 - rank, 3.0], dtype=torch.float64)
        hx.allreduce_max(m)
        assert m.tolist() == [float(world), 10.0, 3.0]
        assert a.periodic_z or ((hx.lo is None) == (rank == 0) and (hx.hi is None) == (rank == world - 1))
        dist.barrier()
        if rank == 0:
            print("SLAB_CHECK_OK cpu world=%d case=%s" % (world, a.case))
        dist.destroy_process_group()
        return 0

    # ---- GPU ----
    from xfluids_b200 import capi
    from xfluids_b200.slab import SlabStepper
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    results = {}
    stream = torch.cuda.Stream(device=dev)
    for overlap in (False, True):
        eng = capi.Engine(mine.block, mine.thermal, mine.scheme, device=local, keepalive=(mine,))
        if a.visc:
            eng.set_transport(mine.transport, keepalive=(mine,))
        eng.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            eng.set_state(Um, Tm)
            st = SlabStepper(eng, mine.bc, rank, world, dev, overlap=overlap)
            st.startup()
            if overlap:
                done, _, err = st.run(a.steps)          # the polling loop of the C++ driver (xf_slab_run)
                assert (done, err) == (a.steps, 0)
            else:
                st.steps(a.steps)
            assert not st.any_error()
            torch.cuda.synchronize()
            results[overlap] = (eng.download(eng.U).reshape(mine.block.Zmax, plane), eng.time()[0])
            st.close()
        eng.close()
    assert np.array_equal(results[False][0], results[True][0]), "overlapped exchange changed the result"
    # host-buffer step (bench e2e on N > 1): the chunked, overlapped form against upload -> step -> download, 2 steps each
    if a.alpha != "GLF" and not a.visc:   # (the chunked host step is switched off with the viscous terms on)
        hres = {}
        for mode in ("plain", "overlapped"):
            eng = capi.Engine(mine.block, mine.thermal, mine.scheme, device=local, keepalive=(mine,))
            if a.visc:
                eng.set_transport(mine.transport, keepalive=(mine,))
            eng.set_stream(stream.cuda_stream)
            eng.L.check(eng.L.dll.xf_set_host_overlap(eng.ctx, 4))
            with torch.cuda.stream(stream):
                eng.set_state(Um, Tm)
                st = SlabStepper(eng, mine.bc, rank, world, dev, overlap=True)
                st.startup()
                st.steps(1)
                hb = np.ascontiguousarray(eng.download(eng.U))
                for _ in range(2):
                    if mode == "plain":
                        eng.upload(eng.U, hb)
                        eng.upload(eng.U1, hb)
                        st.step()
                        hb = np.ascontiguousarray(eng.download(eng.U))
                    else:
                        applied, herr = st.step_host(hb.ctypes.data)
                        assert applied and herr == 0, "overlapped host step not available"
                assert not st.any_error()
                torch.cuda.synchronize()
                hres[mode] = (hb.copy(), eng.time()[0])
                st.close()
            eng.close()
        assert hres["plain"][1] == hres["overlapped"][1], "host-step time differs"
        assert np.array_equal(hres["plain"][0], hres["overlapped"][0]), "overlapped host step differs from upload/step/download"
    mineU, tmine = results[True]
    gathered = [None] * world
    dist.all_gather_object(gathered, (mineU[Bz:Bz + zi], tmine))
    if rank == 0:
        eng = capi.Engine(one.block, one.thermal, one.scheme, device=local, keepalive=(one,))
        if a.visc:
            eng.set_transport(one.transport, keepalive=(one,))
        eng.set_state(U1, T1)
        eng.boundary(eng.U, one.bc)
        assert eng.update_states(eng.U) == 0
        done, t, err = eng.run(one.bc, a.steps)
        assert (done, err) == (a.steps, 0)
        Uone = eng.download(eng.U).reshape(one.block.Zmax, plane)
        for r in range(world):
            assert gathered[r][1] == t, ("time differs", gathered[r][1], t)
            assert np.array_equal(gathered[r][0], Uone[Bz + r * zi:Bz + (r + 1) * zi]), "rank %d slab differs from the undecomposed block" % r
        eng.close()
        print("SLAB_CHECK_OK gpu world=%d case=%s steps=%d strong=%d t=%.9e" % (world, a.case, a.steps, int(a.strong), t))
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
