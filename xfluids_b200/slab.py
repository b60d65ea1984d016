"""z-slab decomposition across the GPUs of one box: the replacement for the reference's MpiTrans (src/mpiPacks/mpiPacks.cpp:3-75,
357-505) and the MPI calls inside FluidBoundaryCondition / GetFluidDt.

One process per GPU (torch.distributed, backend nccl; gloo in the CPU tests).  Rank r owns Z_inner planes; its z faces carry
BC_COPY (99) towards a neighbour rank and the physical boundary condition on the outside (mpiPacks.cpp:44-72; the host Setup
computes that list).  Per RK stage, after the local ghost fill in x and y:

    pack  the Bwidth_Z innermost planes next to each internal face   (FluidMpiCopyKernelZ pack, BCs_kernels.hpp:304-324)
    send  them to that neighbour / receive the neighbour's planes     (MPI_Sendrecv, mpiPacks.cpp:486-489 -> NCCL send/recv)
    unpack into the ghost planes of that face                         (FluidMpiCopyKernelZ unpack)

and once per step the three directional dt maxima are MAX-reduced over the ranks (Fluids.cpp:902-913; max is exact, so every
rank derives the same dt bit for bit).  This module holds the transport-independent part; the device kernels are
xf_halo_pack / xf_halo_unpack of the C ABI."""
import ctypes as C

import torch
import torch.distributed as dist

BC_COPY = 99


class HaloExchanger:
    """Neighbour exchange of packed z-halo buffers + the scalar reductions of one time step."""

    def __init__(self, rank, world, bc, periodic_z=False, group=None):
        self.rank, self.world, self.group = rank, world, group
        # internal faces are the ones the host Setup marked BC_COPY (face 4 = zmin, 5 = zmax)
        self.lo = (rank - 1) % world if bc[4] == BC_COPY else None
        self.hi = (rank + 1) % world if bc[5] == BC_COPY else None
        if not periodic_z:
            assert (self.lo is None) == (rank == 0) or world == 1
            assert (self.hi is None) == (rank == world - 1) or world == 1

    def exchange(self, send_lo, send_hi, recv_lo, recv_hi):
        """send_lo: my inner planes next to zmin -> neighbour rank-1 (lands in ITS zmax ghosts = its recv_hi);
        send_hi: my inner planes next to zmax -> neighbour rank+1 (its recv_lo).  Returns the outstanding requests."""
        ops = []
        if self.hi is not None:
            ops.append(dist.P2POp(dist.isend, send_hi, self.hi, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, self.hi, group=self.group))
        if self.lo is not None:
            ops.append(dist.P2POp(dist.isend, send_lo, self.lo, group=self.group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, self.lo, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t


class _DevPtr:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap library-owned memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def wrap_device(ptr, n, dtype=torch.float64, device=None):
    ts = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevPtr(ptr, (n,), ts), device=device)


class SlabStepper:
    """The time loop of one rank of a z-slab run: XFLUIDS::Evolution's inner loop (src/XFLUIDS.cpp:172-294) with the halo exchange
    where FluidBoundaryCondition has its MPI exchange and the dt MAX-reduction where GetFluidDt has its allreduce.  Everything
    is enqueued on the engine's stream; nothing synchronises with the host inside a step."""

    def __init__(self, eng, bc, rank, world, device, overlap=True):
        self.eng, self.bc, self.rank, self.world = eng, list(bc), rank, world
        # GLF needs the running maxima of |lambda| over the WHOLE domain before every sweep (ConVenction_block.hpp:115-215): a MAX
        # all-reduce of 9 doubles between primitive recovery and sweeps of every stage; that stage runs un-overlapped
        sc = getattr(eng, "scheme", None)
        self.glf = world > 1 and sc is not None and sc.artificial_type == 3 and sc.weno_order != 7
        self.overlap = overlap and world > 1 and eng.block.DimZ and not self.glf
        self.comm = None
        self.hx = HaloExchanger(rank, world, self.bc)
        L = eng.L.dll
        n = L.xf_halo_doubles(eng.ctx)
        self.buf = {k: torch.empty(n, dtype=torch.float64, device=device) for k in ("send_lo", "send_hi", "recv_lo", "recv_hi")}
        self.dtmax = wrap_device(L.xf_device_dtmax(eng.ctx), 3, torch.float64, device)
        self.errors = wrap_device(L.xf_device_errors(eng.ctx), 4, torch.int32, device)
        self.glfmax = wrap_device(L.xf_device_glfmax(eng.ctx), 9, torch.float64, device) if self.glf else None

    def halo(self, field):
        """z exchange of `field` (a device pointer of the engine): pack -> send/recv -> unpack, on the current stream."""
        e, L, b = self.eng, self.eng.L, self.buf
        if self.hx.lo is not None:
            L.check(L.dll.xf_halo_pack(e.ctx, field, 4, b["send_lo"].data_ptr()))
        if self.hx.hi is not None:
            L.check(L.dll.xf_halo_pack(e.ctx, field, 5, b["send_hi"].data_ptr()))
        for r in self.hx.exchange(b["send_lo"], b["send_hi"], b["recv_lo"], b["recv_hi"]):
            r.wait()
        if self.hx.lo is not None:
            L.check(L.dll.xf_halo_unpack(e.ctx, field, 4, b["recv_lo"].data_ptr()))
        if self.hx.hi is not None:
            L.check(L.dll.xf_halo_unpack(e.ctx, field, 5, b["recv_hi"].data_ptr()))

    def startup(self):
        """main.cpp:44-48: BoundaryCondition + UpdateStates on the initial U."""
        e = self.eng
        e.boundary(e.U, self.bc)
        self.halo(e.U)
        assert e.update_states(e.U) == 0
        if self.glf:
            self.hx.allreduce_max(self.glfmax)

    # ---- one stage, exchange blocking (reference order: BC + exchange, UpdateStates, GetLU, UpdateU) ----
    def stage_blocking(self, flag):
        e = self.eng
        UI = e.U if flag == 1 else e.U1
        e.boundary(UI, self.bc)
        self.halo(UI)
        if self.glf:
            L = e.L
            L.check(L.dll.xf_stage_states(e.ctx, e.U, e.U1, flag))
            self.hx.allreduce_max(self.glfmax)
            L.check(L.dll.xf_stage_fluxes(e.ctx, e.U, e.U1, e.LU, flag))
        else:
            e.rk_stage(None, flag)

    # ---- one stage, exchange overlapped with the interior work ----
    def stage_overlapped(self, flag):
        """pack on the compute stream (before the primitive recovery rewrites the species of U), then the exchange and the
        unpack on the communication stream while the compute stream recovers the primitives of all non-ghost planes and runs
        the x and y sweeps; the z ghost planes' primitives, the z sweep and the update wait for the unpack."""
        e, L, b = self.eng, self.eng.L, self.buf
        main = torch.cuda.current_stream()
        if self.comm is None:
            self.comm = torch.cuda.Stream(device=main.device)
            self.ev_packed = torch.cuda.Event()
            self.ev_unpacked = torch.cuda.Event()
        UI = e.U if flag == 1 else e.U1
        e.boundary(UI, self.bc)
        if self.hx.lo is not None:
            L.check(L.dll.xf_halo_pack(e.ctx, UI, 4, b["send_lo"].data_ptr()))
        if self.hx.hi is not None:
            L.check(L.dll.xf_halo_pack(e.ctx, UI, 5, b["send_hi"].data_ptr()))
        self.ev_packed.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_packed)
            for r in self.hx.exchange(b["send_lo"], b["send_hi"], b["recv_lo"], b["recv_hi"]):
                r.wait()
            cs = self.comm.cuda_stream
            if self.hx.lo is not None:
                L.check(L.dll.xf_halo_unpack_on(e.ctx, UI, 4, b["recv_lo"].data_ptr(), cs))
            if self.hx.hi is not None:
                L.check(L.dll.xf_halo_unpack_on(e.ctx, UI, 5, b["recv_hi"].data_ptr(), cs))
            self.ev_unpacked.record(self.comm)
        e.stage_interior(flag)
        main.wait_event(self.ev_unpacked)
        e.stage_finish(flag)

    def step(self, t_end=1e300):
        e = self.eng
        self.hx.allreduce_max(self.dtmax)       # Fluids.cpp:902-913
        e.dt_device(t_end)                      # XFLUIDS.cpp:196-199
        for flag in (1, 2, 3):                  # XFLUIDS.cpp:441-525
            if self.overlap:
                self.stage_overlapped(flag)
            else:
                self.stage_blocking(flag)

    def step_host(self, h_ptr, t_end=1e300):
        """One step from / to this rank's pinned host buffer (AoS U of its slab) with the PCIe copies overlapped like xf_step_host:
        chunked upload under the plane-local part of stage 1, stage 3 in z-chunks with the download behind it; the z-halo
        exchanges sit between the pieces.  Bit-identical to upload -> step() -> download.  Returns False when the overlapped path
        does not apply (GLF, 1-D / 2-D) and the caller has to use the plain sequence."""
        e, L = self.eng, self.eng.L
        if self.glf or not e.block.DimZ:
            return False
        bc = (C.c_int * 6)(*self.bc)
        self.hx.allreduce_max(self.dtmax)
        rc = L.dll.xf_host_begin(e.ctx, C.c_void_p(h_ptr), bc, t_end, e.U, e.U1)
        if rc != 0:
            return False
        self.halo(e.U)
        L.check(L.dll.xf_host_stage1_finish(e.ctx, bc, e.U, e.U1, e.LU))
        if self.overlap:
            self.stage_overlapped(2)
        else:
            self.stage_blocking(2)
        e.boundary(e.U1, self.bc)
        self.halo(e.U1)
        err = C.c_int()
        L.check(L.dll.xf_host_stage3(e.ctx, C.c_void_p(h_ptr), e.U, e.U1, e.LU, C.byref(err)), allow_numeric=True)
        return True

    def steps(self, n, t_end=1e300):
        for _ in range(n):
            self.step(t_end)

    def any_error(self):
        f = self.errors.clone()
        self.hx.allreduce_max(f)
        return bool(f[:3].any().item())
