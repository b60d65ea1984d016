"""z-slab decomposition across the GPUs of one box: Python face of the C++ slab stepper (csrc/xf_slab.cu, C ABI xf_comm_* /
xf_slab_*), the replacement for the reference's MpiTrans (src/mpiPacks/mpiPacks.cpp:3-75, 357-505) and the MPI calls inside
FluidBoundaryCondition / GetFluidDt.

One process per GPU under torchrun.  torch.distributed is plumbing only: it carries the 128-byte NCCL unique id from rank 0 to
the other ranks, after which every exchange (ncclSend / ncclRecv of the packed ghost planes, MAX all-reduce of the dt maxima and
the guard flags) is issued by the C++ stepper on CUDA streams.  Per RK stage, after the local ghost fill in x and y:

    pack  the Bwidth_Z innermost planes next to each internal face   (FluidMpiCopyKernelZ pack, BCs_kernels.hpp:304-324)
    send  them to that neighbour / receive the neighbour's planes     (MPI_Sendrecv, mpiPacks.cpp:486-489 -> NCCL send/recv)
    unpack into the ghost planes of that face                         (FluidMpiCopyKernelZ unpack)

HaloExchanger below is the same protocol over torch.distributed point-to-point operations (gloo in the CPU tests): neighbour
bookkeeping and posting order, including the periodic-z case where both neighbours are the same rank."""
import ctypes as C

import torch
import torch.distributed as dist

BC_COPY = 99


class HaloExchanger:
    """Neighbour exchange of packed z-halo buffers + the scalar reductions of one time step (transport: torch.distributed)."""

    def __init__(self, rank, world, bc, group=None):
        self.rank, self.world, self.group = rank, world, group
        # internal faces are the ones the host Setup marked BC_COPY (face 4 = zmin, 5 = zmax); with a periodic z boundary that
        # includes the outer faces of the first and the last rank, whose neighbour is the rank at the other end
        self.lo = (rank - 1) % world if bc[4] == BC_COPY else None
        self.hi = (rank + 1) % world if bc[5] == BC_COPY else None
        assert world > 1 or (self.lo is None and self.hi is None), "BC_COPY faces on a single rank"

    def exchange(self, send_lo, send_hi, recv_lo, recv_hi):
        """send_lo: my inner planes next to zmin -> neighbour rank-1 (lands in ITS zmax ghosts = its recv_hi);
        send_hi: my inner planes next to zmax -> neighbour rank+1 (its recv_lo).  Returns the outstanding requests.

        Posting order: sends (hi, lo), receives (lo, hi).  Point-to-point operations between one pair of ranks match in posting
        order; when both neighbours are the SAME rank (periodic z on two ranks) my first send -- my upper planes -- must meet
        the peer's first receive, which therefore has to be its lower ghosts."""
        ops = []
        if self.hi is not None:
            ops.append(dist.P2POp(dist.isend, send_hi, self.hi, group=self.group))
        if self.lo is not None:
            ops.append(dist.P2POp(dist.isend, send_lo, self.lo, group=self.group))
        if self.lo is not None:
            ops.append(dist.P2POp(dist.irecv, recv_lo, self.lo, group=self.group))
        if self.hi is not None:
            ops.append(dist.P2POp(dist.irecv, recv_hi, self.hi, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    def allreduce_max(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t


def make_comm(L, rank, world, device_index):
    """xf_comm of this rank; the NCCL unique id travels from rank 0 over torch.distributed (any backend)."""
    idb = (C.c_char * 128)()
    if world > 1:
        payload = [None]
        if rank == 0:
            L.check(L.dll.xf_comm_unique_id(idb))
            payload = [bytes(idb.raw)]
        dist.broadcast_object_list(payload, src=0)
        idb = (C.c_char * 128).from_buffer_copy(payload[0])
    comm = C.c_void_p()
    L.check(L.dll.xf_comm_create(idb, rank, world, device_index, C.byref(comm)))
    return comm


class SlabStepper:
    """The time loop of one rank of a z-slab run: XFLUIDS::Evolution's inner loop (src/XFLUIDS.cpp:172-294) with the halo exchange
    where FluidBoundaryCondition has its MPI exchange and the dt MAX-reduction where GetFluidDt has its allreduce -- all of it in
    the C++ stepper (xf_slab_*); this class only holds the handles.  Everything is enqueued on the CUDA stream that is current
    when the stepper is created; nothing synchronises with the host inside a step."""

    def __init__(self, eng, bc, rank, world, device, overlap=True, comm=None):
        self.eng, self.bc, self.rank, self.world = eng, list(bc), rank, world
        L = self.L = eng.L
        dev_index = device.index if isinstance(device, torch.device) else int(device)
        self._own_comm = comm is None
        self.comm = comm if comm is not None else make_comm(L, rank, world, dev_index)
        sc = eng.scheme
        self.slab = C.c_void_p()
        stream = torch.cuda.current_stream(dev_index).cuda_stream
        L.check(L.dll.xf_slab_create(eng.ctx, self.comm, (C.c_int * 6)(*self.bc), sc.artificial_type, sc.weno_order, C.c_void_p(stream), C.byref(self.slab)))
        if not overlap:
            L.check(L.dll.xf_slab_set_overlap(self.slab, 0))

    def neighbours(self):
        v = (C.c_int * 2)()
        self.L.check(self.L.dll.xf_slab_neighbours(self.slab, v))
        return tuple(None if x < 0 else x for x in v)

    def halo(self, field):
        self.L.check(self.L.dll.xf_slab_halo(self.slab, field))

    def startup(self):
        """main.cpp:44-48: BoundaryCondition + UpdateStates on the initial U."""
        err = C.c_int()
        self.L.check(self.L.dll.xf_slab_startup(self.slab, self.eng.U, C.byref(err)))
        assert err.value == 0, "guards fired on the initial state"

    def step(self, t_end=1e300):
        e = self.eng
        self.L.check(self.L.dll.xf_slab_step(self.slab, e.U, e.U1, e.LU, t_end))

    def steps(self, n, t_end=1e300):
        for _ in range(n):
            self.step(t_end)

    def run(self, n, t_end=1e300):
        """n steps with the host looking at (time, error) like xf_run; returns (steps done, time, error)."""
        e = self.eng
        done, err, t = C.c_int(), C.c_int(), C.c_double()
        self.L.check(self.L.dll.xf_slab_run(self.slab, e.U, e.U1, e.LU, n, t_end, C.byref(done), C.byref(t), C.byref(err)), allow_numeric=True)
        return done.value, t.value, err.value

    def step_host(self, h_ptr, t_end=1e300):
        """One step from / to this rank's pinned host buffer (AoS U of its slab) with the PCIe copies overlapped like xf_step_host:
        chunked upload under the plane-local part of stage 1, stage 3 in z-chunks with the download behind it; the z-halo
        exchanges sit between the pieces.  Bit-identical to upload -> step() -> download.  Returns (applied, error): applied is
        False when the overlapped path does not apply (GLF, 1-D / 2-D) and the caller has to use the plain sequence; error is the
        numeric-guard flag of this rank."""
        e = self.eng
        err = C.c_int()
        rc = self.L.dll.xf_slab_step_host(self.slab, C.c_void_p(h_ptr), t_end, e.U, e.U1, e.LU, C.byref(err))
        if rc == -1:
            return False, 0
        self.L.check(rc, allow_numeric=True)
        return True, err.value

    def any_error(self):
        err = C.c_int()
        self.L.check(self.L.dll.xf_slab_any_error(self.slab, C.byref(err)))
        return bool(err.value)

    def close(self):
        if self.slab:
            self.L.dll.xf_slab_destroy(self.slab)
            self.slab = None
        if self._own_comm and self.comm:
            self.L.dll.xf_comm_destroy(self.comm)
            self.comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
