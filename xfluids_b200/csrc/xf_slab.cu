// xf_slab.cu -- z-slab decomposition across the GPUs of one box, in C++ behind the C ABI (include/xfluids_b200.h "multi-GPU"):
// the replacement for the reference's MpiTrans (src/mpiPacks/mpiPacks.cpp:3-75, 357-505), the MPI exchange inside
// FluidBoundaryCondition (src/solver_BCs/BCs_block.cpp:50-217) and the MPI_Allreduce of GetFluidDt (src/Fluids.cpp:902-913).
//
// Transport: NCCL send/recv + all-reduce(MAX), bound at run time with dlopen("libnccl.so.2") -- inside a torchrun rank that is the
// NCCL torch has already loaded, in the stand-alone `xfluids` executable the system's.  One xf_comm per rank; the ranks may be
// processes (bench.py under torchrun: the 128-byte unique id travels over torch.distributed) or threads of one process
// (`xfluids ... -mpi=1,1,N`: one thread per GPU).  The library works without NCCL as long as no communicator is created.
//
// Per RK stage (xf_slab_stage): local ghost fill -> pack the Bwidth_Z inner planes next to each internal face (on the compute
// stream, BEFORE the primitive recovery renormalises the species) -> send / recv + unpack on a communication stream, while the
// compute stream recovers the primitives of all non-ghost planes and runs the x and y sweeps (xf_stage_interior) -> the z ghost
// planes' primitives, the z sweep and the update wait for the unpack event (xf_stage_finish).  Once per step the three dt maxima are
// MAX-reduced in place on the device (max is exact: every rank derives the same dt bit for bit).  Global Lax-Friedrichs splitting adds
// a MAX all-reduce of the nine running maxima of |lambda| between primitive recovery and sweeps; that stage runs un-overlapped.
//
// Send / receive order inside one group: sends (hi, lo), receives (lo, hi).  With a periodic z boundary on two ranks both
// neighbours are the SAME peer and point-to-point operations match in posting order per peer: my first send (my upper planes)
// must meet the peer's first receive (its LOWER ghosts).  (slab.py of round 1 posted send_hi / recv_hi first and would have swapped
// the halos in that case -- ADVICE r1.)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include "../../include/xfluids_b200.h"

namespace
{
	struct NcclApi
	{
		decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
		decltype(&ncclCommInitRank) CommInitRank = nullptr;
		decltype(&ncclCommDestroy) CommDestroy = nullptr;
		decltype(&ncclSend) Send = nullptr;
		decltype(&ncclRecv) Recv = nullptr;
		decltype(&ncclAllReduce) AllReduce = nullptr;
		decltype(&ncclGroupStart) GroupStart = nullptr;
		decltype(&ncclGroupEnd) GroupEnd = nullptr;
		decltype(&ncclGetErrorString) GetErrorString = nullptr;
		bool ok = false;
		std::string why;
	};
	NcclApi &nccl()
	{
		static NcclApi api;
		static std::once_flag once;
		std::call_once(once, []()
					   {
			void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
			if (!h)
				h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
			if (!h)
			{
				api.why = std::string("cannot load libnccl.so.2: ") + dlerror();
				return;
			}
#define XF_SYM(name)                                                        \
	api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name)); \
	if (!api.name)                                                          \
	{                                                                       \
		api.why = "libnccl lacks nccl" #name;                               \
		return;                                                             \
	}
			XF_SYM(GetUniqueId) XF_SYM(CommInitRank) XF_SYM(CommDestroy) XF_SYM(Send) XF_SYM(Recv) XF_SYM(AllReduce) XF_SYM(GroupStart) XF_SYM(GroupEnd)
				XF_SYM(GetErrorString)
#undef XF_SYM
			api.ok = true; });
		return api;
	}
	thread_local std::string g_slab_err;
	int sfail(int code, const std::string &m)
	{
		g_slab_err = m;
		return code;
	}
} // namespace

#define NC(call)                                                                                                   \
	do                                                                                                             \
	{                                                                                                              \
		ncclResult_t r__ = (call);                                                                                 \
		if (r__ != ncclSuccess)                                                                                    \
			return sfail(XF_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r__));                     \
	} while (0)
#define CUS(call)                                                                                                  \
	do                                                                                                             \
	{                                                                                                              \
		cudaError_t e__ = (call);                                                                                  \
		if (e__ != cudaSuccess)                                                                                    \
			return sfail(XF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                        \
	} while (0)
#define XS(call)                                                                                                   \
	do                                                                                                             \
	{                                                                                                              \
		int r__ = (call);                                                                                          \
		if (r__ != XF_OK)                                                                                          \
			return sfail(r__, std::string(#call) + ": " + xf_last_error());                                        \
	} while (0)

struct xf_comm
{
	int rank = 0, world = 1, device = 0;
	ncclComm_t comm = nullptr;
};

struct xf_slab
{
	xf_ctx *c = nullptr;
	xf_comm *comm = nullptr;
	int bc[6];
	int lo = -1, hi = -1; // neighbour ranks across the zmin / zmax face, -1: physical boundary
	bool glf = false, overlap = true;
	size_t nhalo = 0;
	double *send_lo = nullptr, *send_hi = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
	cudaStream_t main = nullptr, comm_stream = nullptr;
	cudaEvent_t ev_packed = nullptr, ev_unpacked = nullptr;
	double *h_pin = nullptr; // 8 doubles of pinned host memory (time, dt, reductions of host values)
	int *h_err = nullptr;
	double *d_scratch = nullptr; // 16 device doubles
};

extern "C"
{
	const char *xf_slab_last_error(void) { return g_slab_err.c_str(); }

	int xf_comm_unique_id(char id[128])
	{
		if (!nccl().ok)
			return sfail(XF_ERR_COMM, nccl().why);
		ncclUniqueId u;
		static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
		NC(nccl().GetUniqueId(&u));
		std::memcpy(id, &u, 128);
		return XF_OK;
	}
	int xf_comm_create(const char id[128], int rank, int world, int device, xf_comm **out)
	{
		if (!out || (!id && world > 1) || world < 1 || rank < 0 || rank >= world)
			return sfail(XF_ERR_ARG, "xf_comm_create: bad argument");
		CUS(cudaSetDevice(device));
		xf_comm *m = new xf_comm();
		m->rank = rank, m->world = world, m->device = device;
		if (world == 1)
		{ // a single rank never communicates: no NCCL needed
			*out = m;
			return XF_OK;
		}
		if (!nccl().ok)
		{
			delete m;
			return sfail(XF_ERR_COMM, nccl().why);
		}
		ncclUniqueId u;
		std::memcpy(&u, id, 128);
		ncclResult_t r = nccl().CommInitRank(&m->comm, world, u, rank);
		if (r != ncclSuccess)
		{
			delete m;
			return sfail(XF_ERR_COMM, std::string("ncclCommInitRank: ") + nccl().GetErrorString(r));
		}
		*out = m;
		return XF_OK;
	}
	int xf_comm_destroy(xf_comm *m)
	{
		if (!m)
			return XF_OK;
		if (m->comm)
			nccl().CommDestroy(m->comm);
		delete m;
		return XF_OK;
	}
	int xf_comm_rank(const xf_comm *m) { return m->rank; }
	int xf_comm_world(const xf_comm *m) { return m->world; }
	// MAX over the ranks of n device doubles / ints, in place, on `stream`
	int xf_comm_allreduce_max(xf_comm *m, double *d_v, int n, void *stream)
	{
		if (m->world > 1)
			NC(nccl().AllReduce(d_v, d_v, (size_t)n, ncclDouble, ncclMax, m->comm, (cudaStream_t)stream));
		return XF_OK;
	}
	int xf_comm_allreduce_max_int(xf_comm *m, int *d_v, int n, void *stream)
	{
		if (m->world > 1)
			NC(nccl().AllReduce(d_v, d_v, (size_t)n, ncclInt32, ncclMax, m->comm, (cudaStream_t)stream));
		return XF_OK;
	}

	int xf_slab_create(xf_ctx *c, xf_comm *comm, const int bc[6], int artificial_type, int weno_order, void *compute_stream, xf_slab **out)
	{
		if (!c || !comm || !bc || !out)
			return sfail(XF_ERR_ARG, "xf_slab_create: null argument");
		CUS(cudaSetDevice(comm->device));
		xf_slab *s = new xf_slab();
		s->c = c, s->comm = comm;
		std::memcpy(s->bc, bc, sizeof(s->bc));
		// internal faces are the ones the host Setup marked BC_COPY (mpiPacks.cpp:44-72); a periodic z boundary makes the outer faces
		// of the first and the last rank internal too (neighbour = the rank at the other end)
		const int w = comm->world, r = comm->rank;
		s->lo = bc[4] == XF_BC_COPY ? (r - 1 + w) % w : -1;
		s->hi = bc[5] == XF_BC_COPY ? (r + 1) % w : -1;
		if (w == 1 && (s->lo >= 0 || s->hi >= 0))
		{
			delete s;
			return sfail(XF_ERR_ARG, "BC_COPY faces on a single rank (periodic z on one GPU is the local Periodic ghost fill)");
		}
		s->glf = w > 1 && artificial_type == 3 && weno_order != 7;
		s->overlap = w > 1 && !s->glf;
		s->main = (cudaStream_t)compute_stream;
		s->nhalo = xf_halo_doubles(c);
		double **bufs[4] = {&s->send_lo, &s->send_hi, &s->recv_lo, &s->recv_hi};
		for (double **b : bufs)
			CUS(cudaMalloc((void **)b, s->nhalo * sizeof(double)));
		CUS(cudaMalloc((void **)&s->d_scratch, 16 * sizeof(double)));
		CUS(cudaStreamCreateWithFlags(&s->comm_stream, cudaStreamNonBlocking));
		CUS(cudaEventCreateWithFlags(&s->ev_packed, cudaEventDisableTiming));
		CUS(cudaEventCreateWithFlags(&s->ev_unpacked, cudaEventDisableTiming));
		CUS(cudaMallocHost((void **)&s->h_pin, 8 * sizeof(double)));
		CUS(cudaMallocHost((void **)&s->h_err, 4 * sizeof(int)));
		XS(xf_set_stream(c, compute_stream));
		*out = s;
		return XF_OK;
	}
	int xf_slab_destroy(xf_slab *s)
	{
		if (!s)
			return XF_OK;
		cudaSetDevice(s->comm->device);
		cudaStreamSynchronize(s->main);
		if (s->comm_stream)
			cudaStreamSynchronize(s->comm_stream), cudaStreamDestroy(s->comm_stream);
		double *bufs[5] = {s->send_lo, s->send_hi, s->recv_lo, s->recv_hi, s->d_scratch};
		for (double *b : bufs)
			if (b)
				cudaFree(b);
		if (s->ev_packed)
			cudaEventDestroy(s->ev_packed);
		if (s->ev_unpacked)
			cudaEventDestroy(s->ev_unpacked);
		if (s->h_pin)
			cudaFreeHost(s->h_pin);
		if (s->h_err)
			cudaFreeHost(s->h_err);
		delete s;
		return XF_OK;
	}
	int xf_slab_set_overlap(xf_slab *s, int on)
	{
		s->overlap = on && s->comm->world > 1 && !s->glf;
		return XF_OK;
	}
	int xf_slab_neighbours(const xf_slab *s, int lohi[2])
	{
		lohi[0] = s->lo, lohi[1] = s->hi;
		return XF_OK;
	}

	// the exchange proper on `st`: send_lo -> rank lo (lands in ITS zmax ghosts), send_hi -> rank hi (its zmin ghosts)
	static int exchange(xf_slab *s, cudaStream_t st)
	{
		if (s->lo < 0 && s->hi < 0)
			return XF_OK;
		NcclApi &n = nccl();
		NC(n.GroupStart());
		if (s->hi >= 0)
			NC(n.Send(s->send_hi, s->nhalo, ncclDouble, s->hi, s->comm->comm, st));
		if (s->lo >= 0)
			NC(n.Send(s->send_lo, s->nhalo, ncclDouble, s->lo, s->comm->comm, st));
		if (s->lo >= 0)
			NC(n.Recv(s->recv_lo, s->nhalo, ncclDouble, s->lo, s->comm->comm, st));
		if (s->hi >= 0)
			NC(n.Recv(s->recv_hi, s->nhalo, ncclDouble, s->hi, s->comm->comm, st));
		NC(n.GroupEnd());
		return XF_OK;
	}
	// z exchange of `field` on the compute stream: pack -> send/recv -> unpack (MpiTransBuf, mpiPacks.cpp:357-505)
	int xf_slab_halo(xf_slab *s, double *field)
	{
		int rc;
		if (s->lo >= 0)
			XS(xf_halo_pack(s->c, field, 4, s->send_lo));
		if (s->hi >= 0)
			XS(xf_halo_pack(s->c, field, 5, s->send_hi));
		if ((rc = exchange(s, s->main)))
			return rc;
		if (s->lo >= 0)
			XS(xf_halo_unpack(s->c, field, 4, s->recv_lo));
		if (s->hi >= 0)
			XS(xf_halo_unpack(s->c, field, 5, s->recv_hi));
		return XF_OK;
	}
	// main.cpp:44-48: BoundaryCondition + UpdateStates on the initial U
	int xf_slab_startup(xf_slab *s, double *U, int *error)
	{
		int rc;
		XS(xf_boundary(s->c, U, s->bc));
		if ((rc = xf_slab_halo(s, U)))
			return rc;
		XS(xf_update_states(s->c, U, error));
		if (s->glf && (rc = xf_comm_allreduce_max(s->comm, xf_device_glfmax(s->c), 9, s->main)))
			return rc;
		return XF_OK;
	}
	int xf_slab_stage(xf_slab *s, double *U, double *U1, double *LU, int flag)
	{
		if (flag < 1 || flag > 3)
			return sfail(XF_ERR_ARG, "flag must be 1..3");
		if ((s->lo >= 0 || s->hi >= 0) && xf_transport_needs_global_extrema(s->c))
			return sfail(XF_ERR_ARG, "species-diffusion limiter with domain-wide mass-fraction extrema (-diffu-mpi) is not reduced over the slabs: single GPU only");
		double *UI = flag == 1 ? U : U1;
		int rc;
		XS(xf_boundary(s->c, UI, s->bc));
		if (!s->overlap)
		{ // reference order: BC + exchange, UpdateStates, GetLU, UpdateU
			if ((rc = xf_slab_halo(s, UI)))
				return rc;
			if (s->glf)
			{
				XS(xf_stage_states(s->c, U, U1, flag));
				if ((rc = xf_comm_allreduce_max(s->comm, xf_device_glfmax(s->c), 9, s->main)))
					return rc;
				XS(xf_stage_fluxes(s->c, U, U1, LU, flag));
			}
			else
				XS(xf_rk_stage(s->c, U, U1, LU, nullptr, flag));
			return XF_OK;
		}
		if (s->lo >= 0)
			XS(xf_halo_pack(s->c, UI, 4, s->send_lo));
		if (s->hi >= 0)
			XS(xf_halo_pack(s->c, UI, 5, s->send_hi));
		CUS(cudaEventRecord(s->ev_packed, s->main));
		CUS(cudaStreamWaitEvent(s->comm_stream, s->ev_packed, 0));
		if ((rc = exchange(s, s->comm_stream)))
			return rc;
		if (s->lo >= 0)
			XS(xf_halo_unpack_on(s->c, UI, 4, s->recv_lo, s->comm_stream));
		if (s->hi >= 0)
			XS(xf_halo_unpack_on(s->c, UI, 5, s->recv_hi, s->comm_stream));
		CUS(cudaEventRecord(s->ev_unpacked, s->comm_stream));
		XS(xf_stage_interior(s->c, U, U1, LU, flag));
		CUS(cudaStreamWaitEvent(s->main, s->ev_unpacked, 0));
		XS(xf_stage_finish(s->c, U, U1, LU, flag));
		return XF_OK;
	}
	// one time step, nothing synchronises with the host: MAX of the dt maxima over the ranks (Fluids.cpp:902-913), device-resident
	// dt (XFLUIDS.cpp:196-199), three stages (XFLUIDS.cpp:441-525)
	int xf_slab_step(xf_slab *s, double *U, double *U1, double *LU, double t_end)
	{
		int rc;
		if ((rc = xf_comm_allreduce_max(s->comm, xf_device_dtmax(s->c), 3, s->main)))
			return rc;
		XS(xf_dt_device(s->c, t_end));
		for (int flag = 1; flag <= 3; flag++)
			if ((rc = xf_slab_stage(s, U, U1, LU, flag)))
				return rc;
		return XF_OK;
	}
	// error word MAX-reduced over the ranks (every rank takes the same decision); synchronises
	int xf_slab_any_error(xf_slab *s, int *error)
	{
		int rc;
		int *scratch = reinterpret_cast<int *>(s->d_scratch);
		CUS(cudaMemcpyAsync(scratch, xf_device_errors(s->c), 4 * sizeof(int), cudaMemcpyDeviceToDevice, s->main));
		if ((rc = xf_comm_allreduce_max_int(s->comm, scratch, 4, s->main)))
			return rc;
		CUS(cudaMemcpyAsync(s->h_err, scratch, 4 * sizeof(int), cudaMemcpyDeviceToHost, s->main));
		CUS(cudaStreamSynchronize(s->main));
		*error = (s->h_err[0] || s->h_err[1] || s->h_err[2]) ? 1 : 0;
		return XF_OK;
	}
	// nsteps steps of XFLUIDS::Evolution's inner loop on every rank (the multi-GPU form of xf_run): the host looks at (time, dt, error)
	// every step near t_end and every few steps otherwise; time and dt are identical on all ranks (max is exact), the error word is
	// MAX-reduced, so all ranks stop together.
	int xf_slab_run(xf_slab *s, double *U, double *U1, double *LU, int nsteps, double t_end, int *steps_done, double *time_out, int *error)
	{
		double t = 0, dt = 0;
		int done = 0, err = 0, rc;
		XS(xf_get_time(s->c, &t, nullptr));
		int poll = 1;
		while (done < nsteps && t < t_end)
		{
			const int batch = poll < nsteps - done ? poll : nsteps - done;
			for (int q = 0; q < batch; q++)
				if ((rc = xf_slab_step(s, U, U1, LU, t_end)))
					return rc;
			done += batch;
			XS(xf_get_time(s->c, &t, &dt));
			if ((rc = xf_slab_any_error(s, &err)))
				return rc;
			if (err)
				break;
			// dt may grow from step to step: never queue more steps than fit before t_end at four times the current dt
			const double room = dt > 0 ? (t_end - t) / (4.0 * dt) : 1.0;
			poll = room >= 16.0 ? 16 : (room >= 1.0 ? (int)room : 1);
		}
		if (steps_done)
			*steps_done = done;
		if (time_out)
			*time_out = t;
		if (error)
			*error = err;
		return err ? XF_ERR_NUMERIC : XF_OK;
	}
	// one step from / to this rank's pinned host buffer (AoS U of its slab) with the PCIe copies overlapped like xf_step_host and the
	// z-halo exchanges between the pieces; XF_ERR_ARG when the overlapped form does not apply (global Lax-Friedrichs, 1-D / 2-D)
	int xf_slab_step_host(xf_slab *s, double *h_U, double t_end, double *U, double *U1, double *LU, int *error)
	{
		int rc;
		if (s->glf)
			return sfail(XF_ERR_ARG, "xf_slab_step_host: not with global Lax-Friedrichs splitting");
		if ((rc = xf_comm_allreduce_max(s->comm, xf_device_dtmax(s->c), 3, s->main)))
			return rc;
		XS(xf_host_begin(s->c, h_U, s->bc, t_end, U, U1, LU));
		if ((rc = xf_slab_halo(s, U)))
			return rc;
		XS(xf_host_stage1_finish(s->c, s->bc, U, U1, LU));
		if ((rc = xf_slab_stage(s, U, U1, LU, 2)))
			return rc;
		XS(xf_boundary(s->c, U1, s->bc));
		if ((rc = xf_slab_halo(s, U1)))
			return rc;
		rc = xf_host_stage3(s->c, h_U, U, U1, LU, error);
		if (rc != XF_OK && rc != XF_ERR_NUMERIC)
			return sfail(rc, std::string("xf_host_stage3: ") + xf_last_error());
		return rc;
	}
	// MAX over the ranks of n (<= 8) HOST doubles (block-level GetFluidDt: the maxima have been read back already); synchronises
	int xf_slab_allreduce_max_host(xf_slab *s, double *h_v, int n)
	{
		if (n > 8)
			return sfail(XF_ERR_ARG, "at most 8 values");
		int rc;
		std::memcpy(s->h_pin, h_v, n * sizeof(double));
		CUS(cudaMemcpyAsync(s->d_scratch + 8, s->h_pin, n * sizeof(double), cudaMemcpyHostToDevice, s->main));
		if ((rc = xf_comm_allreduce_max(s->comm, s->d_scratch + 8, n, s->main)))
			return rc;
		CUS(cudaMemcpyAsync(s->h_pin, s->d_scratch + 8, n * sizeof(double), cudaMemcpyDeviceToHost, s->main));
		CUS(cudaStreamSynchronize(s->main));
		std::memcpy(h_v, s->h_pin, n * sizeof(double));
		return XF_OK;
	}
}
