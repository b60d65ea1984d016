// xf_capi.cu -- the C-ABI layer of libxfluids_b200.so (include/xfluids_b200.h): context, memory, dispatch to the
// strict / fast kernel flavours, CUDA-graph replay of whole time steps.  No CPU fallback anywhere: every entry
// point ends in a kernel launch or a CUDA memory operation on the context's device and stream.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>
#include "../../include/xfluids_b200.h"
#include "xf_launch.h"
#include "xf_log.cuh"
#include "xf_exp.cuh"

static thread_local std::string g_err;
static int fail(int code, const std::string &m)
{
	g_err = m;
	return code;
}
#define CU(call)                                                                                        \
	do                                                                                                  \
	{                                                                                                   \
		cudaError_t e__ = (call);                                                                       \
		if (e__ != cudaSuccess)                                                                         \
			return fail(XF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));              \
	} while (0)
#define XF_DEVICE(c) CU(cudaSetDevice((c)->device)) /* every entry point that launches or copies: several contexts / devices may live in one process */
#define KL(call)                                                                                        \
	do                                                                                                  \
	{                                                                                                   \
		int r__ = (call);                                                                               \
		if (r__ > 0)                                                                                    \
			return fail(XF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString((cudaError_t)r__)); \
		if (r__ < 0)                                                                                    \
			return fail(XF_ERR_ARG, std::string(#call) + ": unsupported configuration");                \
	} while (0)

struct XfTable
{
	decltype(&xf_strict::launch_prim) prim;
	decltype(&xf_strict::launch_sweeps) sweeps;
	decltype(&xf_strict::launch_lu) lu;
	decltype(&xf_strict::launch_rk) rk;
	decltype(&xf_strict::launch_nan) nan;
	decltype(&xf_strict::launch_bc) bc;
	decltype(&xf_strict::launch_dt) dt;
	decltype(&xf_strict::launch_dt_final) dt_final;
	decltype(&xf_strict::launch_layout) layout;
	decltype(&xf_strict::launch_scalar_pad) scalar_pad;
	decltype(&xf_strict::launch_halo) halo;
	decltype(&xf_strict::launch_sweep_x) sweep_x;
	decltype(&xf_strict::launch_march) march;
	decltype(&xf_strict::launch_visc) visc;
	decltype(&xf_strict::launch_prim_shell) prim_shell;
	decltype(&xf_strict::launch_rk_prim) rk_prim;
};
static const XfTable T_STRICT = {xf_strict::launch_prim, xf_strict::launch_sweeps, xf_strict::launch_lu, xf_strict::launch_rk, xf_strict::launch_nan,
								 xf_strict::launch_bc, xf_strict::launch_dt, xf_strict::launch_dt_final, xf_strict::launch_layout,
								 xf_strict::launch_scalar_pad, xf_strict::launch_halo, xf_strict::launch_sweep_x, xf_strict::launch_march, xf_strict::launch_visc,
								 xf_strict::launch_prim_shell, xf_strict::launch_rk_prim};
// fast mode: FMA contraction in the sweeps / LU / RK only.  Primitive recovery stays strict: its Newton loop stops on an
// absolute tolerance and is capped/limited, so a 1-ulp difference can change the trip count and move T by O(1e-6)
// (measured: 2e-6 relative on U after one jet step with a contracted prim kernel).
static const XfTable T_FAST = {xf_strict::launch_prim, xf_fast::launch_sweeps, xf_fast::launch_lu, xf_fast::launch_rk, xf_fast::launch_nan,
							   xf_fast::launch_bc, xf_fast::launch_dt, xf_fast::launch_dt_final, xf_fast::launch_layout,
							   xf_fast::launch_scalar_pad, xf_fast::launch_halo, xf_fast::launch_sweep_x, xf_fast::launch_march, xf_fast::launch_visc,
							   nullptr, nullptr}; // no update -> recovery fusion in fast mode: the recovery must not be compiled with FMA contraction

struct xf_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	xf_block bl{};
	xf_scheme sc{};
	int ns = 1, cop = 0, ghost = 0, E = 5;
	XfDev d{};
	XfThermo th{};
	XfVisc vs{};
	const XfTable *t = nullptr;
	std::vector<void *> owned;      // device allocations freed in xf_destroy
	double *stage = nullptr;        // device staging for AoS import/export [Ncells*E]
	cudaStream_t copy_stream = nullptr;   // chunked host upload of xf_step_host
	std::vector<cudaEvent_t> chunk_ev;
	int host_chunks = 32;           // z-chunks of the overlapped upload (xf_set_host_overlap; <= 1: plain path)
	double *h_pin = nullptr;        // pinned host scratch (16 doubles)
	int *h_err = nullptr;           // pinned host error word (4 ints)
	long long launches = 0;
	const double *lastUI = nullptr; // field of the last xf_update_states (its component 0 is rho)
	// CUDA graph of one time step (xf_run)
	// [0] a self-contained step; with the update -> recovery fusion also [2] the first step of a batch (leaves the deep cells of U
	// recovered for the next step), [3] a middle step (finds them recovered and leaves them so), [1] the last step of a batch
	cudaGraphExec_t gexec[4] = {nullptr, nullptr, nullptr, nullptr};
	const double *gU = nullptr, *gU1 = nullptr, *gLU = nullptr;
	int gbc[6] = {-1, -1, -1, -1, -1, -1};
	double gt_end = 0;
	long long glaunches[4] = {0, 0, 0, 0}; // kernel launches inside one replay
	// XF_FUSE_PRIM=1 (default 0): the stage update of stages 1 and 2 (and of stage 3 inside a batch of xf_run) goes straight on to the next
	// stage's primitive recovery for the cells no ghost fill reads (k_rk_prim); strict mode with the tiled sweeps only.  Bit-identical, and
	// measured SLOWER than the two kernels (SBI 512^3: 29.4 ms against 9.9 + 10.7 ms per stage, profiles/r02_tuning.md): at the 128
	// registers the recovery needs, 16 warps per SM cannot keep the update's 73 loads per cell in flight
	int fuse = 0;
	// 1 (default; XF_VISC_TAIL=0 switches it off): with the viscous terms on, every sweep of a whole-stage call subtracts the viscous wall
	// flux of its face before it stores, instead of a separate kernel that reads and rewrites the three wall-flux fields
	int visc_tail = 1;
	void drop_graphs()
	{
		for (int i = 0; i < 4; i++)
			if (gexec[i])
				cudaGraphExecDestroy(gexec[i]), gexec[i] = nullptr;
	}
	// TMA tensor maps of the marching sweeps: per (sweep input field, direction); the primitive / Y maps are per direction
	std::map<std::pair<const void *, int>, XfTma> tma;
	// 1 (default): tiled y / z sweeps store the wall fluxes, one kernel forms the divergence and the stage update.  0 (XF_MARCH=1): the
	// TMA-fed marching sweeps with the divergence / update fused in (xf_march.cuh) -- bit-identical results, measured slower on every
	// BASELINE config (profiles/r02_tuning.md), kept selectable for that A/B
	int tiled = 1;
	int gbc_stage[6] = {0, 0, 0, 0, 0, 0}; // face conditions of the stage in flight (the viscous block needs them for the derivative ghost fill)
	size_t ncells() const { return size_t(bl.Xmax) * bl.Ymax * bl.Zmax; }
};

static int dmalloc(xf_ctx *c, double **p, size_t n)
{
	CU(cudaMalloc((void **)p, n * sizeof(double)));
	CU(cudaMemsetAsync(*p, 0, n * sizeof(double), c->stream));
	c->owned.push_back(*p);
	return 0;
}


// ---- roofline denominators measured on the device the bench runs on (bench.py "roofline.peak") ----------------
// FP64: 8 independent DFMA chains per thread, 1024 threads/SM-slot, long enough to hide launch overhead.
__global__ void __launch_bounds__(256) k_peak_dfma(double *out, int iters, double a, double b)
{
	double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++)
	{
		x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
		x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
	}
	const double r = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
	if (r == 123.456)
		out[0] = r;
}
// FP64 issue-rate probes: 8 independent chains of one opcode per thread (OP 0 DADD, 1 DMUL, 2 DFMA); strict builds issue DADD / DMUL where
// a contracted build would issue DFMA, so the honest ceiling of the parity mode is the DADD / DMUL issue rate
template <int OP>
__global__ void __launch_bounds__(256) k_peak_op(double *out, int iters, double a, double b)
{
	double x[8];
#pragma unroll
	for (int q = 0; q < 8; q++)
		x[q] = threadIdx.x * 1e-9 + q;
	for (int i = 0; i < iters; i++)
	{
#pragma unroll
		for (int q = 0; q < 8; q++)
			x[q] = OP == 0 ? __dadd_rn(x[q], b) : (OP == 1 ? __dmul_rn(x[q], a) : fma(x[q], a, b));
	}
	double r = 0;
#pragma unroll
	for (int q = 0; q < 8; q++)
		r += x[q];
	if (r == 123.456)
		out[0] = r;
}
__global__ void __launch_bounds__(256) k_peak_copy(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		out[i] = in[i];
}

__global__ void __launch_bounds__(256) k_log_eval(const double *__restrict__ x, double *__restrict__ y, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		y[i] = xf_log(x[i]);
}

// which: 0 log(x), 1 exp(x), 2 pow(x, y)
__global__ void __launch_bounds__(256) k_math_eval(const double *__restrict__ x, const double *__restrict__ y2, double *__restrict__ y, size_t n, int which)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		y[i] = which == 0 ? xf_log(x[i]) : (which == 1 ? xf_exp(x[i]) : xf_pow(x[i], y2[i]));
}

extern "C"
{
	static int ensure_fw(xf_ctx *c);
	const char *xf_last_error(void) { return g_err.c_str(); }

	// y[i] = xf_log(x[i]) evaluated on `device` (host arrays): the parity test of the device logarithm against the host libm
	int xf_log_eval(int device, const double *h_x, double *h_y, size_t n)
	{
		CU(cudaSetDevice(device));
		double *dx = nullptr, *dy = nullptr;
		CU(cudaMalloc((void **)&dx, n * sizeof(double)));
		if (cudaMalloc((void **)&dy, n * sizeof(double)) != cudaSuccess)
		{
			cudaFree(dx);
			return fail(XF_ERR_CUDA, "xf_log_eval: cudaMalloc");
		}
		cudaError_t e = cudaMemcpy(dx, h_x, n * sizeof(double), cudaMemcpyHostToDevice);
		if (e == cudaSuccess)
		{
			k_log_eval<<<1184, 256>>>(dx, dy, n);
			e = cudaMemcpy(h_y, dy, n * sizeof(double), cudaMemcpyDeviceToHost);
		}
		cudaFree(dx), cudaFree(dy);
		if (e != cudaSuccess)
			return fail(XF_ERR_CUDA, std::string("xf_log_eval: ") + cudaGetErrorString(e));
		return XF_OK;
	}

	int xf_math_eval(int device, int which, const double *h_x, const double *h_y2, double *h_out, size_t n)
	{
		CU(cudaSetDevice(device));
		if (which < 0 || which > 2 || (which == 2 && !h_y2))
			return fail(XF_ERR_ARG, "xf_math_eval: which = 0 (log), 1 (exp), 2 (pow, needs the exponents)");
		double *dx = nullptr, *dy = nullptr, *d2 = nullptr;
		cudaError_t e = cudaMalloc((void **)&dx, n * sizeof(double));
		if (e == cudaSuccess)
			e = cudaMalloc((void **)&dy, n * sizeof(double));
		if (e == cudaSuccess && which == 2)
			e = cudaMalloc((void **)&d2, n * sizeof(double));
		if (e == cudaSuccess)
			e = cudaMemcpy(dx, h_x, n * sizeof(double), cudaMemcpyHostToDevice);
		if (e == cudaSuccess && which == 2)
			e = cudaMemcpy(d2, h_y2, n * sizeof(double), cudaMemcpyHostToDevice);
		if (e == cudaSuccess)
		{
			k_math_eval<<<1184, 256>>>(dx, d2, dy, n, which);
			e = cudaMemcpy(h_out, dy, n * sizeof(double), cudaMemcpyDeviceToHost);
		}
		cudaFree(dx), cudaFree(dy), cudaFree(d2);
		if (e != cudaSuccess)
			return fail(XF_ERR_CUDA, std::string("xf_math_eval: ") + cudaGetErrorString(e));
		return XF_OK;
	}

	int xf_create(const xf_block *bl, const xf_thermal *th, const xf_scheme *sc, int device, xf_ctx **out)
	{
		if (!bl || !th || !sc || !out)
			return fail(XF_ERR_ARG, "null argument");
		int ndev = 0;
		if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device)
			return fail(XF_ERR_CUDA, "no usable CUDA device (this library has no CPU path)");
		CU(cudaSetDevice(device));
		if (sc->weno_order != 5 && sc->weno_order != 6 && sc->weno_order != 7)
			return fail(XF_ERR_ARG, "weno_order must be 5 (WENO5-JS), 6 (WENO-CU6) or 7 (WENO7-JS)");
		if (sc->artificial_type < 1 || sc->artificial_type > 3)
			return fail(XF_ERR_ARG, "artificial_type must be 1 (ROE), 2 (LLF) or 3 (GLF)");
		if (th->cop && (th->num_species < 2 || th->num_species > 5))
			return fail(XF_ERR_ARG, "multi-component path is instantiated for 2..5 species");
		if (bl->Xmax < 1 || bl->Ymax < 1 || bl->Zmax < 1)
			return fail(XF_ERR_ARG, "bad block");
		const int minw = sc->weno_order == 7 ? 4 : 3;
		if ((bl->DimX && bl->Bwidth_X < minw) || (bl->DimY && bl->Bwidth_Y < minw) || (bl->DimZ && bl->Bwidth_Z < minw))
			return fail(XF_ERR_ARG, "ghost width too small for the stencil");
		xf_ctx *c = new xf_ctx();
		c->device = device;
		c->bl = *bl, c->sc = *sc;
		c->cop = th->cop ? 1 : 0, c->ns = th->cop ? th->num_species : 1, c->ghost = th->ghost_species ? 1 : 0;
		c->E = c->cop ? c->ns + 4 : 5;
		c->t = sc->fp_mode ? &T_FAST : &T_STRICT;
		XfDev &d = c->d;
		d.Xmax = bl->Xmax, d.Ymax = bl->Ymax, d.Zmax = bl->Zmax;
		d.Xp = (bl->Xmax + XF_PITCH_ALIGN - 1) / XF_PITCH_ALIGN * XF_PITCH_ALIGN;
		d.Xi = bl->X_inner, d.Yi = bl->Y_inner, d.Zi = bl->Z_inner;
		d.Bx = bl->Bwidth_X, d.By = bl->Bwidth_Y, d.Bz = bl->Bwidth_Z;
		d.DimX = bl->DimX, d.DimY = bl->DimY, d.DimZ = bl->DimZ;
		d.weno = sc->weno_order, d.alpha = sc->artificial_type, d.ghost = c->ghost, d.positivity = sc->positivity ? 1 : 0;
		d.sY = d.Xp, d.sZ = (long long)d.Xp * d.Ymax, d.N = d.sZ * d.Zmax;
		d._dx = bl->_dx, d._dy = bl->_dy, d._dz = bl->_dz, d.CFL = bl->CFLnumber, d.gamma0 = th->ncop_gamma;
		d.dx = bl->dx, d.dy = bl->dy, d.dz = bl->dz;
		// thermo tables (Thermo_device.h coefficient use; products formed exactly as the reference forms them)
		std::memset(&c->th, 0, sizeof(XfThermo));
		c->th.Ru = (6.02214076e26 * 1.380649e-23) * 1.0E-3; // global_setup.h:49-54
		if (c->cop)
		{
			const double _OT = 1.0 / 3.0;
			for (int n = 0; n < c->ns; n++)
			{
				for (int r = 0; r < 3; r++)
				{
					const double *a = th->Hia + n * 21;
					double *h = c->th.hcoef[r][n], *cc = c->th.ccoef[r][n];
					h[0] = -a[0 * 3 + r], h[1] = a[1 * 3 + r], h[2] = a[2 * 3 + r], h[3] = 0.5 * a[3 * 3 + r];
					h[4] = a[4 * 3 + r] * _OT, h[5] = 0.25 * a[5 * 3 + r], h[6] = 0.2 * a[6 * 3 + r], h[7] = th->Hib[n * 6 + 0 * 3 + r];
					for (int m = 0; m < 7; m++)
						cc[m] = a[m * 3 + r];
				}
				c->th.Ri[n] = th->Ri[n], c->th._Wi[n] = th->_Wi[n];
			}
		}
		const size_t N = (size_t)d.N;
		const char *march = std::getenv("XF_MARCH");
		c->tiled = (march && march[0] == '1') ? 0 : 1;
		const char *fuse = std::getenv("XF_FUSE_PRIM");
		c->fuse = (fuse && fuse[0] == '1') ? 1 : 0;
		const char *vtail = std::getenv("XF_VISC_TAIL");
		c->visc_tail = (vtail && vtail[0] == '0') ? 0 : 1;
		// every failure from here on goes through one exit that releases what has been allocated; the first failure wins
		int rc = 0;
		auto alloc = [&](double **p, size_t n)
		{ if (!rc) rc = dmalloc(c, p, n); };
		if (cudaDeviceGetAttribute(&d.nsm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
			rc = fail(XF_ERR_CUDA, "cudaDeviceGetAttribute");
		if (c->cop && N >= (size_t)0xffffffffu)
			rc = fail(XF_ERR_ARG, "block too large for 32-bit cell indices");
		double *prim5 = nullptr; // u, v, w, p, c contiguous: one 4-D TMA tensor map covers them (xf_march.cuh)
		alloc(&prim5, 5 * N);
		d.u = prim5, d.v = prim5 + N, d.w = prim5 + 2 * N, d.p = prim5 + 3 * N, d.c = prim5 + 4 * N;
		alloc(&d.H, N), alloc(&d.T, N);
		if (c->cop)
		{
			double **cop_arr[] = {&d.g3, &d.dpdrho, &d.e, &d.prho};
			for (double **p : cop_arr)
				alloc(p, N);
			alloc(&d.y, N * c->ns);
			alloc(&d.dpdrhoi, N * (c->ns - 1));
		}
		alloc(&d.red, XF_RED_COUNT);
		double *errp = nullptr;
		alloc(&errp, 2);
		d.err = reinterpret_cast<int *>(errp);
		if (c->cop)
		{ // hard-cell list of the two-pass primitive recovery: one 32-bit index per cell at most, + the counter
			double *hp = nullptr;
			alloc(&hp, N / 2 + 2);
			d.hard_count = reinterpret_cast<unsigned *>(hp);
			d.hard_ids = d.hard_count + 2;
		}
		if (!rc && c->tiled)
			rc = ensure_fw(c);
		if (!rc && cudaMallocHost((void **)&c->h_pin, 16 * sizeof(double)) != cudaSuccess)
			rc = fail(XF_ERR_CUDA, "cudaMallocHost");
		if (!rc && cudaMallocHost((void **)&c->h_err, 4 * sizeof(int)) != cudaSuccess)
			rc = fail(XF_ERR_CUDA, "cudaMallocHost");
		if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess)
			rc = fail(XF_ERR_CUDA, "cudaStreamSynchronize");
		if (rc)
		{
			const std::string keep = g_err;
			xf_destroy(c);
			g_err = keep;
			return rc;
		}
		*out = c;
		return XF_OK;
	}

	int xf_destroy(xf_ctx *c)
	{
		if (!c)
			return XF_OK;
		cudaSetDevice(c->device);
		cudaStreamSynchronize(c->stream);
		c->drop_graphs();
		for (void *p : c->owned)
			cudaFree(p);
		if (c->stage)
			cudaFree(c->stage);
		for (cudaEvent_t e : c->chunk_ev)
			cudaEventDestroy(e);
		if (c->copy_stream)
			cudaStreamDestroy(c->copy_stream);
		if (c->h_pin)
			cudaFreeHost(c->h_pin);
		if (c->h_err)
			cudaFreeHost(c->h_err);
		delete c;
		return XF_OK;
	}
	int xf_set_stream(xf_ctx *c, void *s)
	{
		c->stream = (cudaStream_t)s;
		c->drop_graphs();
		return XF_OK;
	}
	int xf_synchronize(xf_ctx *c)
	{
		XF_DEVICE(c);
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}
	size_t xf_pitch(const xf_ctx *c) { return (size_t)c->d.Xp; }
	size_t xf_field_stride(const xf_ctx *c) { return (size_t)c->d.N; }
	size_t xf_field_doubles(const xf_ctx *c) { return (size_t)c->d.N * c->E; }
	int xf_emax(const xf_ctx *c) { return c->E; }
	long long xf_launch_count(const xf_ctx *c) { return c->launches; }
	double *xf_device_dtmax(xf_ctx *c) { return c->d.red + XF_RED_DTMAX; }
	int *xf_device_errors(xf_ctx *c) { return c->d.err; }
	double *xf_device_glfmax(xf_ctx *c) { return c->d.red + XF_RED_GLF; }

	int xf_field_alloc(xf_ctx *c, double **p)
	{
		CU(cudaSetDevice(c->device));
		CU(cudaMalloc((void **)p, xf_field_doubles(c) * sizeof(double)));
		CU(cudaMemsetAsync(*p, 0, xf_field_doubles(c) * sizeof(double), c->stream));
		return XF_OK;
	}
	int xf_field_free(xf_ctx *c, double *p)
	{
		XF_DEVICE(c);
		CU(cudaStreamSynchronize(c->stream));
		CU(cudaFree(p));
		return XF_OK;
	}
	static int ensure_stage(xf_ctx *c)
	{
		if (!c->stage)
			CU(cudaMalloc((void **)&c->stage, c->ncells() * c->E * sizeof(double)));
		return 0;
	}
	int xf_upload_aos(xf_ctx *c, double *d_field, const double *h_aos)
	{
		XF_DEVICE(c);
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		CU(cudaMemcpyAsync(c->stage, h_aos, c->ncells() * c->E * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		KL(c->t->layout(c->d, c->E, d_field, c->stage, 1, c->stream, 0, -1));
		c->launches++;
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}
	int xf_download_aos(xf_ctx *c, const double *d_field, double *h_aos)
	{
		XF_DEVICE(c);
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		KL(c->t->layout(c->d, c->E, const_cast<double *>(d_field), c->stage, 0, c->stream, 0, -1));
		c->launches++;
		CU(cudaMemcpyAsync(h_aos, c->stage, c->ncells() * c->E * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}
	static double *scalar_by_name(xf_ctx *c, const char *name)
	{
		XfDev &d = c->d;
		if (!std::strcmp(name, "T")) return d.T;
		if (!std::strcmp(name, "p")) return d.p;
		if (!std::strcmp(name, "u")) return d.u;
		if (!std::strcmp(name, "v")) return d.v;
		if (!std::strcmp(name, "w")) return d.w;
		if (!std::strcmp(name, "H")) return d.H;
		if (!std::strcmp(name, "c")) return d.c;
		if (!std::strcmp(name, "rho") && c->lastUI) return const_cast<double *>(c->lastUI); // component 0 of the field the primitives were derived from
		if (!std::strncmp(name, "UI", 2) && name[2] >= '0' && name[2] < '0' + c->E && !name[3] && c->lastUI) // component n of that field
			return const_cast<double *>(c->lastUI) + (size_t)(name[2] - '0') * d.N;
		if (c->cop && name[0] == 'y' && name[1] >= '0' && name[1] < '0' + c->ns && !name[2])
			return d.y + (size_t)(name[1] - '0') * d.N;
		// viscous work arrays (after xf_set_transport): "visc", "therm", "Vde<0..8>", "Dkm<k>", "hi<k>"
		const XfVisc &v = c->vs;
		if (v.on && v.Vde)
		{
			if (!std::strcmp(name, "visc")) return v.va;
			if (!std::strcmp(name, "therm")) return v.tca;
			if (!std::strncmp(name, "Vde", 3) && name[3] >= '0' && name[3] <= '8' && !name[4]) return v.Vde + (size_t)(name[3] - '0') * d.N;
			if (!std::strncmp(name, "Dkm", 3) && name[3] >= '0' && name[3] < '0' + c->ns && !name[4]) return v.Dkm + (size_t)(name[3] - '0') * d.N;
			if (!std::strncmp(name, "hi", 2) && name[2] >= '0' && name[2] < '0' + c->ns && !name[3]) return v.hi + (size_t)(name[2] - '0') * d.N;
		}
		return nullptr;
	}
	int xf_set_scalar(xf_ctx *c, const char *name, const double *h)
	{
		XF_DEVICE(c);
		double *p = scalar_by_name(c, name);
		if (!p)
			return fail(XF_ERR_ARG, std::string("unknown scalar ") + name);
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		CU(cudaMemcpyAsync(c->stage, h, c->ncells() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		KL(c->t->scalar_pad(c->d, p, c->stage, 1, c->stream));
		c->launches++;
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}
	int xf_get_scalar(xf_ctx *c, const char *name, double *h)
	{
		XF_DEVICE(c);
		double *p = scalar_by_name(c, name);
		if (!p)
			return fail(XF_ERR_ARG, std::string("unknown scalar ") + name);
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		KL(c->t->scalar_pad(c->d, p, c->stage, 0, c->stream));
		c->launches++;
		CU(cudaMemcpyAsync(h, c->stage, c->ncells() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}
	int xf_get_wallflux_aos(xf_ctx *c, int dir, double *h_aos)
	{
		if (dir < 0 || dir > 2 || !c->d.Fw[dir])
			return fail(XF_ERR_ARG, "inactive direction");
		return xf_download_aos(c, c->d.Fw[dir], h_aos);
	}

	// ---- block-level entry points ---------------------------------------------------------------
	int xf_boundary(xf_ctx *c, double *U, const int bc[6])
	{
		XF_DEVICE(c);
		std::memcpy(c->gbc_stage, bc, 6 * sizeof(int));
		KL(c->t->bc(c->d, c->E, c->cop, U, bc, c->stream, &c->launches, 7, -1, -1));
		return XF_OK;
	}
	// planes [k0, k1) ; reset_dt: zero the dt maxima first (the first call of a gather)
	static int update_states_range(xf_ctx *c, double *U, bool gather_dt, bool reset_dt, int k0, int k1)
	{
		XF_DEVICE(c);
		int flags = 0;
		if (gather_dt)
		{
			if (reset_dt)
				CU(cudaMemsetAsync(c->d.red + XF_RED_DTMAX, 0, 3 * sizeof(double), c->stream));
			flags |= 1;
		}
		if (c->sc.artificial_type == 3 && c->sc.weno_order != 7)
			flags |= 2; // GLF running maxima; for SCHEME_ORDER 7 eigen_local == 0 and the maxima stay 0
		KL(c->t->prim(c->d, c->th, c->ns, c->cop, U, flags, c->stream, &c->launches, k0, k1));
		c->lastUI = U;
		return XF_OK;
	}
	static int update_states(xf_ctx *c, double *U, bool gather_dt) { return update_states_range(c, U, gather_dt, true, 0, c->d.Zmax); }
	// ---- update -> recovery fusion (k_rk_prim / k_prim_shell) ----
	static bool can_fuse(const xf_ctx *c) { return c->fuse && c->tiled && c->t->rk_prim != nullptr; }
	static int prim_flags(const xf_ctx *c, bool gather_dt) { return (gather_dt ? 1 : 0) | ((c->sc.artificial_type == 3 && c->sc.weno_order != 7) ? 2 : 0); }
	// the cells the previous stage's k_rk_prim left out: ghosts and the inner cells the ghost fill has just read.  The dt maxima were
	// reset ahead of that kernel, the list of unconverged cells holds its entries.
	static int update_states_shell(xf_ctx *c, double *U, bool gather_dt)
	{
		XF_DEVICE(c);
		KL(c->t->prim_shell(c->d, c->th, c->ns, c->cop, U, prim_flags(c, gather_dt), c->stream, &c->launches));
		c->lastUI = U;
		return XF_OK;
	}

	// ---- wall-flux fields of the block-level API (FluxFw / Gw / Hw): allocated on first use, the fused path never stores wall fluxes ----
	static int ensure_fw(xf_ctx *c)
	{
		XfDev &d = c->d;
		for (int dir = 0; dir < 3; dir++)
			if (((dir == 0 && d.DimX) || (dir == 1 && d.DimY) || (dir == 2 && d.DimZ)) && !d.Fw[dir])
			{
				int rc = dmalloc(c, &d.Fw[dir], (size_t)d.N * c->E);
				if (rc)
					return rc;
				c->drop_graphs(); // XfDev travels by value inside captured launches
			}
		return 0;
	}

	int xf_set_transport(xf_ctx *c, const xf_transport *tr)
	{
		XF_DEVICE(c);
		if (!tr)
			return fail(XF_ERR_ARG, "null transport");
		XfVisc &v = c->vs;
		const bool was_on = v.on != 0;
		double *keep[5] = {v.Vde, v.va, v.tca, v.Dkm, v.hi};
		double *keep_lim = v.lim;
		std::memset(&v, 0, sizeof(v));
		v.Vde = keep[0], v.va = keep[1], v.tca = keep[2], v.Dkm = keep[3], v.hi = keep[4], v.lim = keep_lim;
		v.on = tr->visc ? 1 : 0, v.heat = (tr->visc && tr->visc_heat) ? 1 : 0, v.diffu = (tr->visc && tr->visc_diffu) ? 1 : 0;
		c->drop_graphs(); // the parameter block travels by value inside captured launches
		if (!v.on)
			return XF_OK;
		if (!c->tiled)
			return fail(XF_ERR_ARG, "the viscous terms need the wall fluxes in memory: not with XF_MARCH=1");
		if (!tr->fit_visc || !tr->Wi || (v.heat && !tr->fit_therm) || (v.diffu && !tr->fit_Dkj))
			return fail(XF_ERR_ARG, "transport fits missing");
		const int ns = c->ns;
		for (int n = 0; n < ns; n++)
		{
			for (int m = 0; m < 4; m++)
			{
				v.fit_visc[n][m] = tr->fit_visc[n * 4 + m];
				if (v.heat)
					v.fit_therm[n][m] = tr->fit_therm[n * 4 + m];
			}
			v.Wi[n] = tr->Wi[n];
			if (v.diffu)
				for (int q = 0; q < ns; q++)
					for (int m = 0; m < 4; m++)
						v.fit_Dkj[n * ns + q][m] = tr->fit_Dkj[(n * ns + q) * 4 + m];
		}
		v.Yil_limiter = tr->Yil_limiter, v.Dim_limiter = tr->Dim_limiter, v.dim_max0 = tr->dim_max0;
		// the molar-mass factors of PHI are per-pair constants: evaluated here with the host's pow(), the function the reference calls
		v.dkj_sym = v.diffu ? 1 : 0;
		for (int k = 0; k < ns; k++)
			for (int i = 0; i < ns; i++)
			{
				v.phiW[k * ns + i] = std::pow(v.Wi[i] / v.Wi[k], 0.25);
				v.phiS[k * ns + i] = std::pow(1.0 + v.Wi[k] / v.Wi[i], -0.5);
				if (v.diffu && std::memcmp(v.fit_Dkj[i * ns + k], v.fit_Dkj[k * ns + i], 4 * sizeof(double)) != 0)
					v.dkj_sym = 0;
			}
		if (!was_on || !v.Vde)
		{
			const size_t N = (size_t)c->d.N;
			int rc;
			if ((!v.Vde && (rc = dmalloc(c, &v.Vde, 9 * N))) || (!v.va && (rc = dmalloc(c, &v.va, N))) || (!v.tca && (rc = dmalloc(c, &v.tca, N))) ||
				(!v.Dkm && (rc = dmalloc(c, &v.Dkm, N * ns))) || (!v.hi && (rc = dmalloc(c, &v.hi, N * ns))) || (!v.lim && (rc = dmalloc(c, &v.lim, 4 * XF_MAXS))))
				return rc;
		}
		return ensure_fw(c);
	}

	// 1 when the species-diffusion limiter depends on domain-wide extrema of the mass fractions (Dim_max != 0: the reference's MPI build,
	// ConVenction_block.hpp:489-503 reduces them over the ranks); a slab driver would have to MIN / MAX-reduce them every stage
	int xf_transport_needs_global_extrema(const xf_ctx *c) { return (c && c->vs.on && c->vs.diffu && c->vs.dim_max0 != 0.0) ? 1 : 0; }

	// ---- TMA tensor maps of the marching sweeps (cuTensorMapEncodeTiled through the runtime's driver entry point: no libcuda link) ----
	typedef CUresult (*xf_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
									 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	static xf_encode_fn tma_encoder()
	{
		static xf_encode_fn fn = nullptr;
		if (!fn)
		{
			void *p = nullptr;
			cudaDriverEntryPointQueryResult q;
			if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
				fn = (xf_encode_fn)p;
		}
		return fn;
	}
	// [ncomp][Zmax][Ymax][Xp] doubles at `base` as a 4-D tensor (x, y, z, component); box = XF_MW cells in x by TF cells along `dir`
	static int make_map(xf_ctx *c, CUtensorMap *m, const double *base, int ncomp, int dir, int rows = 0, int width = XF_MW)
	{
		const XfDev &d = c->d;
		xf_encode_fn enc = tma_encoder();
		if (!enc)
			return fail(XF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
		const cuuint32_t TF = rows > 0 ? (cuuint32_t)rows : (cuuint32_t)XF_MTF(d.weno);
		const cuuint64_t gdim[4] = {(cuuint64_t)d.Xp, (cuuint64_t)d.Ymax, (cuuint64_t)d.Zmax, (cuuint64_t)ncomp};
		const cuuint64_t gstr[3] = {(cuuint64_t)d.Xp * 8, (cuuint64_t)d.sZ * 8, (cuuint64_t)d.N * 8};
		const cuuint32_t box[4] = {(cuuint32_t)width, dir == 1 ? TF : 1u, dir == 2 ? TF : 1u, (cuuint32_t)ncomp};
		const cuuint32_t es[4] = {1, 1, 1, 1};
		const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double *>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
							   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS)
			return fail(XF_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
		return 0;
	}
	static int tma_for(xf_ctx *c, const double *UI, int dir, const XfTma **out)
	{
		const auto key = std::make_pair((const void *)UI, dir);
		auto it = c->tma.find(key);
		if (it == c->tma.end())
		{
			XfTma t;
			std::memset(&t, 0, sizeof(t));
			int rc;
			if ((rc = make_map(c, &t.U, UI, c->E, dir)) || (rc = make_map(c, &t.P, c->d.u, 5, dir)))
				return rc;
			if (c->cop && (rc = make_map(c, &t.Y, c->d.y, c->ns - 1, dir)))
				return rc;
			it = c->tma.emplace(key, t).first;
		}
		*out = &it->second;
		return 0;
	}

	// the conserved pencil of one tile of the tiled y / z sweeps: XF_TW_ cells x XF_TILE_ROWS rows x Emax components (cache key: dir + 8)
	static int tile_map_for(xf_ctx *c, const double *UI, int dir, const CUtensorMap **out)
	{
		const auto key = std::make_pair((const void *)UI, dir + 8);
		auto it = c->tma.find(key);
		if (it == c->tma.end())
		{
			XfTma t;
			std::memset(&t, 0, sizeof(t));
			int rc;
			if ((rc = make_map(c, &t.U, UI, c->E, dir, XF_TILE_ROWS(c->d.weno), XF_TW_)))
				return rc;
			it = c->tma.emplace(key, t).first;
		}
		*out = &it->second.U;
		return 0;
	}
	// tiled sweeps of `dirmask` (wall fluxes stored in Fw)
	// visc_tail: every sweep subtracts the viscous wall flux before it stores (the viscous prelude has run on this stage's primitives)
	static int tiled_sweeps(xf_ctx *c, const double *UI, int dirmask, int kp0, int kp1, int tz0, int tz1, bool visc_tail = false)
	{
		const CUtensorMap *tmy = nullptr, *tmz = nullptr;
		int rc;
		if ((dirmask & 2) && c->d.DimY && (rc = tile_map_for(c, UI, 1, &tmy)))
			return rc;
		if ((dirmask & 4) && c->d.DimZ && (rc = tile_map_for(c, UI, 2, &tmz)))
			return rc;
		const XfViscF vf = xf_visc_face_args(c->vs);
		KL(c->t->sweeps(c->d, c->ns, c->cop, UI, c->stream, &c->launches, dirmask, kp0, kp1, tz0, tz1, tmy, tmz, visc_tail ? &vf : nullptr));
		return XF_OK;
	}

	// A march shorter than a few waves of blocks (2-D grids: one column per XF_MW cells of x) is cut into segments along the sweep; each
	// segment restages NST - 1 rows and recomputes one seed face.  Segment length = m TF - 1 cells, i.e. m full iterations of faces.
	static void plan_segments(const xf_ctx *c, int ntr, int len, XfMarchArgs *a)
	{
		const XfDev &d = c->d;
		const int TF = XF_MTF(d.weno);
		const long long cols = (long long)((d.Xi + XF_MW - 1) / XF_MW) * ntr, target = 4LL * d.nsm;
		a->nseg = 1, a->seglen = len;
		if (cols >= target || len < 8 * TF)
			return;
		long long want = (target + cols - 1) / cols;
		if (want > len / (4 * TF))
			want = len / (4 * TF);
		if (want <= 1)
			return;
		const int m = (int)((len / want + 1 + TF - 1) / TF);
		a->seglen = m * TF - 1;
		a->nseg = (len + a->seglen - 1) / a->seglen;
	}

	// The sweeps of one RK stage.  dirmask: the directions to run in this call; the x and y sweeps cover the z-planes [kp0, kp1), the z
	// sweep the cells (planes) [ka, kb) (absolute indices; < 0: all inner planes).  Every direction adds its part of the divergence to LU in
	// the reference's x -> y -> z order (UpdateFluidLU, Reconstruction_kernels.hpp:201-234); with `finish` the last active direction goes on
	// to the NaN guard and the stage update in the same kernel, and neither wall fluxes nor LU reach HBM.
	// nflags >= 0 (tiled sweeps, strict mode): the update goes on to the NEXT stage's primitive recovery of the deep cells with these gather flags
	static int stage_sweeps(xf_ctx *c, double *U, double *U1, double *LU, int flag, int dirmask, int kp0, int kp1, int ka, int kb, bool finish, int nflags = -1)
	{
		XF_DEVICE(c);
		const XfDev &d = c->d;
		double *UI = flag == 1 ? U : U1;
		const bool allz = ka < 0;
		if (kp0 < 0)
			kp0 = d.Bz, kp1 = d.Bz + d.Zi;
		if (ka < 0)
			ka = d.Bz, kb = d.Bz + d.Zi;
		if (c->tiled)
		{ // round-1 path: tiled sweeps store the wall fluxes, one kernel forms the divergence and the update
			if (!allz)
				return fail(XF_ERR_ARG, "the tiled sweeps cover whole blocks in z only");
			int rc;
			// a whole stage in one call: the viscous prelude (derivatives, transport coefficients, limiter extrema) runs first and every sweep
			// subtracts the viscous flux of its face before storing -- Fw is written once.  Split stages (z-slab overlap: the prelude needs the
			// z-ghost primitives the x / y sweeps do not wait for) keep the stand-alone pass after the last sweep.
			const bool tail = c->vs.on && c->visc_tail && finish && dirmask == 7;
			if (tail)
				KL(c->t->visc(d, c->th, c->vs, c->ns, c->cop, UI, c->gbc_stage, c->stream, &c->launches, 1));
			if ((rc = tiled_sweeps(c, UI, dirmask, kp0, kp1, -1, -1, tail)))
				return rc;
			if (finish)
			{
				if (c->vs.on && !tail)
				{ // viscous wall fluxes are subtracted once every direction's inviscid wall flux is in memory (ConVenction_block.hpp:424-575)
					KL(c->t->visc(d, c->th, c->vs, c->ns, c->cop, UI, c->gbc_stage, c->stream, &c->launches, 3));
				}
				if (nflags >= 0)
				{
					if (nflags & 1)
						CU(cudaMemsetAsync(d.red + XF_RED_DTMAX, 0, 3 * sizeof(double), c->stream));
					KL(c->t->rk_prim(d, c->th, c->ns, c->cop, U, U1, d.red + XF_RED_DT, flag, 1, nflags, c->stream, &c->launches));
					return XF_OK;
				}
				KL(c->t->rk(d, c->E, U, U1, LU, 0.0, d.red + XF_RED_DT, flag, 1, 1, c->stream, -1, -1));
				c->launches++;
			}
			return XF_OK;
		}
		const int first_dir = d.DimX ? 0 : (d.DimY ? 1 : 2), last_dir = d.DimZ ? 2 : (d.DimY ? 1 : 0);
		XfMarchArgs a;
		std::memset(&a, 0, sizeof(a));
		a.flag = flag, a.guard = 1, a.U = U, a.U1 = U1, a.LU = LU, a.dt_dev = d.red + XF_RED_DT, a.nseg = 1;
		bool rk_pass = false;
		int rc;
		if ((dirmask & 1) && d.DimX)
		{
			a.mode = XF_MODE_ACC, a.first = 1, a.t0 = kp0, a.t1 = kp1;
			KL(c->t->sweep_x(d, c->ns, c->cop, UI, a, c->stream));
			c->launches++;
			rk_pass = finish && last_dir == 0; // 1-D: the chunks of the x sweep overlap, no in-place update there
		}
		if ((dirmask & 2) && d.DimY)
		{
			const XfTma *tm;
			if ((rc = tma_for(c, UI, 1, &tm)))
				return rc;
			a.mode = (finish && last_dir == 1) ? XF_MODE_RK : XF_MODE_ACC, a.first = first_dir == 1, a.t0 = kp0, a.t1 = kp1, a.ca = d.By, a.cb = d.By + d.Yi;
			plan_segments(c, kp1 - kp0, d.Yi, &a);
			if (a.mode == XF_MODE_RK && a.nseg > 1 && flag == 2)
				a.mode = XF_MODE_ACC, rk_pass = true; // stage 2 updates its own input: not across segment boundaries (xf_march.cuh)
			KL(c->t->march(d, c->ns, c->cop, *tm, UI, a, 1, c->stream));
			c->launches++;
		}
		if ((dirmask & 4) && d.DimZ)
		{
			const XfTma *tm;
			if ((rc = tma_for(c, UI, 2, &tm)))
				return rc;
			a.mode = finish ? XF_MODE_RK : XF_MODE_ACC, a.first = first_dir == 2, a.t0 = d.By, a.t1 = d.By + d.Yi, a.ca = ka, a.cb = kb, a.nseg = 1, a.seglen = kb - ka;
			KL(c->t->march(d, c->ns, c->cop, *tm, UI, a, 2, c->stream));
			c->launches++;
		}
		if (rk_pass)
		{
			KL(c->t->rk(d, c->E, U, U1, LU, 0.0, d.red + XF_RED_DT, flag, 1, 0, c->stream, -1, -1));
			c->launches++;
		}
		return XF_OK;
	}
	int xf_error_flags(xf_ctx *c, int flags[4])
	{
		XF_DEVICE(c);
		CU(cudaMemcpyAsync(c->h_err, c->d.err, 4 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		for (int i = 0; i < 4; i++)
			flags[i] = c->h_err[i];
		return XF_OK;
	}
	int xf_clear_errors(xf_ctx *c)
	{
		XF_DEVICE(c);
		CU(cudaMemsetAsync(c->d.err, 0, 4 * sizeof(int), c->stream));
		return XF_OK;
	}
	int xf_update_states(xf_ctx *c, double *U, int *error)
	{
		int rc = update_states(c, U, true);
		if (rc)
			return rc;
		if (error)
		{
			int f[4];
			if ((rc = xf_error_flags(c, f)))
				return rc;
			*error = (f[0] || f[1]) ? 1 : 0;
		}
		return XF_OK;
	}
	int xf_get_lu(xf_ctx *c, const double *U, double *LU)
	{
		XF_DEVICE(c);
		// block-level form: the wall fluxes of the three directions are stored (the reference's FluxFw / Gw / Hw, readable through
		// xf_get_wallflux_aos), then UpdateFluidLU forms the divergence
		int rc;
		if ((rc = ensure_fw(c)))
			return rc;
		const XfDev &d = c->d;
		if (c->tiled)
		{
			if ((rc = tiled_sweeps(c, U, 7, -1, -1, -1, -1)))
				return rc;
		}
		else
		{
			XfMarchArgs a;
			std::memset(&a, 0, sizeof(a));
			a.mode = XF_MODE_FW, a.nseg = 1, a.t0 = d.Bz, a.t1 = d.Bz + d.Zi;
			if (d.DimX)
			{
				KL(c->t->sweep_x(d, c->ns, c->cop, U, a, c->stream));
				c->launches++;
			}
			if (d.DimY)
			{
				const XfTma *tm;
				if ((rc = tma_for(c, U, 1, &tm)))
					return rc;
				a.Fw = d.Fw[1], a.ca = d.By, a.cb = d.By + d.Yi;
				plan_segments(c, d.Zi, d.Yi, &a);
				KL(c->t->march(d, c->ns, c->cop, *tm, U, a, 1, c->stream));
				c->launches++;
			}
			if (d.DimZ)
			{
				const XfTma *tm;
				if ((rc = tma_for(c, U, 2, &tm)))
					return rc;
				a.Fw = d.Fw[2], a.t0 = d.By, a.t1 = d.By + d.Yi, a.ca = d.Bz, a.cb = d.Bz + d.Zi, a.nseg = 1, a.seglen = d.Zi;
				KL(c->t->march(d, c->ns, c->cop, *tm, U, a, 2, c->stream));
				c->launches++;
			}
		}
		if (c->vs.on)
			KL(c->t->visc(d, c->th, c->vs, c->ns, c->cop, U, c->gbc_stage, c->stream, &c->launches, 3));
		KL(c->t->lu(c->d, c->E, LU, c->stream));
		c->launches++;
		return XF_OK;
	}
	int xf_estimate_nan(xf_ctx *c, const double *UI, const double *LU, int *error)
	{
		XF_DEVICE(c);
		KL(c->t->nan(c->d, c->E, UI, LU, c->stream));
		c->launches++;
		if (error)
		{
			int f[4], rc;
			if ((rc = xf_error_flags(c, f)))
				return rc;
			*error = f[2] ? 1 : 0;
		}
		return XF_OK;
	}
	int xf_update_u_rk3(xf_ctx *c, double *U, double *U1, const double *LU, double dt, int flag)
	{
		XF_DEVICE(c);
		if (flag < 1 || flag > 3)
			return fail(XF_ERR_ARG, "flag must be 1..3");
		KL(c->t->rk(c->d, c->E, U, U1, LU, dt, nullptr, flag, 0, 0, c->stream, -1, -1));
		c->launches++;
		return XF_OK;
	}
	int xf_get_dt(xf_ctx *c, double *dt, double uvw_c_max[3])
	{
		XF_DEVICE(c);
		// stand-alone reduction pass over the stored primitives, like the reference's GetDt; rho is component 0 of
		// the field the primitives were last derived from (the reference keeps a separate rho array)
		if (!c->lastUI)
			return fail(XF_ERR_ARG, "xf_get_dt before any xf_update_states");
		CU(cudaMemsetAsync(c->d.red + XF_RED_DTMAX, 0, 3 * sizeof(double), c->stream));
		KL(c->t->dt(c->d, c->lastUI, c->stream));
		c->launches++;
		CU(cudaMemcpyAsync(c->h_pin, c->d.red + XF_RED_DTMAX, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		// uvw_c_max as this GetDt leaves it is what the positivity-preserving limiter of the following GetLU calls reads
		CU(cudaMemcpyAsync(c->d.red + XF_RED_PPL, c->d.red + XF_RED_DTMAX, 3 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		const double *m = c->h_pin;
		if (uvw_c_max)
			uvw_c_max[0] = m[0], uvw_c_max[1] = m[1], uvw_c_max[2] = m[2];
		const double dtref = m[0] * c->bl._dx + m[1] * c->bl._dy + m[2] * c->bl._dz;
		*dt = c->bl.CFLnumber / dtref;
		return XF_OK;
	}

	// ---- fused path --------------------------------------------------------------------------------
	int xf_dt_device(xf_ctx *c, double t_end)
	{
		XF_DEVICE(c);
		KL(c->t->dt_final(c->d, t_end, c->stream));
		c->launches++;
		return XF_OK;
	}
	// deep_valid: the previous stage's k_rk_prim has recovered the deep cells of this stage's input; fuse_next: this stage's update does the
	// same for the next stage (whose input is U1 after stages 1 and 2, U after stage 3: stage 1 of the next step)
	static int rk_stage_ex(xf_ctx *c, double *U, double *U1, double *LU, const int bc[6], int flag, bool deep_valid, bool fuse_next)
	{
		if (flag < 1 || flag > 3)
			return fail(XF_ERR_ARG, "flag must be 1..3");
		double *UI = flag == 1 ? U : U1;
		int rc;
		if (bc && (rc = xf_boundary(c, UI, bc)))
			return rc;
		// the dt of the NEXT step is computed from the primitives of stage 3 (XFLUIDS.cpp:196 reads fdata as left
		// by the last UpdateStates) -> gather the maxima there
		if ((rc = deep_valid ? update_states_shell(c, UI, flag == 3) : update_states(c, UI, flag == 3)))
			return rc;
		// x, y sweeps accumulate their part of the divergence in LU; the last direction adds its own and applies NaN guard + RK update
		return stage_sweeps(c, U, U1, LU, flag, 7, -1, -1, -1, -1, true, fuse_next ? prim_flags(c, flag == 2) : -1);
	}
	int xf_rk_stage(xf_ctx *c, double *U, double *U1, double *LU, const int bc[6], int flag) { return rk_stage_ex(c, U, U1, LU, bc, flag, false, false); }
	// ---- one stage split around the z-halo exchange (multi-GPU overlap; include/xfluids_b200.h) --------------
	int xf_stage_interior(xf_ctx *c, double *U, double *U1, double *LU, int flag)
	{
		if (flag < 1 || flag > 3 || !c->d.DimZ)
			return fail(XF_ERR_ARG, "xf_stage_interior: flag 1..3 and an active z dimension");
		double *UI = flag == 1 ? U : U1;
		int rc;
		if ((rc = update_states_range(c, UI, flag == 3, true, c->d.Bz, c->d.Zmax - c->d.Bz)))
			return rc;
		return stage_sweeps(c, U, U1, LU, flag, 3, -1, -1, -1, -1, false);
	}
	int xf_stage_finish(xf_ctx *c, double *U, double *U1, double *LU, int flag)
	{
		if (flag < 1 || flag > 3 || !c->d.DimZ)
			return fail(XF_ERR_ARG, "xf_stage_finish: flag 1..3 and an active z dimension");
		double *UI = flag == 1 ? U : U1;
		int rc;
		if ((rc = update_states_range(c, UI, flag == 3, false, 0, c->d.Bz)))
			return rc;
		if ((rc = update_states_range(c, UI, flag == 3, false, c->d.Zmax - c->d.Bz, c->d.Zmax)))
			return rc;
		return stage_sweeps(c, U, U1, LU, flag, 4, -1, -1, -1, -1, true);
	}
	// ---- one stage split between primitive recovery and sweeps (multi-GPU GLF: the 9 running maxima of |lambda| are MAX-reduced over
	//      the ranks in between, like the reference's MPI build does for eigen_block) ----------------------------------------------
	int xf_stage_states(xf_ctx *c, double *U, double *U1, int flag)
	{
		if (flag < 1 || flag > 3)
			return fail(XF_ERR_ARG, "flag must be 1..3");
		return update_states(c, flag == 1 ? U : U1, flag == 3);
	}
	int xf_stage_fluxes(xf_ctx *c, double *U, double *U1, double *LU, int flag)
	{
		if (flag < 1 || flag > 3)
			return fail(XF_ERR_ARG, "flag must be 1..3");
		return stage_sweeps(c, U, U1, LU, flag, 7, -1, -1, -1, -1, true);
	}
	int xf_get_time(xf_ctx *c, double *time, double *last_dt)
	{
		XF_DEVICE(c);
		CU(cudaMemcpyAsync(c->h_pin, c->d.red + XF_RED_DT, 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		if (last_dt)
			*last_dt = c->h_pin[0];
		if (time)
			*time = c->h_pin[1];
		return XF_OK;
	}
	// time steps taken with dt > 0 since the context was created (device counter; call after xf_get_time, which fetched it)
	static long long steps_taken(const xf_ctx *c) { return (long long)c->h_pin[2]; }
	int xf_set_time(xf_ctx *c, double time)
	{
		XF_DEVICE(c);
		c->h_pin[8] = time;
		CU(cudaMemcpyAsync(c->d.red + XF_RED_TIME, c->h_pin + 8, sizeof(double), cudaMemcpyHostToDevice, c->stream));
		CU(cudaStreamSynchronize(c->stream));
		return XF_OK;
	}

	// first_valid: the deep cells of U are already recovered (by the previous step of the batch); fuse_last: leave them so for the next
	static int enqueue_step(xf_ctx *c, double *U, double *U1, double *LU, const int bc[6], double t_end, bool first_valid = false, bool fuse_last = false)
	{
		int rc;
		const bool fuse = can_fuse(c);
		if ((rc = xf_dt_device(c, t_end)))
			return rc;
		for (int flag = 1; flag <= 3; flag++)
			if ((rc = rk_stage_ex(c, U, U1, LU, bc, flag, fuse && (flag > 1 || first_valid), fuse && (flag < 3 || fuse_last))))
				return rc;
		return XF_OK;
	}
	static int capture_step(xf_ctx *c, int which, double *U, double *U1, double *LU, const int bc[6], double t_end)
	{
		cudaStream_t cap = c->stream;
		cudaStream_t own = nullptr;
		if (cap == nullptr)
		{ // the legacy default stream cannot be captured
			CU(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
			CU(cudaStreamSynchronize(nullptr));
			c->stream = own;
		}
		const long long l0 = c->launches;
		CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
		int rc = enqueue_step(c, U, U1, LU, bc, t_end, (which & 1) != 0, (which & 2) != 0);
		cudaGraph_t g = nullptr;
		cudaError_t e = cudaStreamEndCapture(c->stream, &g);
		c->glaunches[which] = c->launches - l0;
		c->launches = l0;
		if (own)
			c->stream = cap;
		if (rc || e != cudaSuccess)
		{
			if (own)
				cudaStreamDestroy(own);
			return rc ? rc : fail(XF_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
		}
		CU(cudaGraphInstantiate(&c->gexec[which], g, 0));
		CU(cudaGraphDestroy(g));
		if (own)
			CU(cudaStreamDestroy(own));
		return XF_OK;
	}
	int xf_run(xf_ctx *c, double *U, double *U1, double *LU, const int bc[6], int nsteps, double t_end, int *steps_done, double *time_out, int *error)
	{
		CU(cudaSetDevice(c->device));
		// (re)capture one step when pointers / BCs / t_end change
		const bool same = c->gU == U && c->gU1 == U1 && c->gLU == LU && c->gt_end == t_end && !std::memcmp(c->gbc, bc, 6 * sizeof(int));
		if (!same)
		{
			c->drop_graphs();
			c->gU = U, c->gU1 = U1, c->gLU = LU, c->gt_end = t_end;
			std::memcpy(c->gbc, bc, 6 * sizeof(int));
		}
		const bool fuse = can_fuse(c);
		double t = 0, dt = 0;
		int rc, queued = 0, err = 0;
		auto replay = [&](int which) -> int
		{
			if (!c->gexec[which] && (rc = capture_step(c, which, U, U1, LU, bc, t_end)))
				return rc;
			CU(cudaGraphLaunch(c->gexec[which], c->stream));
			c->launches += c->glaunches[which];
			return XF_OK;
		};
		if ((rc = xf_get_time(c, &t, nullptr)))
			return rc;
		const long long steps0 = steps_taken(c);
		int poll = 1; // steps between host checks of (time, error)
		while (queued < nsteps && t < t_end)
		{
			int batch = poll < nsteps - queued ? poll : nsteps - queued;
			// every batch starts from and ends in the state the reference's loop has between steps (U as the stage-3 update left it):
			// only inside a batch does the stage-3 update go on to the next step's recovery
			for (int s = 0; s < batch; s++)
				if ((rc = replay((fuse && batch > 1) ? (s == 0 ? 2 : (s == batch - 1 ? 1 : 3)) : 0)))
					return rc;
			queued += batch;
			int f[4];
			if ((rc = xf_get_time(c, &t, &dt)) || (rc = xf_error_flags(c, f)))
				return rc;
			if (f[0] || f[1] || f[2])
			{
				err = 1;
				break;
			}
			// far from t_end: check less often; near it: every step, like the reference's loop (XFLUIDS.cpp:172-199).  dt may GROW from
			// step to step (start-up transients): never queue more steps than fit before t_end at four times the current dt, so that no
			// replay runs with dt clipped to 0 (such a step would still renormalise species and re-round U -- ADVICE r1)
			const double room = dt > 0 ? (t_end - t) / (4.0 * dt) : 1.0;
			poll = room >= 16.0 ? 16 : (room >= 1.0 ? (int)room : 1);
		}
		// steps actually taken (dt > 0), counted on the device
		const int done = (int)(steps_taken(c) - steps0);
		if (steps_done)
			*steps_done = done;
		if (time_out)
			*time_out = t;
		if (error)
			*error = err;
		return err ? XF_ERR_NUMERIC : XF_OK;
	}

	// ---- per-kernel timing of one eager step (bench.py roofline; CUDA events on the launching stream) ----
	// ms[0] dt, ms[1] boundary fill, ms[2] primitive recovery, ms[3..5] sweep x/y/z, ms[6] divergence + RK update,
	// each summed over the three stages of one time step; ms[7] = whole step.
	int xf_profile_step(xf_ctx *c, double *U, double *U1, double *LU, const int bc[6], double t_end, float ms[8])
	{
		CU(cudaSetDevice(c->device));
		const int NE = 1 + 3 * 6 + 1;
		struct Events
		{ // destroyed on every exit path
			cudaEvent_t ev[NE];
			int n = 0;
			~Events()
			{
				for (int i = 0; i < n; i++)
					cudaEventDestroy(ev[i]);
			}
		} evs;
		cudaEvent_t *ev = evs.ev;
		for (int i = 0; i < NE; i++)
		{
			CU(cudaEventCreate(&ev[i]));
			evs.n = i + 1;
		}
		int e = 0, rc;
		const bool fuse = can_fuse(c);
		CU(cudaEventRecord(ev[e++], c->stream));
		if ((rc = xf_dt_device(c, t_end)))
			return rc;
		for (int flag = 1; flag <= 3; flag++)
		{
			double *UI = flag == 1 ? U : U1;
			CU(cudaEventRecord(ev[e++], c->stream));
			if ((rc = xf_boundary(c, UI, bc)))
				return rc;
			CU(cudaEventRecord(ev[e++], c->stream));
			if ((rc = (fuse && flag > 1) ? update_states_shell(c, UI, flag == 3) : update_states(c, UI, flag == 3)))
				return rc;
			const int last_dir = c->d.DimZ ? 2 : (c->d.DimY ? 1 : 0);
			const bool tail = c->tiled && c->vs.on && c->visc_tail;
			if (tail) // (lands in the "prim" bucket ms[2]; the wall-flux part is inside the sweeps' buckets)
				KL(c->t->visc(c->d, c->th, c->vs, c->ns, c->cop, UI, c->gbc_stage, c->stream, &c->launches, 1));
			for (int dir = 0; dir < 3; dir++)
			{
				CU(cudaEventRecord(ev[e++], c->stream));
				if ((rc = c->tiled ? tiled_sweeps(c, UI, 1 << dir, -1, -1, -1, -1, tail) : stage_sweeps(c, U, U1, LU, flag, 1 << dir, -1, -1, -1, -1, dir == last_dir)))
					return rc;
			}
			CU(cudaEventRecord(ev[e++], c->stream));
			if (c->tiled)
			{
				if (c->vs.on && !tail)
					KL(c->t->visc(c->d, c->th, c->vs, c->ns, c->cop, UI, c->gbc_stage, c->stream, &c->launches, 3));
				if (fuse && flag < 3)
				{ // the step as xf_run's self-contained graph runs it: ms[6] holds the deep cells' recovery of stages 2 and 3, ms[2] the rest
					const int nflags = prim_flags(c, flag == 2);
					if (nflags & 1)
						CU(cudaMemsetAsync(c->d.red + XF_RED_DTMAX, 0, 3 * sizeof(double), c->stream));
					KL(c->t->rk_prim(c->d, c->th, c->ns, c->cop, U, U1, c->d.red + XF_RED_DT, flag, 1, nflags, c->stream, &c->launches));
				}
				else
				{
					KL(c->t->rk(c->d, c->E, U, U1, LU, 0.0, c->d.red + XF_RED_DT, flag, 1, 1, c->stream, -1, -1));
					c->launches++;
				}
			}
		}
		CU(cudaEventRecord(ev[e++], c->stream));
		CU(cudaStreamSynchronize(c->stream));
		for (int i = 0; i < 8; i++)
			ms[i] = 0.f;
		float t;
		CU(cudaEventElapsedTime(&t, ev[0], ev[1]));
		ms[0] = t;
		for (int st = 0; st < 3; st++)
			for (int q = 0; q < 6; q++)
			{
				CU(cudaEventElapsedTime(&t, ev[1 + st * 6 + q], ev[2 + st * 6 + q]));
				ms[1 + q] += t;
			}
		CU(cudaEventElapsedTime(&t, ev[0], ev[NE - 1]));
		ms[7] = t;
		return XF_OK;
	}

	// FP64 FMA rate (TFLOP/s, FMA = 2 flop) and device copy bandwidth (GB/s, read + write) of `device`, best of 5.
	int xf_measure_peaks(int device, double *dfma_tflops, double *copy_gbs)
	{
		CU(cudaSetDevice(device));
		cudaDeviceProp pr;
		CU(cudaGetDeviceProperties(&pr, device));
		cudaEvent_t a, b;
		CU(cudaEventCreate(&a));
		CU(cudaEventCreate(&b));
		double *out = nullptr;
		CU(cudaMalloc((void **)&out, 64));
		const int iters = 1 << 14, blocks = pr.multiProcessorCount * 8;
		double best = 0;
		for (int rep = 0; rep < 6; rep++)
		{
			CU(cudaEventRecord(a, 0));
			k_peak_dfma<<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
			CU(cudaEventRecord(b, 0));
			CU(cudaEventSynchronize(b));
			float ms;
			CU(cudaEventElapsedTime(&ms, a, b));
			const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
			if (rep && tf > best)
				best = tf;
		}
		if (dfma_tflops)
			*dfma_tflops = best;
		if (copy_gbs)
		{
			const size_t n = (size_t)1 << 27; // 2 GiB per buffer as double2
			double2 *p = nullptr, *q = nullptr;
			CU(cudaMalloc((void **)&p, n * sizeof(double2)));
			CU(cudaMalloc((void **)&q, n * sizeof(double2)));
			CU(cudaMemset(p, 0, n * sizeof(double2)));
			double bw = 0;
			for (int rep = 0; rep < 6; rep++)
			{
				CU(cudaEventRecord(a, 0));
				k_peak_copy<<<pr.multiProcessorCount * 16, 256>>>(p, q, n);
				CU(cudaEventRecord(b, 0));
				CU(cudaEventSynchronize(b));
				float ms;
				CU(cudaEventElapsedTime(&ms, a, b));
				const double g = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
				if (rep && g > bw)
					bw = g;
			}
			*copy_gbs = bw;
			cudaFree(p), cudaFree(q);
		}
		cudaFree(out);
		cudaEventDestroy(a), cudaEventDestroy(b);
		return XF_OK;
	}

	// achieved host <-> device copy rates (GB/s) of this context's device for `bytes` of pinned host memory, through the AoS staging
	// buffer: the denominators of the end-to-end leg (what the PCIe link of this rank delivers when nothing else runs)
	int xf_measure_pcie(xf_ctx *c, void *h_pinned, size_t bytes, double *h2d_gbs, double *d2h_gbs)
	{
		CU(cudaSetDevice(c->device));
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		const size_t cap = c->ncells() * c->E * sizeof(double);
		if (bytes > cap)
			bytes = cap;
		cudaEvent_t a, b;
		CU(cudaEventCreate(&a));
		CU(cudaEventCreate(&b));
		float ms = 0;
		int rc = XF_OK;
		for (int dir = 0; dir < 2 && rc == XF_OK; dir++)
		{
			cudaError_t e = cudaEventRecord(a, c->stream);
			if (e == cudaSuccess)
				e = dir == 0 ? cudaMemcpyAsync(c->stage, h_pinned, bytes, cudaMemcpyHostToDevice, c->stream) : cudaMemcpyAsync(h_pinned, c->stage, bytes, cudaMemcpyDeviceToHost, c->stream);
			if (e == cudaSuccess)
				e = cudaEventRecord(b, c->stream);
			if (e == cudaSuccess)
				e = cudaEventSynchronize(b);
			if (e == cudaSuccess)
				e = cudaEventElapsedTime(&ms, a, b);
			if (e != cudaSuccess)
				rc = fail(XF_ERR_CUDA, std::string("xf_measure_pcie: ") + cudaGetErrorString(e));
			else
				*(dir == 0 ? h2d_gbs : d2h_gbs) = bytes / (ms * 1e-3) / 1e9;
		}
		cudaEventDestroy(a), cudaEventDestroy(b);
		return rc;
	}

	// FP64 instruction issue rates of `device` in 1e12 thread-instructions per second: [0] DADD, [1] DMUL, [2] DFMA (best of 5 each)
	int xf_measure_fp64_issue(int device, double tinst[3])
	{
		CU(cudaSetDevice(device));
		cudaDeviceProp pr;
		CU(cudaGetDeviceProperties(&pr, device));
		cudaEvent_t a, b;
		CU(cudaEventCreate(&a));
		CU(cudaEventCreate(&b));
		double *out = nullptr;
		CU(cudaMalloc((void **)&out, 64));
		const int iters = 1 << 14, blocks = pr.multiProcessorCount * 8;
		for (int op = 0; op < 3; op++)
		{
			double best = 0;
			for (int rep = 0; rep < 6; rep++)
			{
				CU(cudaEventRecord(a, 0));
				if (op == 0)
					k_peak_op<0><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
				else if (op == 1)
					k_peak_op<1><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
				else
					k_peak_op<2><<<blocks, 256>>>(out, iters, 0.999999, 1e-7);
				CU(cudaEventRecord(b, 0));
				CU(cudaEventSynchronize(b));
				float ms;
				CU(cudaEventElapsedTime(&ms, a, b));
				const double t = 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
				if (rep && t > best)
					best = t;
			}
			tinst[op] = best;
		}
		cudaFree(out);
		cudaEventDestroy(a), cudaEventDestroy(b);
		return XF_OK;
	}

	// ---- halo -----------------------------------------------------------------------------------
	size_t xf_halo_doubles(const xf_ctx *c) { return (size_t)c->E * c->d.Bz * (size_t)c->d.sZ; }
	int xf_halo_pack_on(xf_ctx *c, const double *U, int face, double *buf, void *stream)
	{
		cudaStream_t keep = c->stream;
		c->stream = (cudaStream_t)stream;
		const int rc = xf_halo_pack(c, U, face, buf);
		c->stream = keep;
		return rc;
	}
	int xf_halo_unpack_on(xf_ctx *c, double *U, int face, const double *buf, void *stream)
	{
		cudaStream_t keep = c->stream;
		c->stream = (cudaStream_t)stream;
		const int rc = xf_halo_unpack(c, U, face, buf);
		c->stream = keep;
		return rc;
	}
	int xf_halo_pack(xf_ctx *c, const double *U, int face, double *buf)
	{
		XF_DEVICE(c);
		if (!c->d.DimZ || (face != 4 && face != 5))
			return fail(XF_ERR_ARG, "halo faces are 4 (zmin) / 5 (zmax) of an active z dimension");
		const int k0 = face == 4 ? c->d.Bz : c->d.Zmax - 2 * c->d.Bz;
		KL(c->t->halo(c->d, c->E, const_cast<double *>(U), buf, k0, 1, c->stream));
		c->launches++;
		return XF_OK;
	}
	int xf_halo_unpack(xf_ctx *c, double *U, int face, const double *buf)
	{
		XF_DEVICE(c);
		if (!c->d.DimZ || (face != 4 && face != 5))
			return fail(XF_ERR_ARG, "halo faces are 4 (zmin) / 5 (zmax) of an active z dimension");
		const int k0 = face == 4 ? 0 : c->d.Zmax - c->d.Bz;
		KL(c->t->halo(c->d, c->E, U, const_cast<double *>(buf), k0, 0, c->stream));
		c->launches++;
		return XF_OK;
	}

	// ---- host-buffer step (end-to-end timing path) ----------------------------------------------------
	void *xf_host_alloc_pinned(size_t bytes)
	{
		void *p = nullptr;
		return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
	}
	void xf_host_free_pinned(void *p) { cudaFreeHost(p); }
	int xf_set_host_overlap(xf_ctx *c, int chunks)
	{
		c->host_chunks = chunks < 0 ? 0 : chunks;
		return XF_OK;
	}
	// One step from a host buffer with the upload overlapped: the AoS image goes up in z-chunks on a copy stream while the compute
	// stream converts each arrived chunk to the SoA layout and runs the plane-local part of stage 1 on it (x / y ghost fill,
	// primitive recovery, x and y sweeps).  The planes within 2 Bz of the z faces wait for the z ghost fill, which must read the
	// conserved variables BEFORE the primitive recovery renormalises the species (the same ordering rule as the z-halo split).
	// Every cell goes through the same kernels on the same inputs as in the plain path: the result is bit-identical.
	// ---- the overlapped host step in three pieces, so that a multi-GPU caller can put its z-halo exchanges between them -------------
	//   xf_host_begin          chunked upload + the plane-local part of stage 1 on the planes that do not feed the z ghost fill / halo
	//   [z-halo exchange of U by the caller: the planes it packs, Bz..2Bz-1 and their mirror, are still un-renormalised]
	//   xf_host_stage1_finish  z ghost fill, the waiting planes, z sweep, update
	//   [stage 2 by the caller; then ghost fill + z-halo exchange of U1]
	//   xf_host_stage3         primitive recovery of U1 (gathers the next dt), then sweeps / update / SoA->AoS / download per z-chunk
	static bool host_overlap_ok(xf_ctx *c)
	{
		const XfDev &d = c->d;
		return !c->vs.on && c->host_chunks > 1 && d.DimX && d.DimY && d.DimZ && d.Zmax >= 6 * d.Bz && d.Zmax >= c->host_chunks && c->sc.artificial_type != 3;
	}
	int xf_host_begin(xf_ctx *c, double *h_U, const int bc[6], double t_end, double *U, double *U1, double *LU)
	{
		if (!host_overlap_ok(c))
			return fail(XF_ERR_ARG, "xf_host_begin: needs a 3-D block, host_chunks > 1 and ROE / LLF splitting");
		const XfDev &d = c->d;
		const int Bz = d.Bz, Zmax = d.Zmax;
		int rc;
		if (ensure_stage(c))
			return XF_ERR_CUDA;
		if (!c->copy_stream)
			CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
		const int nch = c->host_chunks;
		if ((int)c->chunk_ev.size() < nch + 1)
		{
			const size_t old = c->chunk_ev.size();
			c->chunk_ev.resize(nch + 1);
			for (size_t i = old; i < c->chunk_ev.size(); i++)
				CU(cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming));
		}
		const size_t plane_aos = (size_t)d.Xmax * d.Ymax * c->E; // doubles per z-plane of the AoS image
		// the copy stream must not overwrite the staging buffer while an earlier call still reads it
		CU(cudaEventRecord(c->chunk_ev[nch], c->stream));
		CU(cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[nch], 0));
		for (int ch = 0; ch < nch; ch++)
		{
			const int z0 = (int)((long long)Zmax * ch / nch), z1 = (int)((long long)Zmax * (ch + 1) / nch);
			CU(cudaMemcpyAsync(c->stage + plane_aos * z0, h_U + plane_aos * z0, plane_aos * (z1 - z0) * sizeof(double), cudaMemcpyHostToDevice, c->copy_stream));
			CU(cudaEventRecord(c->chunk_ev[ch], c->copy_stream));
		}
		if ((rc = xf_dt_device(c, t_end))) // dt from the maxima the previous step left (the device state is the host buffer's)
			return rc;
		const int lo = 2 * Bz, hi = Zmax - 2 * Bz; // planes [lo, hi) do not feed the z ghost fill
		for (int ch = 0; ch < nch; ch++)
		{
			const int z0 = (int)((long long)Zmax * ch / nch), z1 = (int)((long long)Zmax * (ch + 1) / nch);
			CU(cudaStreamWaitEvent(c->stream, c->chunk_ev[ch], 0));
			KL(c->t->layout(d, c->E, U, c->stage, 1, c->stream, (long long)z0 * d.Ymax, (long long)(z1 - z0) * d.Ymax));
			c->launches++;
			CU(cudaMemcpy2DAsync(U1 + (size_t)z0 * d.sZ, (size_t)d.N * sizeof(double), U + (size_t)z0 * d.sZ, (size_t)d.N * sizeof(double),
								 (size_t)(z1 - z0) * d.sZ * sizeof(double), c->E, cudaMemcpyDeviceToDevice, c->stream));
			KL(c->t->bc(d, c->E, c->cop, U, bc, c->stream, &c->launches, 3, z0, z1));
			const int p0 = z0 > lo ? z0 : lo, p1 = z1 < hi ? z1 : hi;
			if (p1 > p0)
			{
				if ((rc = update_states_range(c, U, false, false, p0, p1)))
					return rc;
				if ((rc = stage_sweeps(c, U, U1, LU, 1, 3, p0, p1, -1, -1, false)))
					return rc;
			}
		}
		return XF_OK;
	}
	int xf_host_stage1_finish(xf_ctx *c, const int bc[6], double *U, double *U1, double *LU)
	{
		const XfDev &d = c->d;
		const int Bz = d.Bz, Zmax = d.Zmax, lo = 2 * Bz, hi = Zmax - 2 * Bz;
		int rc;
		// the z ghost fill, then the planes that waited for it
		KL(c->t->bc(d, c->E, c->cop, U, bc, c->stream, &c->launches, 4, -1, -1));
		if ((rc = update_states_range(c, U, false, false, 0, lo)) || (rc = update_states_range(c, U, false, false, hi, Zmax)))
			return rc;
		if ((rc = stage_sweeps(c, U, U1, LU, 1, 3, Bz, lo, -1, -1, false)) || (rc = stage_sweeps(c, U, U1, LU, 1, 3, hi, Zmax - Bz, -1, -1, false)))
			return rc;
		return stage_sweeps(c, U, U1, LU, 1, 4, -1, -1, -1, -1, true);
	}
	int xf_host_stage3(xf_ctx *c, double *h_U, double *U, double *U1, double *LU, int *error)
	{
		const XfDev &d = c->d;
		const int Bz = d.Bz, Zmax = d.Zmax, nch = c->host_chunks;
		const size_t plane_aos = (size_t)d.Xmax * d.Ymax * c->E;
		int rc;
		if (!host_overlap_ok(c) || !c->copy_stream)
			return fail(XF_ERR_ARG, "xf_host_stage3 without xf_host_begin");
		// stage 3 in z-chunks with the download behind it: primitive recovery of U1 on the whole block (its ghost fill / halo was the caller's; gathers the
		// dt maxima of the next step), then per chunk of z tiles the three sweeps, the update of the planes whose two z faces are
		// now known, the SoA->AoS conversion of those planes and their copy to the host on the copy stream.
		if ((rc = update_states(c, U1, true)))
			return rc;
		// chunks of inner planes.  Tiled sweeps: whole z tiles (TF faces each); the planes [ka, kb) whose two z faces are known after tiles
		// < t1 are updated.  Marching sweeps: the cells one segment of the march updates, m TF - 1 planes (m full iterations of faces,
		// the first face only seeds the divergence).
		const int TF = c->tiled ? xf_strict::z_tile_faces() : XF_MTF(d.weno), Zi = d.Zi, ntz = xf_strict::xf_z_tiles(d);
		int clen = ((Zi + nch - 1) / nch + 1 + TF - 1) / TF * TF - 1;
		clen = clen < TF - 1 ? TF - 1 : clen;
		const int nd = c->tiled ? (nch < ntz ? nch : ntz) : (Zi + clen - 1) / clen;
		if ((int)c->chunk_ev.size() < nch + 1 + nd)
		{
			const size_t old = c->chunk_ev.size();
			c->chunk_ev.resize(nch + 1 + nd);
			for (size_t i = old; i < c->chunk_ev.size(); i++)
				CU(cudaEventCreateWithFlags(&c->chunk_ev[i], cudaEventDisableTiming));
		}
		for (int ch = 0; ch < nd; ch++)
		{
			int ka, kb; // inner planes (0-based) [ka, kb) this chunk updates
			if (c->tiled)
			{
				const int t0 = (int)((long long)ntz * ch / nd), t1 = (int)((long long)ntz * (ch + 1) / nd);
				// tiles [t0, t1) = z faces Bz - 1 + TF t0 .. Bz - 2 + TF t1
				ka = TF * t0 - 1, kb = TF * t1 - 1;
				ka = ka < 0 ? 0 : ka, kb = kb > Zi ? Zi : kb;
				if (ch == nd - 1)
					kb = Zi;
				if ((rc = tiled_sweeps(c, U1, 3, Bz + ka, Bz + kb, -1, -1)))
					return rc;
				if ((rc = tiled_sweeps(c, U1, 4, -1, -1, t0, t1)))
					return rc;
				KL(c->t->rk(d, c->E, U, U1, LU, 0.0, d.red + XF_RED_DT, 3, 1, 1, c->stream, ka, kb));
				c->launches++;
			}
			else
			{
				ka = ch * clen, kb = (ch + 1) * clen < Zi ? (ch + 1) * clen : Zi;
				if ((rc = stage_sweeps(c, U, U1, LU, 3, 3, Bz + ka, Bz + kb, -1, -1, false)) || (rc = stage_sweeps(c, U, U1, LU, 3, 4, -1, -1, Bz + ka, Bz + kb, true)))
					return rc;
			}
			// planes to ship: the updated inner planes, plus the z ghost planes with the first / last chunk
			const int z0 = ch == 0 ? 0 : Bz + ka, z1 = ch == nd - 1 ? Zmax : Bz + kb;
			KL(c->t->layout(d, c->E, U, c->stage, 0, c->stream, (long long)z0 * d.Ymax, (long long)(z1 - z0) * d.Ymax));
			c->launches++;
			CU(cudaEventRecord(c->chunk_ev[nch + 1 + ch], c->stream));
			CU(cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[nch + 1 + ch], 0));
			CU(cudaMemcpyAsync(h_U + plane_aos * z0, c->stage + plane_aos * z0, plane_aos * (z1 - z0) * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
		}
		int f[4];
		if ((rc = xf_error_flags(c, f))) // synchronises the compute stream
			return rc;
		CU(cudaStreamSynchronize(c->copy_stream));
		const int err = (f[0] || f[1] || f[2]) ? 1 : 0;
		if (error)
			*error = err;
		return err ? XF_ERR_NUMERIC : XF_OK;
	}
	static int step_host_overlapped(xf_ctx *c, double *h_U, const int bc[6], double t_end, double *U, double *U1, double *LU, int *steps_done, int *error)
	{
		int rc;
		if ((rc = xf_host_begin(c, h_U, bc, t_end, U, U1, LU)) || (rc = xf_host_stage1_finish(c, bc, U, U1, LU)) || (rc = xf_rk_stage(c, U, U1, LU, bc, 2)) ||
			(rc = xf_boundary(c, U1, bc)))
			return rc;
		if (steps_done)
			*steps_done = 1;
		return xf_host_stage3(c, h_U, U, U1, LU, error);
	}
	int xf_step_host(xf_ctx *c, double *h_U, const int bc[6], int nsteps, double t_end, double *U, double *U1, double *LU, int *steps_done, int *error)
	{
		int rc;
		// (host_overlap_ok: not for GLF, which needs the block-wide maxima of |lambda| of THIS stage's primitives before any sweep starts,
		// ConVenction_block.hpp:115-215 -- no chunk-wise overlap of primitive recovery and sweeps there)
		const bool overlap = nsteps == 1 && host_overlap_ok(c);
		if (overlap)
			return step_host_overlapped(c, h_U, bc, t_end, U, U1, LU, steps_done, error); // downloads as it goes
		else
		{
			if ((rc = xf_upload_aos(c, U, h_U)))
				return rc;
			CU(cudaMemcpyAsync(U1, U, xf_field_doubles(c) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
			rc = xf_run(c, U, U1, LU, bc, nsteps, t_end, steps_done, nullptr, error);
		}
		if (rc && rc != XF_ERR_NUMERIC)
			return rc;
		int rc2 = xf_download_aos(c, U, h_U);
		return rc2 ? rc2 : rc;
	}
}
