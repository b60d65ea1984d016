// xf_kernels.cu -- hand-written sm_100a kernels of the per-stage right-hand side.  Compiled twice:
//   -DXF_NS=xf_strict -fmad=false   (parity mode: no FMA contraction)
//   -DXF_NS=xf_fast   -fmad=true    (same formulas, contraction allowed)
// Kernels (one launch each unless noted; a<n> = the SURVEY 8(a) function rows they cover):
//   k_prim, k_prim_hard   cons->prim + Newton T + per-cell pressure derivatives (+ dt / GLF maxima, guards)            a3,a4,a5,a1,a6',a10
//   k_prim_shell          the same for the cells around the deep block (opt-in update -> recovery fusion)
//   k_sweep<.., PP, ACC, VISC>  characteristic WENO flux at the faces of one direction, stencil staged in shared memory (the conserved
//                         pencil of a y / z tile by one TMA bulk-tensor copy), limiter and viscous wall flux in the tail        a6,a7 (+ f1, f3)
//   k_lu                  flux divergence (block-level API)                                                                    a8
//   k_rk<E, fused>        flux divergence + NaN guard + SSP-RK3 stage update: LU never touches HBM on the fused path           a8,a9,a10
//   k_rk_prim             k_rk going straight on to the next stage's primitive recovery (opt-in, XF_FUSE_PRIM=1)
//   k_march               TMA-fed marching y / z sweep with divergence / update fused in (xf_march.cuh; opt-in, XF_MARCH=1)
//   k_vde, k_vde_bc, k_transport, k_yi_minmax, k_visc_limits, k_visc_flux3   viscous block (xf_visc.cuh)                        f3
//   k_bc                  ghost-cell fill, one launch per direction                                                            a2
//   k_dt                  stand-alone CFL maxima (API parity with GetDt)                                                       a1
//   k_dt_final            dt = CFL/sum, clip to t_end, advance device time
//   k_layout / k_scalar_pad / k_halo   AoS <-> SoA at the boundary, z-slab halo pack / unpack
#include <cuda_runtime.h>
#include <cstdio>
#include "xf_math.cuh"
#include "xf_visc_face.cuh"
#include "xf_launch.h"
#include "xf_tma.cuh"

#ifndef XF_NS
#error "XF_NS must be defined"
#endif

namespace XF_NS
{

// ---------------------------------------------------------------------------------------------
// reductions: warp shuffle -> one atomic per warp on the bit pattern (values are >= 0, so the
// IEEE ordering equals the unsigned-integer ordering)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_atomic_max_pos(double *addr, double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
	if ((threadIdx.x & 31) == 0)
		atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ bool xf_bad(double x) { return (x < 0) || isnan(x) || isinf(x); }

// ---------------------------------------------------------------------------------------------
// k_prim: Updaterhoyi + UpdateFuidStatesKernel (Update_kernels.hpp:5-48, Update_device.hpp:7-54) over ALL cells
// incl. ghosts, guards (Estimate_kernels.hpp) on inner cells, and the per-cell halves of ReconstructSoundSpeed.
// flags: bit0 gather dt maxima, bit1 gather GLF maxima
//
// The Newton iteration for T has a data-dependent trip count (1-3 in smooth flow, 20-25 at shock fronts, the
// 100-iteration cap at fresh discontinuities), so a single pass leaves most lanes of every warp that
// touches a front idle (ncu: 8.5 of 32 lanes active on average).  Two passes instead, same iteration sequence per cell:
//   k_prim       every cell, at most XF_NEWTON_FAST iterations; cells that have not met the stop criterion park their
//                current T / iteration count and append their index to a device list
//   k_prim_hard  one thread per list entry continues the same iteration to the stop criterion or the cap
// Both end in the same epilogue.
// ---------------------------------------------------------------------------------------------
#ifndef XF_NEWTON_FAST_
#define XF_NEWTON_FAST_ 3
#endif
constexpr int XF_NEWTON_FAST = XF_NEWTON_FAST_;
#ifndef XF_PRIM_MINB
#define XF_PRIM_MINB 5   // (measured: 5 -> 22.5 ms, 6 -> 25.2, 8 -> 23.3, 10 -> 24.3 at 512x512x256) resident 128-thread blocks per SM k_prim is compiled for (register cap 65536 / (128 * XF_PRIM_MINB))
#endif

template <class C>
struct PrimCell
{
	double rho, rho1, U1, U2, U3, U4, u, v, w, q2, tme, R, Wm;
	double yi[C::NS];
};

// The limited Newton iteration of get_T (Mixing_device.h:171-193) from iteration `it` on, written around ONE evaluation
// site of the species enthalpies / mixture Cp (the loop is not unrolled; the evaluation that follows the last update is the
// one the epilogue needs, at the final T).  Returns false when `limit` iterations have been done without meeting the stop
// criterion and `park` is set (T then holds the iterate to resume from); with park == false the limit is the reference's
// 100-iteration cap, after which T is used as it is.
template <class C>
XF_DEV bool xf_newton(const XfThermo &th, const double *yi, double e, double R, double &T, int it, int limit, bool park, double *hi, double &Cp)
{
	bool conv = false;
#pragma unroll 1
	for (;;)
	{
		if (!conv && it == limit)
		{
			if (park)
				return false;
			conv = true;
		}
		xf_species_h<C>(th, T, hi);
		Cp = xf_mix_cp<C>(th, yi, T);
		if (conv)
			return true;
		double h = 0.0;
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			h += hi[n] * yi[n];
		const double func_T = h - R * T - e;
		const double dfunc_T = Cp - R;
		double df = xf_min(func_T / (dfunc_T + 1.0e-30), 1e-3 * T);
		df = xf_max(df, -1e-2 * T);
		T = T - df;
		it++;
		conv = fabs(df) <= 1.0e-6;
	}
}

template <class C>
XF_DEV void prim_epilogue(const XfDev &d, const XfThermo &th, long long id, bool inner, const PrimCell<C> &pc, double T, const double *hi, double Cp,
						  int flags, double *dtm, double *glf)
{
	constexpr int NS = C::NS, NC = C::NC;
	const double rho = pc.rho, rho1 = pc.rho1, u = pc.u, v = pc.v, w = pc.w, q2 = pc.q2;
	double p, gamma;
	if constexpr (C::COP)
	{
		const double R = pc.R, Wm = pc.Wm;
		const double *yi = pc.yi;
		p = rho * R * T;
		// Cp = get_CopCp(T), hi = species enthalpies at T: evaluated by the caller (xf_newton's last evaluation)
		// 4-argument get_CopGamma (Mixing_device.h:114-126)
		const double CopW = 1.0 / Wm;
		const double g4 = Cp / (Cp - th.Ru / CopW);
		gamma = (g4 > 1.0) ? g4 : -1.0;
		const double H = (pc.U4 + p) * rho1;
		// per-cell pieces of ReconstructSoundSpeed (Utils_device.hpp:102-131)
		const double Cv = Cp - th.Ru * Wm;
		const double g3 = Cp / Cv; // 3-argument get_CopGamma
		const double prho = p / rho;
		const double e_l = H - 0.5 * q2 - prho;
		const double RN = th.Ri[NC];
		const double RNT = RN * T;
		d.dpdrho[id] = (g3 - 1.0) * (0.5 * q2 - hi[NC] + Cp * RNT / R);
#pragma unroll
		for (int n = 0; n < NC; n++)
		{
			const double hN_minus_hi = -hi[n] + hi[NC];
			const double Ri_minus_RN = (th.Ri[n] - RN);
			d.dpdrhoi[n * d.N + id] = (g3 - 1.0) * (hN_minus_hi + Cp * Ri_minus_RN * T / R);
		}
		d.g3[id] = g3, d.e[id] = e_l, d.prho[id] = prho;
		d.T[id] = T, d.H[id] = H;
	}
	else
	{
		gamma = d.gamma0;
		p = (d.gamma0 - 1.0) * rho * pc.tme;
		d.H[id] = (pc.U4 + p) * rho1;
	}
	const double cc = sqrt(gamma * p * rho1);
	d.u[id] = u, d.v[id] = v, d.w[id] = w, d.p[id] = p, d.c[id] = cc;

	// guards (flag only): EstimateYiKernel / EstimatePrimitiveVarKernel, inner cells
	if (inner)
	{
		bool e0 = xf_bad(rho);
		if constexpr (C::COP)
		{
#pragma unroll
			for (int n = 0; n < NS; n++)
				e0 = e0 || isnan(pc.yi[n]) || isinf(pc.yi[n]);
		}
		if (e0)
			d.err[0] = 1;
		if (xf_bad(rho) || xf_bad(p) || (C::COP && xf_bad(T)))
			d.err[1] = 1;
	}
	if (flags & 1)
	{ // GetDt: hard-coded 1.4, all cells incl. ghosts (GlobalDt_block.hpp:34-66)
		const double c_local = sqrt(1.4 * p / rho);
		dtm[0] = fabs(u) + c_local, dtm[1] = fabs(v) + c_local, dtm[2] = fabs(w) + c_local;
	}
	if (flags & 2)
	{ // GetLocalEigen maxima (Eigen_value.hpp:19-28): u_d - c, u_d, u_d + c
		glf[0] = fabs(u - cc), glf[1] = fabs(u), glf[2] = fabs(u + cc);
		glf[3] = fabs(v - cc), glf[4] = fabs(v), glf[5] = fabs(v + cc);
		glf[6] = fabs(w - cc), glf[7] = fabs(w), glf[8] = fabs(w + cc);
	}
}

__device__ __forceinline__ void prim_reduce(const XfDev &d, int flags, const double *dtm, const double *glf)
{
	if (flags & 1)
	{
		if (d.DimX) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 0, dtm[0]);
		if (d.DimY) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 1, dtm[1]);
		if (d.DimZ) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 2, dtm[2]);
	}
	if (flags & 2)
	{
#pragma unroll
		for (int q = 0; q < 9; q++)
			if ((q < 3 && d.DimX) || (q >= 3 && q < 6 && d.DimY) || (q >= 6 && d.DimZ))
				warp_atomic_max_pos(d.red + XF_RED_GLF + q, glf[q]);
	}
}

__device__ __forceinline__ bool cell_is_inner(const XfDev &d, long long id)
{
	const int i = int(id % d.Xp);
	const long long row = id / d.Xp;
	const int j = int(row % d.Ymax), k = int(row / d.Ymax);
	return i >= d.Bx && i < d.Xmax - d.Bx && j >= d.By && j < d.Ymax - d.By && k >= d.Bz && k < d.Zmax - d.Bz;
}

#ifndef XF_PRIM_PIPE
#define XF_PRIM_PIPE 8 // 128-cell chunks per block, the next chunk's U and T prefetched with cp.async while the current one is computed (0: off).  Measured 512x256x256: off 10.1, 2: 8.79, 4: 8.43, 8: 8.26 ms per step
#endif
__device__ __forceinline__ void xf_cp_async8(double *smem_dst, const double *gmem_src)
{
	const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem_src));
}
// one cell of the primitive recovery; Uc[0..E-1] = the cell's conserved variables, Tc = its Newton warm start
template <class C>
__device__ __forceinline__ void prim_cell(const XfDev &d, const XfThermo &th, double *__restrict__ U, long long id, int i, int j, int k,
										  const double *Uc, double Tc, int flags, double *dtm, double *glf)
{
	constexpr int NS = C::NS, NC = C::NC;
	PrimCell<C> pc;
	pc.rho = Uc[0];
	pc.rho1 = 1.0 / pc.rho;
	if constexpr (C::COP)
	{
		double *yi = pc.yi;
		if (d.ghost)
		{ // GhostSpecies: renormalise and write back into U (Update_device.hpp:15-22)
			yi[NC] = 0.0;
			double sum_yi = 0.0;
#pragma unroll
			for (int ii = 0; ii < NC; ii++)
				yi[ii] = Uc[5 + ii] * pc.rho1, sum_yi += yi[ii];
			sum_yi = 1.0 / sum_yi;
#pragma unroll
			for (int ii = 0; ii < NC; ii++)
				yi[ii] *= sum_yi, U[(5 + ii) * d.N + id] = pc.rho * yi[ii];
		}
		else
		{
			yi[NC] = 1.0;
#pragma unroll
			for (int ii = 0; ii < NC; ii++)
				yi[ii] = Uc[5 + ii] * pc.rho1, yi[NC] += -yi[ii];
		}
#pragma unroll
		for (int n = 0; n < NS; n++)
			d.y[n * d.N + id] = yi[n];
	}
	pc.U1 = Uc[1], pc.U2 = Uc[2], pc.U3 = Uc[3], pc.U4 = Uc[4];
	pc.u = pc.U1 * pc.rho1, pc.v = pc.U2 * pc.rho1, pc.w = pc.U3 * pc.rho1;
	pc.q2 = pc.u * pc.u + pc.v * pc.v + pc.w * pc.w;
	pc.tme = pc.U4 * pc.rho1 - 0.5 * pc.q2;
	double T = 0.0, Cp = 0.0, hi[NS];
	bool done = true;
	if constexpr (C::COP)
	{
		double Wm = 0.0; // sum yi/Wi
#pragma unroll
		for (int n = 0; n < NS; n++)
			Wm += pc.yi[n] * th._Wi[n];
		pc.Wm = Wm, pc.R = Wm * th.Ru;
		T = Tc;
		done = xf_newton<C>(th, pc.yi, pc.tme, pc.R, T, 0, XF_NEWTON_FAST, true, hi, Cp);
		if (!done)
		{ // park: T after XF_NEWTON_FAST steps; k_prim_hard resumes at iteration XF_NEWTON_FAST + 1
			d.T[id] = T;
			const unsigned slot = atomicAdd(d.hard_count, 1u);
			d.hard_ids[slot] = (unsigned)id;
		}
	}
	if (done)
	{
		const bool inner = i >= d.Bx && i < d.Xmax - d.Bx && j >= d.By && j < d.Ymax - d.By && k >= d.Bz && k < d.Zmax - d.Bz;
		double dt1[3] = {0.0, 0.0, 0.0}, gl1[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
		prim_epilogue<C>(d, th, id, inner, pc, T, hi, Cp, flags, dt1, gl1);
		if (flags & 1)
		{
#pragma unroll
			for (int q = 0; q < 3; q++)
				dtm[q] = fmax(dtm[q], dt1[q]);
		}
		if (flags & 2)
		{
#pragma unroll
			for (int q = 0; q < 9; q++)
				glf[q] = fmax(glf[q], gl1[q]);
		}
	}
}

// grid: x = chunks of 128 * max(XF_PRIM_PIPE, 1) cells of one z-plane's linear index space, y = plane (lin0 = the first plane);
// 32-bit index arithmetic inside the plane (a 64-bit division per cell costs more instructions than the whole epilogue).
// ncu (source-level sampling) had 30 % of this kernel's samples on the first use of the cell's freshly loaded U -- DRAM latency that
// 20 resident warps per SM cannot hide -- hence the cp.async double buffer: each thread prefetches ITS OWN next cell (no barrier).
template <class C>
__global__ void __launch_bounds__(128, XF_PRIM_MINB) k_prim(XfDev d, XfThermo th, double *__restrict__ U, int flags, long long lin0 /* first z-plane */, long long lin1,
																	 int split = 0x7fffffff, int gap = 0 /* planes lin0 + y, skipping `gap` planes from the split-th on */)
{
	constexpr int E = C::E, NV = E + (C::COP ? 1 : 0), CH = XF_PRIM_PIPE > 0 ? XF_PRIM_PIPE : 1;
	const int kq = int(lin0) + int(blockIdx.y) + (int(blockIdx.y) >= split ? gap : 0);
	const long long pbase = (long long)kq * d.sZ;
	double dtm[3] = {0.0, 0.0, 0.0}, glf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	(void)lin1;
#if XF_PRIM_PIPE > 0
	__shared__ double buf[2][NV][128];
	auto issue = [&](int c, int b)
	{
		const unsigned q = (blockIdx.x * CH + c) * 128u + threadIdx.x;
		if (q < (unsigned)d.sZ)
		{
#pragma unroll
			for (int n = 0; n < E; n++)
				xf_cp_async8(&buf[b][n][threadIdx.x], U + n * d.N + pbase + q);
			if constexpr (C::COP)
				xf_cp_async8(&buf[b][E][threadIdx.x], d.T + pbase + q);
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
	};
	issue(0, 0);
#endif
#pragma unroll 1
	for (int c = 0; c < CH; c++)
	{
		const unsigned q = (blockIdx.x * CH + c) * 128u + threadIdx.x;
		const unsigned jq = q / (unsigned)d.Xp, iq = q - jq * (unsigned)d.Xp;
		const bool active = q < (unsigned)d.sZ && int(iq) < d.Xmax;
		double Uc[E], Tc = 0.0;
#if XF_PRIM_PIPE > 0
		if (c + 1 < CH)
		{
			issue(c + 1, (c + 1) & 1);
			asm volatile("cp.async.wait_group 1;" ::: "memory");
		}
		else
			asm volatile("cp.async.wait_group 0;" ::: "memory");
		if (active)
		{
#pragma unroll
			for (int n = 0; n < E; n++)
				Uc[n] = buf[c & 1][n][threadIdx.x];
			if constexpr (C::COP)
				Tc = buf[c & 1][E][threadIdx.x];
		}
#else
		if (active)
		{
#pragma unroll
			for (int n = 0; n < E; n++)
				Uc[n] = U[n * d.N + pbase + q];
			if constexpr (C::COP)
				Tc = d.T[pbase + q];
		}
#endif
		if (active)
			prim_cell<C>(d, th, U, pbase + q, int(iq), int(jq), kq, Uc, Tc, flags, dtm, glf);
	}
	prim_reduce(d, flags, dtm, glf);
}

// Primitive recovery of the shell of the middle z-planes [dp.zlo, dp.zhi): the rows outside [ylo, yhi) whole, of the others the cells
// outside [xlo, xhi).  (The planes outside [zlo, zhi) are whole planes: k_prim.)
template <class C>
__global__ void __launch_bounds__(128, XF_PRIM_MINB) k_prim_shell(XfDev d, XfThermo th, XfDeep dp, double *__restrict__ U, int flags)
{
	constexpr int E = C::E;
	const int kq = dp.zlo + int(blockIdx.y);
	const unsigned t = blockIdx.x * 128u + threadIdx.x;
	const unsigned nfull = unsigned(dp.ylo + (d.Ymax - dp.yhi)) * unsigned(d.Xp), nside = unsigned(dp.xlo + (d.Xmax - dp.xhi));
	int i, j;
	bool active;
	if (t < nfull)
	{
		const unsigned jr = t / unsigned(d.Xp);
		i = int(t - jr * unsigned(d.Xp));
		j = int(jr) < dp.ylo ? int(jr) : dp.yhi + (int(jr) - dp.ylo);
		active = i < d.Xmax;
	}
	else
	{
		const unsigned q = t - nfull, jr = nside ? q / nside : 0u, ii = nside ? q - jr * nside : 0u;
		j = dp.ylo + int(jr);
		i = int(ii) < dp.xlo ? int(ii) : dp.xhi + (int(ii) - dp.xlo);
		active = nside > 0 && j < dp.yhi;
	}
	double dtm[3] = {0.0, 0.0, 0.0}, glf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	if (active)
	{
		const long long id = ((long long)kq * d.Ymax + j) * d.Xp + i;
		double Uc[E], Tc = 0.0;
#pragma unroll
		for (int n = 0; n < E; n++)
			Uc[n] = U[n * d.N + id];
		if constexpr (C::COP)
			Tc = d.T[id];
		prim_cell<C>(d, th, U, id, i, j, kq, Uc, Tc, flags, dtm, glf);
	}
	prim_reduce(d, flags, dtm, glf);
}

template <class C>
__global__ void __launch_bounds__(128) k_prim_hard(XfDev d, XfThermo th, const double *__restrict__ U, int flags)
{
	constexpr int NS = C::NS;
	const unsigned count = *d.hard_count;
	const unsigned stride = gridDim.x * blockDim.x;
	// warp-uniform trip count: every lane of a warp takes part in the shuffles of prim_reduce
	for (unsigned base = blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < count; base += stride)
	{
		const unsigned q = base + (threadIdx.x & 31u);
		double dtm[3] = {0.0, 0.0, 0.0}, glf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
		if (q < count)
		{
			const long long id = d.hard_ids[q];
			PrimCell<C> pc;
			pc.rho = U[id];
			pc.rho1 = 1.0 / pc.rho;
#pragma unroll
			for (int n = 0; n < NS; n++)
				pc.yi[n] = d.y[n * d.N + id];
			pc.U1 = U[1 * d.N + id], pc.U2 = U[2 * d.N + id], pc.U3 = U[3 * d.N + id], pc.U4 = U[4 * d.N + id];
			pc.u = pc.U1 * pc.rho1, pc.v = pc.U2 * pc.rho1, pc.w = pc.U3 * pc.rho1;
			pc.q2 = pc.u * pc.u + pc.v * pc.v + pc.w * pc.w;
			pc.tme = pc.U4 * pc.rho1 - 0.5 * pc.q2;
			double Wm = 0.0;
#pragma unroll
			for (int n = 0; n < NS; n++)
				Wm += pc.yi[n] * th._Wi[n];
			pc.Wm = Wm, pc.R = Wm * th.Ru;
			double T = d.T[id], Cp, hi[NS];
			xf_newton<C>(th, pc.yi, pc.tme, pc.R, T, XF_NEWTON_FAST, 100, false, hi, Cp);
			prim_epilogue<C>(d, th, id, cell_is_inner(d, id), pc, T, hi, Cp, flags, dtm, glf);
		}
		prim_reduce(d, flags, dtm, glf);
	}
}

// stand-alone GetDt maxima from the stored primitives (rho is U[0])
__global__ void __launch_bounds__(256) k_dt(XfDev d, const double *__restrict__ rho)
{
	const long long lin = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	double m0 = 0.0, m1 = 0.0, m2 = 0.0;
	if (lin < d.N && int(lin % d.Xp) < d.Xmax)
	{
		const double c_local = sqrt(1.4 * d.p[lin] / rho[lin]);
		m0 = fabs(d.u[lin]) + c_local, m1 = fabs(d.v[lin]) + c_local, m2 = fabs(d.w[lin]) + c_local;
	}
	if (d.DimX) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 0, m0);
	if (d.DimY) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 1, m1);
	if (d.DimZ) warp_atomic_max_pos(d.red + XF_RED_DTMAX + 2, m2);
}

// dt = CFL / (max_x*_dx + max_y*_dy + max_z*_dz), clipped to t_end; time += dt; maxima reset for the next gather
__global__ void k_dt_final(XfDev d, double t_end)
{
	double *r = d.red;
	const double dtref = r[XF_RED_DTMAX + 0] * d._dx + r[XF_RED_DTMAX + 1] * d._dy + r[XF_RED_DTMAX + 2] * d._dz;
	double dt = d.CFL / dtref;
	const double t = r[XF_RED_TIME];
	if (t + dt > t_end)
		dt = t_end - t;
	r[XF_RED_DT] = dt;
	r[XF_RED_TIME] = t + dt;
	if (dt > 0.0)
		r[XF_RED_STEPS] += 1.0;
	// the positivity-preserving limiter of the coming stages reads uvw_c_max as this GetDt left it (ConVenction_block.hpp:332)
	r[XF_RED_PPL + 0] = r[XF_RED_DTMAX + 0], r[XF_RED_PPL + 1] = r[XF_RED_DTMAX + 1], r[XF_RED_PPL + 2] = r[XF_RED_DTMAX + 2];
	r[XF_RED_DTMAX + 0] = 0.0, r[XF_RED_DTMAX + 1] = 0.0, r[XF_RED_DTMAX + 2] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// k_sweep: ReconstructFlux{X,Y,Z} (Reconstruction_kernels.hpp:8-199).  One thread per face; the stencil's conserved
// variables, (halved) physical fluxes (GetPhysFlux, Update_device.hpp:81-110) and local wave speeds are staged once per
// tile in shared memory as SoA pencils [component][cell], so a face reads its NST cells conflict-free.
//   DIR 0: tile = 128 consecutive cells of the linear index space (rows are contiguous; non-face cells idle)
//   DIR 1/2: tile = 32 (x) x TF faces along the sweep; staged rows = TF + NST - 1
// ---------------------------------------------------------------------------------------------
template <class C, int WENO>
struct SmemStencil
{
	const double *sU, *sF, *sL; // [E][ncell], [E][ncell], [3][ncell]
	int ncell, base, stride;    // cell index of stencil slot s: base + s*stride
	__device__ __forceinline__ double U(int s, int n) const { return sU[n * ncell + base + s * stride]; }
	__device__ __forceinline__ double F(int s, int n) const { return sF[n * ncell + base + s * stride]; }
	__device__ __forceinline__ double lam(int s, int t) const { return sL[t * ncell + base + s * stride]; }
};

template <class C, int DIR>
__device__ __forceinline__ void stage_cell(const XfDev &d, const double *__restrict__ U, long long id, bool valid,
										   double *sU, double *sF, double *sL, int ncell, int c)
{
	constexpr int E = C::E, NC = C::NC;
	if (!valid)
	{
#pragma unroll
		for (int n = 0; n < E; n++)
			sU[n * ncell + c] = 0.0, sF[n * ncell + c] = 0.0;
		sL[c] = 0.0, sL[ncell + c] = 0.0, sL[2 * ncell + c] = 0.0;
		return;
	}
	double Uc[E];
#pragma unroll
	for (int n = 0; n < E; n++)
		Uc[n] = U[n * d.N + id];
	const double u = d.u[id], v = d.v[id], w = d.w[id], p = d.p[id], cc = d.c[id];
	const double un = DIR == 0 ? u : (DIR == 1 ? v : w);
	const double m = Uc[1 + DIR]; // rho*u_d
	double Fc[E];
	Fc[0] = m;
	Fc[1] = DIR == 0 ? m * u + p : m * u;
	Fc[2] = DIR == 1 ? m * v + p : m * v;
	Fc[3] = DIR == 2 ? m * w + p : m * w;
	Fc[4] = (Uc[4] + p) * un;
#pragma unroll
	for (int s = 0; s < NC; s++)
		Fc[5 + s] = m * d.y[s * d.N + id];
	// the physical flux is staged HALVED: every use below is linear in F and 0.5 * F commutes with every rounding of the projection
	// sums, so 0.5 * (sum_k F_k l_k) == sum_k (0.5 F_k) l_k bit for bit -- one multiplication per staged cell instead of one per
	// stencil point, field and face in the Lax-Friedrichs split
#pragma unroll
	for (int n = 0; n < E; n++)
		sU[n * ncell + c] = Uc[n], sF[n * ncell + c] = 0.5 * Fc[n];
	sL[c] = fabs(un - cc), sL[ncell + c] = fabs(un), sL[2 * ncell + c] = fabs(un + cc);
}

template <class C>
__device__ __forceinline__ void load_side(const XfDev &d, const double *__restrict__ U, long long id, XfSide<C> &s)
{
	s.rho = U[id], s.u = d.u[id], s.v = d.v[id], s.w = d.w[id], s.H = d.H[id], s.p = d.p[id];
	if constexpr (C::COP)
	{
		s.g3 = d.g3[id], s.dpdrho = d.dpdrho[id], s.e = d.e[id], s.prho = d.prho[id];
#pragma unroll
		for (int n = 0; n < C::NC; n++)
			s.y[n] = d.y[n * d.N + id], s.dpdrhoi[n] = d.dpdrhoi[n * d.N + id];
	}
}

// the same without rho (the marching sweeps take rho from the staged copy of U)
template <class C>
__device__ __forceinline__ void side_scalars(const XfDev &d, long long id, XfSide<C> &s)
{
	s.u = d.u[id], s.v = d.v[id], s.w = d.w[id], s.H = d.H[id], s.p = d.p[id];
	if constexpr (C::COP)
	{
		s.g3 = d.g3[id], s.dpdrho = d.dpdrho[id], s.e = d.e[id], s.prho = d.prho[id];
#pragma unroll
		for (int n = 0; n < C::NC; n++)
			s.y[n] = d.y[n * d.N + id], s.dpdrhoi[n] = d.dpdrhoi[n * d.N + id];
	}
}

#ifndef XF_MINB_X
#define XF_MINB_X 4   // resident blocks per SM the x sweep is compiled for (register cap 65536 / (XF_MINB_X * 128))
#endif
#ifndef XF_MINB_YZ
#define XF_MINB_YZ 2  // same for the y / z sweeps (256 threads per block)
#endif
#ifndef XF_TX_
#define XF_TX_ 128
#endif
#ifndef XF_SWEEP_TMA
#define XF_SWEEP_TMA 1 // y / z sweeps: the conserved-variable pencil of a tile is one TMA bulk-tensor copy straight into its shared-memory slot (0: per-thread loads)
#endif
#ifndef XF_SIDE_EARLY
#define XF_SIDE_EARLY 0 // 1: issue the face-side loads before the stencil staging (their latency overlaps the staging)
#endif
// x-sweep: faces (cells) per block: 128 (4 blocks/SM), except WENO-CU6, whose long instruction stream (with the limiter even more so) runs
// better in 256-thread blocks, 2 per SM, like the y / z sweeps: with four 128-thread blocks at four different places of the code
// "no instruction" was its top stall reason (ncu, CU6 + limiter: 2.24 stalled warps per issue in x, absent in y / z).  Measured
// (512x256x256, x sweep ms per step, 128 vs 256): WENO5+PP 20.5 / 21.0, CU6 33.9 / 33.3, CU6+PP 41.6 / 36.0, WENO7 36.3 / 37.2, WENO7+PP 38.8 / 39.4
#ifndef XF_TX_BIG
#define XF_TX_BIG 256
#endif
template <int WENO, bool PP>
struct XfTx
{
	static constexpr int V = (WENO != 6) ? XF_TX_ : XF_TX_BIG;
	static constexpr int MINB = (WENO != 6) ? XF_MINB_X : (512 / XF_TX_BIG);
};
constexpr int XF_TW = XF_TW_; // y/z sweeps: tile width in x
constexpr int XF_TF = XF_TF_; // y/z sweeps: faces per tile along the sweep

// ACC (x sweep of the marching path only): instead of storing the wall flux, write LU = 0.0 + (F_{i-1} - F_i) * _dx -- a compile-time
// switch: as a run-time branch the extra tail cost the plain x sweep 6.7 % (75.6 vs 70.8 ms per step at 512^3; four 128-thread blocks
// per SM at four places of a long instruction stream are sensitive to its length)
template <class C, int DIR, int WENO, bool PP, bool ACC = false, bool VISC = false>
__global__ void __launch_bounds__(DIR == 0 ? XfTx<WENO, PP>::V : XF_TW * XF_TF, DIR == 0 ? XfTx<WENO, PP>::MINB : XF_MINB_YZ) k_sweep(XfDev d, const __grid_constant__ CUtensorMap tmU /* y / z sweeps: the sweep input as a 4-D tensor, box = one tile's stencil rows */,
		const double *__restrict__ U, double *__restrict__ Fw, int kp0 /* x / y sweeps: first z-plane of the plane range; z sweep: first tile of the tile range */,
		double *__restrict__ LU /* ACC */, XfViscF vf = XfViscF() /* VISC: the viscous wall flux is subtracted before the store */)
{
	constexpr int mode = ACC ? XF_MODE_ACC : XF_MODE_FW;
	constexpr int E = C::E, NST = XfStencil<WENO>::NST, P = XfStencil<WENO>::P, XF_TX = XfTx<WENO, PP>::V;
	extern __shared__ __align__(128) double smem[];
	XfSide<C> sl, sr;
	XfRoe<C> R;
	SmemStencil<C, WENO> st;
	long long id_l;
	bool valid;
	int sweep_tile = 0, ipos = 0;

	if constexpr (DIR == 0)
	{
		constexpr int ncell = XF_TX + NST - 1;
		double *sU = smem, *sF = smem + E * ncell, *sL = smem + 2 * E * ncell;
		// grid: x = XF_TX-cell chunks of the linear index space of one z-plane, y = inner plane (32-bit index arithmetic;
		// stencils of valid faces never leave their row, so cells outside the plane only feed idle threads).  ACC mode: a cell's
		// divergence needs the face on its left too, so consecutive chunks overlap by one face (thread 0 only seeds)
		const int kpl = kp0 + blockIdx.y;
		const int q0 = (mode == XF_MODE_FW) ? blockIdx.x * XF_TX : blockIdx.x * (XF_TX - 1) - 1;
		const long long id0 = (long long)kpl * d.sZ + q0;
		id_l = id0 + threadIdx.x;
		const unsigned ql = unsigned(q0) + threadIdx.x;
		const int j = int(ql / unsigned(d.Xp)), i = int(ql - unsigned(j) * unsigned(d.Xp));
		ipos = i;
		valid = ql < unsigned(d.sZ) && i >= d.Bx - 1 && i < d.Bx + d.Xi && j >= d.By && j < d.By + d.Yi;
		if (XF_SIDE_EARLY && valid)
			load_side<C>(d, U, id_l, sl), load_side<C>(d, U, id_l + 1, sr);
		for (int c = threadIdx.x; c < ncell; c += XF_TX)
		{
			const int q = q0 - P + c;
			const bool ok = q >= 0 && q < int(d.sZ) && int(unsigned(q) % unsigned(d.Xp)) < d.Xmax;
			stage_cell<C, DIR>(d, U, ok ? id0 - P + c : 0, ok, sU, sF, sL, ncell, c);
		}
		if (XF_SIDE_EARLY == 2 && valid)
			xf_roe_state<C>(sl, sr, d.gamma0, R);
		__syncthreads();
		st.sU = sU, st.sF = sF, st.sL = sL, st.ncell = ncell, st.base = threadIdx.x, st.stride = 1;
		if (!XF_SIDE_EARLY && valid)
			load_side<C>(d, U, id_l, sl), load_side<C>(d, U, id_l + 1, sr);
	}
	else
	{
		constexpr int nrow = XF_TF + NST - 1, ncell = nrow * XF_TW;
		double *sU = smem, *sF = smem + E * ncell, *sL = smem + 2 * E * ncell;
		const int tx = threadIdx.x % XF_TW, ty = threadIdx.x / XF_TW;
		// block -> (x chunk, tile along the sweep, transverse index) = blockIdx.(x, y, z) for both sweeps: blocks that share halo rows /
		// planes run close together, so the 5 (7) halo planes of a z tile come from L2 instead of DRAM a second time (ncu: 43.9 GB
		// read per launch with the z tile as the slowest block index, vs 30.2 GB in x / y)
		const int bx = blockIdx.x;
		sweep_tile = blockIdx.y + (DIR == 2 ? kp0 : 0);
		const int i = d.Bx + bx * XF_TW + tx;
		// faces along the sweep start at B-1 ; the other transverse index is inner
		int j, k, f0;
		long long sS; // cell stride along the sweep
		if constexpr (DIR == 1)
			f0 = d.By - 1 + sweep_tile * XF_TF, k = kp0 + blockIdx.z, j = 0, sS = d.sY;
		else
			f0 = d.Bz - 1 + sweep_tile * XF_TF, j = d.By + blockIdx.z, k = 0, sS = d.sZ;
		const int nmax = DIR == 1 ? d.Ymax : d.Zmax;
		const bool iok = i < d.Bx + d.Xi;
		const int qf = f0 + ty; // left cell of this thread's face
		const int qend = DIR == 1 ? d.By + d.Yi : d.Bz + d.Zi;
		valid = iok && qf < qend;
		id_l = DIR == 1 ? ((long long)k * d.Ymax + qf) * d.Xp + i : ((long long)qf * d.Ymax + j) * d.Xp + i;
		if (XF_SIDE_EARLY && valid)
			load_side<C>(d, U, id_l, sl), load_side<C>(d, U, id_l + sS, sr);
		if constexpr (XF_SWEEP_TMA != 0)
		{ // the conserved variables of the whole tile (nrow rows x XF_TW cells x E components) arrive by ONE cp.async.bulk.tensor in
		  // their final place (the staged pencil of U is a plain copy; rows / columns outside the arrays are zero-filled by the
		  // hardware); meanwhile every thread loads the primitives of its (at most two) cells, then forms 0.5 F and the wave speeds
			unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + (2 * E + 3) * ncell);
			if (threadIdx.x == 0)
			{
				xf_mbar_init(bar, 1);
				xf_mbar_expect_tx(bar, unsigned(E * ncell * sizeof(double)));
				xf_tma_load_4d(sU, &tmU, d.Bx + bx * XF_TW, DIR == 1 ? f0 - P : j, DIR == 1 ? k : f0 - P, 0, bar);
			}
			constexpr int NR = (nrow + XF_TF - 1) / XF_TF, NC = C::NC;
			double pu[NR], pv[NR], pw[NR], pp_[NR], pc[NR], py[NR][NC > 0 ? NC : 1];
#pragma unroll
			for (int it = 0; it < NR; it++)
			{
				const int r = ty + it * XF_TF, q = f0 - P + r;
				const bool ok = r < nrow && iok && q >= 0 && q < nmax;
				const long long id = ok ? (DIR == 1 ? ((long long)k * d.Ymax + q) * d.Xp + i : ((long long)q * d.Ymax + j) * d.Xp + i) : 0;
				pu[it] = ok ? d.u[id] : 0.0, pv[it] = ok ? d.v[id] : 0.0, pw[it] = ok ? d.w[id] : 0.0, pp_[it] = ok ? d.p[id] : 0.0, pc[it] = ok ? d.c[id] : 0.0;
#pragma unroll
				for (int s = 0; s < NC; s++)
					py[it][s] = ok ? d.y[s * d.N + id] : 0.0;
			}
			__syncthreads(); // the initialised barrier is visible to every thread
			xf_mbar_wait(bar, 0);
#pragma unroll
			for (int it = 0; it < NR; it++)
			{
				const int r = ty + it * XF_TF, c = r * XF_TW + tx;
				if (r < nrow)
				{
					const double un = DIR == 1 ? pv[it] : pw[it], m = sU[(1 + DIR) * ncell + c];
					sF[0 * ncell + c] = 0.5 * m;
					sF[1 * ncell + c] = 0.5 * (m * pu[it]);
					sF[2 * ncell + c] = 0.5 * (DIR == 1 ? m * pv[it] + pp_[it] : m * pv[it]);
					sF[3 * ncell + c] = 0.5 * (DIR == 2 ? m * pw[it] + pp_[it] : m * pw[it]);
					sF[4 * ncell + c] = 0.5 * ((sU[4 * ncell + c] + pp_[it]) * un);
#pragma unroll
					for (int s = 0; s < NC; s++)
						sF[(5 + s) * ncell + c] = 0.5 * (m * py[it][s]);
					sL[c] = fabs(un - pc[it]), sL[ncell + c] = fabs(un), sL[2 * ncell + c] = fabs(un + pc[it]);
				}
			}
		}
		else
		{
			for (int r = ty; r < nrow; r += XF_TF)
			{
				const int q = f0 - P + r; // index along the sweep of this staged row
				const bool ok = iok && q >= 0 && q < nmax;
				const long long id = DIR == 1 ? ((long long)k * d.Ymax + q) * d.Xp + i : ((long long)q * d.Ymax + j) * d.Xp + i;
				stage_cell<C, DIR>(d, U, ok ? id : 0, ok, sU, sF, sL, ncell, r * XF_TW + tx);
			}
		}
		if (XF_SIDE_EARLY == 2 && valid)
			xf_roe_state<C>(sl, sr, d.gamma0, R);
		__syncthreads();
		st.sU = sU, st.sF = sF, st.sL = sL, st.ncell = ncell, st.base = ty * XF_TW + tx, st.stride = XF_TW;
		if (!XF_SIDE_EARLY && valid)
			load_side<C>(d, U, id_l, sl), load_side<C>(d, U, id_l + sS, sr);
	}
	if (!valid && !(DIR == 0 && mode != XF_MODE_FW))
		return;
	double F[E];
	if (valid)
	{
		if (XF_SIDE_EARLY != 2)
			xf_roe_state<C>(sl, sr, d.gamma0, R);
		double glf[3] = {d.red[XF_RED_GLF + DIR * 3 + 0], d.red[XF_RED_GLF + DIR * 3 + 1], d.red[XF_RED_GLF + DIR * 3 + 2]};
		xf_face_flux<C, DIR, WENO>(st, R, d.alpha, glf, DIR == 0 ? d.dx : (DIR == 1 ? d.dy : d.dz), F);
		if constexpr (PP)
		{ // (a template parameter: as a run-time branch the limiter cost the unlimited sweeps 2-4 %)  PositivityPreservingKernel runs over the inner cells (ConVenction_block.hpp:330-410): the face below the first inner
		  // cell (the first face of every pencil) is never limited
			bool lim;
			if constexpr (DIR == 0)
				lim = ipos >= d.Bx;
			else
				lim = (sweep_tile * XF_TF + threadIdx.x / XF_TW) > 0;
			if (lim)
				xf_positivity<C, WENO>(st, d.red[XF_RED_PPL + DIR], d.CFL, F);
		}
	}
	if constexpr (DIR == 0 && ACC)
	{
		{ // UpdateFluidLU, x part (Reconstruction_kernels.hpp:219-221): the left face's flux comes from the neighbouring thread through the
		  // (now idle) stencil buffer; LU0 = 0.0; LU0 += (F_{i-1} - F_i) * _dx
			__syncthreads();
			double *ex = smem;
			if (valid)
			{
#pragma unroll
				for (int n = 0; n < E; n++)
					ex[n * XF_TX + threadIdx.x] = F[n];
			}
			__syncthreads();
			if (valid && threadIdx.x >= 1 && ipos >= d.Bx)
			{
#pragma unroll
				for (int n = 0; n < E; n++)
					LU[n * d.N + id_l] = 0.0 + (ex[n * XF_TX + threadIdx.x - 1] - F[n]) * d._dx;
			}
		}
	}
	else
	{
		if constexpr (VISC)
		{ // GetWallViscousFlux{X,Y,Z} (ConVenction_block.hpp:506-575): Flux_wall -= F_wall_v, after the limiter, on the flux still in registers
			double Fv[E];
			visc_face_flux<C, DIR>(d, vf, U, id_l, Fv);
#pragma unroll
			for (int n = 0; n < E; n++)
				F[n] -= Fv[n];
		}
#pragma unroll
		for (int n = 0; n < E; n++)
			Fw[n * d.N + id_l] = F[n];
	}
}

// ---------------------------------------------------------------------------------------------
// k_lu: UpdateFluidLU (Reconstruction_kernels.hpp:201-234); k_rk: UpdateURK3rdKernel (Update_kernels.hpp:64-94) with
// EstimateFluidNANKernel (Fluids.cpp:47-87) folded in; k_rk<E, true>: both, LU kept in registers.
// One thread per inner cell, x fastest.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool inner_cell(const XfDev &d, long long &id)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int i = int(t % d.Xi);
	const long long r = t / d.Xi;
	const int j = int(r % d.Yi);
	const long long k = r / d.Yi;
	if (k >= d.Zi)
		return false;
	id = ((k + d.Bz) * d.Ymax + (j + d.By)) * d.Xp + (i + d.Bx);
	return true;
}
// Block order of the fused divergence + update: consecutive blocks walk along z for one (x chunk, y) column, so that the
// lower z-face flux a block needs (Fw_z at k-1) was read by the block just before it and the lower y-face flux by a block
// Zi launches earlier (~28 MB of traffic at 512^3) -- both still in L2.  With the x-fastest order of inner_cell the reuse
// distance of Fw_z is one whole plane of every array (~120 MB) and the k-1 values come from DRAM a second time.
#ifndef XF_RK_TILE
#define XF_RK_TILE 8 // > 0: walk the (y, z) plane of blocks in XF_RK_TILE x XF_RK_TILE tiles (z fastest inside a tile, tiles y fastest); 0: z-fastest columns.  Measured 512x512x256: 18.0 -> 15.0 ms per step (tile 4, 8, 16 alike)
#endif
// ka, nz: inner z-planes [ka, ka + nz) of the block (0 <= ka, ka + nz <= Zi)
__device__ __forceinline__ bool inner_cell_zfast(const XfDev &d, long long &id, int ka, int nz)
{
	const long long b = blockIdx.x;
	int j, k, xc;
	if (XF_RK_TILE > 0)
	{
		constexpr int TT = XF_RK_TILE > 0 ? XF_RK_TILE : 1;
		// (a run-time tile height for thin plane ranges / 2-D blocks was tried: no gain in 2-D, and the extra code path made the
		// compiler hold fewer loads in flight in the 3-D kernel, 29.8 -> 36.0 ms per step; empty tile slots exit at once)
		const int nty = (d.Yi + TT - 1) / TT, ntz = (nz + TT - 1) / TT;
		const unsigned per_x = unsigned(nty) * unsigned(ntz) * (TT * TT);
		xc = int(b / per_x);
		const unsigned r = unsigned(b - (long long)xc * per_x);
		const unsigned tile = r / (TT * TT), w = r % (TT * TT);
		k = int(tile / nty) * TT + int(w % TT);
		j = int(tile % nty) * TT + int(w / TT);
		if (j >= d.Yi || k >= nz)
			return false;
	}
	else
	{
		k = int(b % nz);
		const long long r = b / nz;
		j = int(r % d.Yi);
		xc = int(r / d.Yi);
	}
	const int i = xc * blockDim.x + threadIdx.x;
	if (i >= d.Xi)
		return false;
	id = ((long long)(k + ka + d.Bz) * d.Ymax + (j + d.By)) * d.Xp + (i + d.Bx);
	return true;
}
__device__ __forceinline__ double lu_of(const XfDev &d, long long o)
{
	double LU0 = 0.0;
	if (d.DimX)
		LU0 += (d.Fw[0][o - 1] - d.Fw[0][o]) * d._dx;
	if (d.DimY)
		LU0 += (d.Fw[1][o - d.sY] - d.Fw[1][o]) * d._dy;
	if (d.DimZ)
		LU0 += (d.Fw[2][o - d.sZ] - d.Fw[2][o]) * d._dz;
	return LU0;
}
template <int E>
__global__ void __launch_bounds__(256) k_lu(XfDev d, double *__restrict__ LU)
{
	long long id;
	if (!inner_cell(d, id))
		return;
#pragma unroll
	for (int n = 0; n < E; n++)
		LU[n * d.N + id] = lu_of(d, n * d.N + id);
}
__device__ __forceinline__ double rk_of(double U, double U1, double LU, double dt, int flag)
{
	if (flag == 1)
		return U + dt * LU;
	if (flag == 2)
		return 0.75 * U + 0.25 * U1 + 0.25 * dt * LU;
	return (U + 2.0 * U1 + 2.0 * dt * LU) * (1.0 / 3.0);
}
// dt_dev != nullptr: read dt from the device (graph-replayable); guard: check UI (U1,U1,U for flag 1,2,3) and LU
template <int E, bool FUSED_LU>
__global__ void __launch_bounds__(256) k_rk(XfDev d, double *__restrict__ U, double *__restrict__ U1, const double *__restrict__ LU,
											double dt_host, const double *__restrict__ dt_dev, int flag, int guard, int ka, int nz)
{
	long long id;
	if (FUSED_LU ? !inner_cell_zfast(d, id, ka, nz) : !inner_cell(d, id))
		return;
	const double dt = dt_dev ? *dt_dev : dt_host;
	bool bad = false;
#pragma unroll
	for (int n = 0; n < E; n++)
	{
		const long long o = n * d.N + id;
		const double lu = FUSED_LU ? lu_of(d, o) : LU[o];
		const double u0 = U[o];
		const double u1 = (flag == 1) ? 0.0 : U1[o];
		if (guard)
		{
			const double ui = (flag == 3) ? u0 : ((flag == 1) ? U1[o] : u1);
			bad = bad || isnan(ui) || isinf(ui) || isnan(lu) || isinf(lu) || (n == 0 && ui < 0);
		}
		const double r = rk_of(u0, u1, lu, dt, flag);
		if (flag == 3)
			U[o] = r;
		else
			U1[o] = r;
	}
	if (guard && bad)
		d.err[2] = 1;
}
// k_rk_prim: the stage update of k_rk<E, true> going straight on to the NEXT stage's primitive recovery for the deep cells (XfDeep), whose
// updated conserved variables are still in registers.  Shell cells are only updated; BC fill and k_prim_shell follow.  nflags: the
// gather flags of the next stage's recovery.  Same block order as k_rk (the wall fluxes of the lower faces come out of L2).
#ifndef XF_RKP_MINB
#define XF_RKP_MINB 4
#endif
template <class C>
__global__ void __launch_bounds__(128, XF_RKP_MINB) k_rk_prim(XfDev d, XfThermo th, XfDeep dp, double *__restrict__ U, double *__restrict__ U1,
															   const double *__restrict__ dt_dev, int flag, int guard, int nflags)
{
	constexpr int E = C::E;
	long long id;
	double dtm[3] = {0.0, 0.0, 0.0}, glf[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	if (inner_cell_zfast(d, id, 0, d.Zi))
	{
		const int i = int(id % d.Xp);
		const long long row = id / d.Xp;
		const int j = int(row % d.Ymax), k = int(row / d.Ymax);
		const bool deep = i >= dp.xlo && i < dp.xhi && j >= dp.ylo && j < dp.yhi && k >= dp.zlo && k < dp.zhi;
		const double dt = *dt_dev;
		double *__restrict__ UO = (flag == 3) ? U : U1;
		double Un[E], Tc = 0.0;
		if constexpr (C::COP)
			if (deep)
				Tc = d.T[id];
		bool bad = false;
#pragma unroll
		for (int n = 0; n < E; n++)
		{
			const long long o = n * d.N + id;
			const double lu = lu_of(d, o);
			const double u0 = U[o];
			const double u1 = (flag == 1) ? 0.0 : U1[o];
			if (guard)
			{
				const double ui = (flag == 3) ? u0 : ((flag == 1) ? U1[o] : u1);
				bad = bad || isnan(ui) || isinf(ui) || isnan(lu) || isinf(lu) || (n == 0 && ui < 0);
			}
			Un[n] = rk_of(u0, u1, lu, dt, flag);
			if (!(C::COP && n >= 5 && d.ghost && deep)) // GhostSpecies: the recovery stores the renormalised partial densities
				UO[o] = Un[n];
		}
		if (guard && bad)
			d.err[2] = 1;
		if (deep)
			prim_cell<C>(d, th, UO, id, i, j, k, Un, Tc, nflags, dtm, glf);
	}
	prim_reduce(d, nflags, dtm, glf);
}
template <int E>
__global__ void __launch_bounds__(256) k_nan(XfDev d, const double *__restrict__ UI, const double *__restrict__ LU)
{
	long long id;
	if (!inner_cell(d, id))
		return;
	bool bad = UI[id] < 0;
#pragma unroll
	for (int n = 0; n < E; n++)
	{
		const double a = UI[n * d.N + id], b = LU[n * d.N + id];
		bad = bad || isnan(a) || isinf(a) || isnan(b) || isinf(b);
	}
	if (bad)
		d.err[2] = 1;
}

#include "xf_march.cuh"

// ---------------------------------------------------------------------------------------------
// k_bc: FluidBCKernel{X,Y,Z} (BCs_kernels.hpp:9-258) launched as in BCs_block.cpp:81-213: one thread per ghost
// index g < Bw and transverse position, handling the min face then the max face.  BC_COPY faces are skipped
// (filled by the halo exchange).
// ---------------------------------------------------------------------------------------------
template <int E, int DIR>
__device__ __forceinline__ void bc_apply(const XfDev &d, double *__restrict__ U, int BC, int i, int j, int k,
										 int mirror_offset, int index_inner, int sign, bool cop)
{
	const int Bw = DIR == 0 ? d.Bx : (DIR == 1 ? d.By : d.Bz);
	const int inner = DIR == 0 ? d.Xi : (DIR == 1 ? d.Yi : d.Zi);
	const int g = DIR == 0 ? i : (DIR == 1 ? j : k);
	const long long id = ((long long)k * d.Ymax + j) * d.Xp + i;
	auto tid = [&](int t) -> long long
	{ return DIR == 0 ? ((long long)k * d.Ymax + j) * d.Xp + t : (DIR == 1 ? ((long long)k * d.Ymax + t) * d.Xp + i : ((long long)t * d.Ymax + j) * d.Xp + i); };
	switch (BC)
	{
	case 2:
	{ // Symmetry
		const long long t = tid(2 * (Bw + mirror_offset) - 1 - g);
#pragma unroll
		for (int n = 0; n < E; n++)
			U[n * d.N + id] = (n == 1 + DIR) ? -U[n * d.N + t] : U[n * d.N + t];
	}
	break;
	case 3:
	{ // Periodic
		const long long t = tid(g + sign * inner);
#pragma unroll
		for (int n = 0; n < E; n++)
			U[n * d.N + id] = U[n * d.N + t];
	}
	break;
	case 1:
	{ // Outflow
		const long long t = tid(index_inner);
#pragma unroll
		for (int n = 0; n < E; n++)
			U[n * d.N + id] = U[n * d.N + t];
	}
	break;
	case 4:
	case 5:
	case 6:
	{ // nslipWall everywhere; viscWall / slipWall in X only (no-ops in Y,Z: BCs_kernels.hpp:167-171,242-246)
		if (BC != 4 && DIR != 0)
			break;
		const long long t = tid(2 * (Bw + mirror_offset) - 1 - g);
#pragma unroll
		for (int n = 0; n < E; n++)
		{
			if (n >= 5 && !cop)
				continue;
			U[n * d.N + id] = (n >= 1 && n <= 3) ? -U[n * d.N + t] : U[n * d.N + t];
		}
	}
	break;
	default: // Inflow, innerBlock, BC_COPY
		break;
	}
}
template <int E, int DIR>
__global__ void __launch_bounds__(256) k_bc(XfDev d, double *__restrict__ U, int bc_min, int bc_max, int cop, int k0, int k1 /* z-planes [k0, k1) for DIR 0, 1 */)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int i, j, k, g;
	if constexpr (DIR == 0)
	{ // threads over (j,k) x g ; g fastest is pointless for coalescing here, ghost columns are 4 wide
		g = int(t % d.Bx);
		const long long r = t / d.Bx;
		j = int(r % d.Ymax), k = k0 + int(r / d.Ymax);
		if (k >= k1)
			return;
		bc_apply<E, 0>(d, U, bc_min, g, j, k, 0, d.Bx, 1, cop);
		bc_apply<E, 0>(d, U, bc_max, g + d.Xmax - d.Bx, j, k, d.Xi, d.Xmax - d.Bx - 1, -1, cop);
	}
	else if constexpr (DIR == 1)
	{
		i = int(t % d.Xmax);
		const long long r = t / d.Xmax;
		g = int(r % d.By), k = k0 + int(r / d.By);
		if (k >= k1)
			return;
		bc_apply<E, 1>(d, U, bc_min, i, g, k, 0, d.By, 1, cop);
		bc_apply<E, 1>(d, U, bc_max, i, g + d.Ymax - d.By, k, d.Yi, d.Ymax - d.By - 1, -1, cop);
	}
	else
	{
		i = int(t % d.Xmax);
		const long long r = t / d.Xmax;
		j = int(r % d.Ymax), g = int(r / d.Ymax);
		if (g >= d.Bz)
			return;
		bc_apply<E, 2>(d, U, bc_min, i, j, g, 0, d.Bz, 1, cop);
		bc_apply<E, 2>(d, U, bc_max, i, j, g + d.Zmax - d.Bz, d.Zi, d.Zmax - d.Bz - 1, -1, cop);
	}
}

// ---------------------------------------------------------------------------------------------
// layout kernels: reference AoS [cell][E] (unpadded) <-> SoA [E][cell] (x-padded), staged through shared memory so
// that both sides are coalesced; scalar arrays padded/unpadded; z-slab halo pack/unpack.
// ---------------------------------------------------------------------------------------------
template <int E, bool TO_SOA>
__global__ void __launch_bounds__(128) k_layout(XfDev d, double *__restrict__ soa, double *__restrict__ aos, long long row0)
{
	__shared__ double tile[128 * E];
	const long long row = row0 + blockIdx.x;          // k*Ymax + j  (grid.x: up to 2^31-1 rows)
	const int i0 = blockIdx.y * 128;
	const int ni = min(128, d.Xmax - i0);
	const long long abase = (row * d.Xmax + i0) * E;  // AoS doubles
	const long long sbase = row * d.Xp + i0;
	if (TO_SOA)
	{
		for (int q = threadIdx.x; q < ni * E; q += 128)
			tile[q] = aos[abase + q];
		__syncthreads();
		if ((int)threadIdx.x < ni)
#pragma unroll
			for (int n = 0; n < E; n++)
				soa[n * d.N + sbase + threadIdx.x] = tile[threadIdx.x * E + n];
	}
	else
	{
		if ((int)threadIdx.x < ni)
#pragma unroll
			for (int n = 0; n < E; n++)
				tile[threadIdx.x * E + n] = soa[n * d.N + sbase + threadIdx.x];
		__syncthreads();
		for (int q = threadIdx.x; q < ni * E; q += 128)
			aos[abase + q] = tile[q];
	}
}
__global__ void k_scalar_pad(XfDev d, double *__restrict__ padded, double *__restrict__ flat, int to_padded)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int i = int(t % d.Xmax);
	const long long row = t / d.Xmax;
	if (row >= (long long)d.Ymax * d.Zmax)
		return;
	if (to_padded)
		padded[row * d.Xp + i] = flat[t];
	else
		flat[t] = padded[row * d.Xp + i];
}
// halo buffer layout: [E][Bz planes][Ymax][Xp] ; face 4: inner planes Bz..2Bz-1 (pack) / ghosts 0..Bz-1 (unpack),
// face 5: inner planes Zmax-2Bz..Zmax-Bz-1 (pack) / ghosts Zmax-Bz..Zmax-1 (unpack)
template <int E>
__global__ void __launch_bounds__(256) k_halo(XfDev d, double *__restrict__ U, double *__restrict__ buf, int k0, int pack)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long slab = (long long)d.Bz * d.sZ;
	if (t >= slab)
		return;
#pragma unroll
	for (int n = 0; n < E; n++)
	{
		if (pack)
			buf[n * slab + t] = U[n * d.N + (long long)k0 * d.sZ + t];
		else
			U[n * d.N + (long long)k0 * d.sZ + t] = buf[n * slab + t];
	}
}

// =================================================================================================
//  launchers (C++ linkage inside the namespace; selected at run time by the C-ABI layer)
// =================================================================================================
#define XF_CHECK_LAUNCH()                         \
	do                                            \
	{                                             \
		cudaError_t e__ = cudaGetLastError();     \
		if (e__ != cudaSuccess)                   \
			return (int)e__;                      \
	} while (0)

#include "xf_visc.cuh"

template <class C>
static int prim_t(const XfDev &d, const XfThermo &th, double *U, int flags, cudaStream_t s, long long *launches, int k0, int k1)
{
	// z-planes [k0, k1) of the block (all of it: 0, Zmax)
	if (k1 <= k0)
		return 0;
	constexpr int CHB = 128 * (XF_PRIM_PIPE > 0 ? XF_PRIM_PIPE : 1);
	const dim3 nb((unsigned)((d.sZ + CHB - 1) / CHB), (unsigned)(k1 - k0));
	if constexpr (C::COP)
	{
		cudaError_t e = cudaMemsetAsync(d.hard_count, 0, sizeof(unsigned), s);
		if (e != cudaSuccess)
			return (int)e;
	}
	k_prim<C><<<nb, 128, 0, s>>>(d, th, U, flags, (long long)k0, (long long)k1);
	XF_CHECK_LAUNCH();
	++*launches;
	if constexpr (C::COP)
	{
		k_prim_hard<C><<<d.nsm * 8, 128, 0, s>>>(d, th, U, flags);
		XF_CHECK_LAUNCH();
		++*launches;
	}
	return 0;
}

// (the opt-in to > 48 KB of dynamic shared memory is per device and per kernel: set on every launch -- microseconds, no stream
// operation, legal during graph capture -- rather than behind a per-process flag that would only cover the first device used)
template <class C, int DIR, int WENO, bool PP>
static int sweep_pp_t(const XfDev &d, const double *U, cudaStream_t s, int kp0, int kp1, int mode = XF_MODE_FW, double *LU = nullptr, const CUtensorMap *tm = nullptr,
					  const XfViscF *vf = nullptr)
{
	static const CUtensorMap no_map{};
	if (kp1 <= kp0)
		return 0;
	constexpr int E = C::E, NST = XfStencil<WENO>::NST;
	if constexpr (DIR == 0)
	{
		constexpr int XF_TX = XfTx<WENO, PP>::V;
		constexpr size_t smem = size_t(2 * E + 3) * (XF_TX + NST - 1) * sizeof(double);
		cudaFuncSetAttribute(k_sweep<C, DIR, WENO, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		const int cells = (mode == XF_MODE_FW) ? XF_TX : XF_TX - 1; // ACC: chunks overlap by one face
		const dim3 g((unsigned)((d.sZ + cells - 1) / cells), (unsigned)(kp1 - kp0));
		if (mode == XF_MODE_FW && vf)
		{
			cudaFuncSetAttribute(k_sweep<C, DIR, WENO, PP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			k_sweep<C, DIR, WENO, PP, false, true><<<g, XF_TX, smem, s>>>(d, no_map, U, d.Fw[0], kp0, LU, *vf);
		}
		else if (mode == XF_MODE_FW)
			k_sweep<C, DIR, WENO, PP, false><<<g, XF_TX, smem, s>>>(d, no_map, U, d.Fw[0], kp0, LU);
		else
		{
			cudaFuncSetAttribute(k_sweep<C, DIR, WENO, PP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			k_sweep<C, DIR, WENO, PP, true><<<g, XF_TX, smem, s>>>(d, no_map, U, d.Fw[0], kp0, LU);
		}
	}
	else
	{
		constexpr size_t smem = size_t(2 * E + 3) * (XF_TF + NST - 1) * XF_TW * sizeof(double) + 16; // + the TMA mbarrier
		if (XF_SWEEP_TMA && !tm)
			return -1;
		cudaFuncSetAttribute(k_sweep<C, DIR, WENO, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		dim3 g;
		if (DIR == 1)
			g.x = (d.Xi + XF_TW - 1) / XF_TW, g.y = (d.Yi + 1 + XF_TF - 1) / XF_TF, g.z = kp1 - kp0;
		else
			g.x = (d.Xi + XF_TW - 1) / XF_TW, g.y = kp1 - kp0, g.z = d.Yi; // kp0, kp1: tile range of the z sweep
		if (vf)
		{
			cudaFuncSetAttribute(k_sweep<C, DIR, WENO, PP, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
			k_sweep<C, DIR, WENO, PP, false, true><<<g, XF_TW * XF_TF, smem, s>>>(d, tm ? *tm : no_map, U, d.Fw[DIR], kp0, nullptr, *vf);
		}
		else
			k_sweep<C, DIR, WENO, PP><<<g, XF_TW * XF_TF, smem, s>>>(d, tm ? *tm : no_map, U, d.Fw[DIR], kp0, nullptr);
		(void)kp1;
	}
	XF_CHECK_LAUNCH();
	return 0;
}
template <class C, int DIR, int WENO>
static int sweep_t(const XfDev &d, const double *U, cudaStream_t s, int kp0, int kp1, int mode = XF_MODE_FW, double *LU = nullptr, const CUtensorMap *tm = nullptr,
				   const XfViscF *vf = nullptr)
{
#ifdef XF_ONLY_SBI
	if (d.positivity || WENO != 5)
		return -1;
	if constexpr (WENO == 5)
		return sweep_pp_t<C, DIR, WENO, false>(d, U, s, kp0, kp1, mode, LU, tm, vf);
#else
	return d.positivity ? sweep_pp_t<C, DIR, WENO, true>(d, U, s, kp0, kp1, mode, LU, tm, vf) : sweep_pp_t<C, DIR, WENO, false>(d, U, s, kp0, kp1, mode, LU, tm, vf);
#endif
}
// x sweep of the fused path: z-planes [a.t0, a.t1), a.mode = XF_MODE_ACC (LU = 0.0 + d/dx part) or XF_MODE_FW
template <class C>
static int sweep_x_t(const XfDev &d, const double *U, const XfMarchArgs &a, cudaStream_t s)
{
	if (d.weno == 7)
		return sweep_t<C, 0, 7>(d, U, s, a.t0, a.t1, a.mode, a.LU);
	if (d.weno == 6)
		return sweep_t<C, 0, 6>(d, U, s, a.t0, a.t1, a.mode, a.LU);
	return sweep_t<C, 0, 5>(d, U, s, a.t0, a.t1, a.mode, a.LU);
}
// marching y / z sweep (xf_march.cuh)
template <class C, int DIR, int WENO, bool PP>
static int march_pp_t(const XfDev &d, const XfTma &tm, const double *UI, const XfMarchArgs &a, cudaStream_t s)
{
	using G = XfMarchGeom<C, WENO>;
	if (a.t1 <= a.t0 || a.cb <= a.ca)
		return 0;
	cudaFuncSetAttribute(k_march<C, DIR, WENO, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::smem_bytes);
	const dim3 g((unsigned)((d.Xi + G::W - 1) / G::W), (unsigned)(a.t1 - a.t0), (unsigned)a.nseg);
	k_march<C, DIR, WENO, PP><<<g, G::NT, G::smem_bytes, s>>>(d, tm.U, tm.P, tm.Y, UI, a);
	XF_CHECK_LAUNCH();
	return 0;
}
template <class C, int DIR>
static int march_t(const XfDev &d, const XfTma &tm, const double *UI, const XfMarchArgs &a, cudaStream_t s)
{
#ifdef XF_ONLY_SBI
	if (d.weno != 5 || d.positivity)
		return -1;
	return march_pp_t<C, DIR, 5, false>(d, tm, UI, a, s);
#endif
	if (d.weno == 7)
		return d.positivity ? march_pp_t<C, DIR, 7, true>(d, tm, UI, a, s) : march_pp_t<C, DIR, 7, false>(d, tm, UI, a, s);
	if (d.weno == 6)
		return d.positivity ? march_pp_t<C, DIR, 6, true>(d, tm, UI, a, s) : march_pp_t<C, DIR, 6, false>(d, tm, UI, a, s);
	return d.positivity ? march_pp_t<C, DIR, 5, true>(d, tm, UI, a, s) : march_pp_t<C, DIR, 5, false>(d, tm, UI, a, s);
}
// kp0, kp1: z-plane range [kp0, kp1) of the x and y sweeps (absolute plane indices inside [Bz, Bz + Zi)); the z sweep always
// covers the block
template <class C>
static int sweeps_t(const XfDev &d, const double *U, cudaStream_t s, long long *launches, int dirmask, int kp0, int kp1, int tz0, int tz1, const CUtensorMap *tmy, const CUtensorMap *tmz,
					const XfViscF *vf)
{
	int rc = 0;
	const bool dx = d.DimX && (dirmask & 1), dy = d.DimY && (dirmask & 2), dz = d.DimZ && (dirmask & 4);
	if (d.weno == 7)
	{
		if (dx) rc |= sweep_t<C, 0, 7>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, nullptr, vf), ++*launches;
		if (dy) rc |= sweep_t<C, 1, 7>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, tmy, vf), ++*launches;
		if (dz) rc |= sweep_t<C, 2, 7>(d, U, s, tz0, tz1, XF_MODE_FW, nullptr, tmz, vf), ++*launches;
	}
	else if (d.weno == 6)
	{
		if (dx) rc |= sweep_t<C, 0, 6>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, nullptr, vf), ++*launches;
		if (dy) rc |= sweep_t<C, 1, 6>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, tmy, vf), ++*launches;
		if (dz) rc |= sweep_t<C, 2, 6>(d, U, s, tz0, tz1, XF_MODE_FW, nullptr, tmz, vf), ++*launches;
	}
	else
	{
		if (dx) rc |= sweep_t<C, 0, 5>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, nullptr, vf), ++*launches;
		if (dy) rc |= sweep_t<C, 1, 5>(d, U, s, kp0, kp1, XF_MODE_FW, nullptr, tmy, vf), ++*launches;
		if (dz) rc |= sweep_t<C, 2, 5>(d, U, s, tz0, tz1, XF_MODE_FW, nullptr, tmz, vf), ++*launches;
	}
	return rc;
}

#ifdef XF_ONLY_SBI // tuning builds: only the Emax = 9 configuration is instantiated (seconds instead of minutes per variant)
#define XF_DISPATCH_CFG(ns, cop, ...)                          \
	if ((cop) && (ns) == 5) { using C = XfCfg<5, true>; __VA_ARGS__; } \
	else return -1;
#else
#define XF_DISPATCH_CFG(ns, cop, ...)                          \
	if (!(cop)) { using C = XfCfg<1, false>; __VA_ARGS__; }    \
	else if ((ns) == 2) { using C = XfCfg<2, true>; __VA_ARGS__; } \
	else if ((ns) == 3) { using C = XfCfg<3, true>; __VA_ARGS__; } \
	else if ((ns) == 4) { using C = XfCfg<4, true>; __VA_ARGS__; } \
	else if ((ns) == 5) { using C = XfCfg<5, true>; __VA_ARGS__; } \
	else return -1;
#endif

#define XF_DISPATCH_E(E_, ...)                          \
	switch (E_)                                         \
	{                                                   \
	case 5: { constexpr int E = 5; __VA_ARGS__; } break; \
	case 6: { constexpr int E = 6; __VA_ARGS__; } break; \
	case 7: { constexpr int E = 7; __VA_ARGS__; } break; \
	case 8: { constexpr int E = 8; __VA_ARGS__; } break; \
	case 9: { constexpr int E = 9; __VA_ARGS__; } break; \
	default: return -1;                                 \
	}

// primitive recovery of every cell outside the deep set (after the ghost fill); the list of unconverged cells is NOT reset: the deep
// cells' entries (k_rk_prim) are still in it, k_prim_hard finishes both
template <class C>
static int prim_shell_t(const XfDev &d, const XfThermo &th, double *U, int flags, cudaStream_t s, long long *launches)
{
	const XfDeep dp = xf_deep_cells(d);
	constexpr int CHB = 128 * (XF_PRIM_PIPE > 0 ? XF_PRIM_PIPE : 1);
	const int nslab = dp.zlo + (d.Zmax - dp.zhi), nmid = dp.zhi - dp.zlo;
	if (nslab > 0)
	{
		k_prim<C><<<dim3((unsigned)((d.sZ + CHB - 1) / CHB), (unsigned)nslab), 128, 0, s>>>(d, th, U, flags, 0LL, 0LL, dp.zlo, nmid);
		XF_CHECK_LAUNCH();
		++*launches;
	}
	if (nmid > 0)
	{
		const long long per = (long long)(dp.ylo + (d.Ymax - dp.yhi)) * d.Xp + (long long)(dp.yhi - dp.ylo) * (dp.xlo + (d.Xmax - dp.xhi));
		if (per > 0)
		{
			k_prim_shell<C><<<dim3((unsigned)((per + 127) / 128), (unsigned)nmid), 128, 0, s>>>(d, th, dp, U, flags);
			XF_CHECK_LAUNCH();
			++*launches;
		}
	}
	if constexpr (C::COP)
	{
		k_prim_hard<C><<<d.nsm * 8, 128, 0, s>>>(d, th, U, flags);
		XF_CHECK_LAUNCH();
		++*launches;
	}
	return 0;
}
template <class C>
static int rk_prim_t(const XfDev &d, const XfThermo &th, double *U, double *U1, const double *dt_dev, int flag, int guard, int nflags, cudaStream_t s, long long *launches)
{
	if constexpr (C::COP)
	{
		cudaError_t e = cudaMemsetAsync(d.hard_count, 0, sizeof(unsigned), s);
		if (e != cudaSuccess)
			return (int)e;
	}
	constexpr int TT = XF_RK_TILE > 0 ? XF_RK_TILE : 1;
	long long nb = (long long)((d.Xi + 127) / 128) * d.Yi * d.Zi;
	if (XF_RK_TILE > 0)
		nb = (long long)((d.Xi + 127) / 128) * ((d.Yi + TT - 1) / TT) * ((d.Zi + TT - 1) / TT) * (TT * TT);
	k_rk_prim<C><<<(unsigned)nb, 128, 0, s>>>(d, th, xf_deep_cells(d), U, U1, dt_dev, flag, guard, nflags);
	XF_CHECK_LAUNCH();
	++*launches;
	return 0;
}
int launch_prim_shell(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, int flags, cudaStream_t s, long long *launches)
{
	XF_DISPATCH_CFG(ns, cop, return prim_shell_t<C>(d, th, U, flags, s, launches));
}
int launch_rk_prim(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, double *U1, const double *dt_dev, int flag, int guard, int nflags, cudaStream_t s,
				   long long *launches)
{
	XF_DISPATCH_CFG(ns, cop, return rk_prim_t<C>(d, th, U, U1, dt_dev, flag, guard, nflags, s, launches));
}
int launch_prim(const XfDev &d, const XfThermo &th, int ns, int cop, double *U, int flags, cudaStream_t s, long long *launches, int k0, int k1)
{
	XF_DISPATCH_CFG(ns, cop, return prim_t<C>(d, th, U, flags, s, launches, k0, k1));
}
// kp0, kp1: z-plane range of the x / y sweeps (< 0: all inner planes); tz0, tz1: tile range of the z sweep (tile t = faces
// Bz - 1 + XF_TF t ... ; < 0: all xf_z_tiles() of them)
int xf_z_tiles(const XfDev &d) { return (d.Zi + 1 + XF_TF - 1) / XF_TF; }
int launch_sweeps(const XfDev &d, int ns, int cop, const double *U, cudaStream_t s, long long *launches, int dirmask, int kp0, int kp1, int tz0, int tz1,
				  const CUtensorMap *tmy, const CUtensorMap *tmz, const XfViscF *vf)
{
	if (kp0 < 0)
		kp0 = d.Bz, kp1 = d.Bz + d.Zi;
	if (tz0 < 0)
		tz0 = 0, tz1 = xf_z_tiles(d);
	XF_DISPATCH_CFG(ns, cop, return sweeps_t<C>(d, U, s, launches, dirmask, kp0, kp1, tz0, tz1, tmy, tmz, vf));
}
int z_tile_faces() { return XF_TF; }
int launch_visc(const XfDev &d, const XfThermo &th, const XfVisc &vs, int ns, int cop, const double *U, const int bc[6], cudaStream_t s, long long *launches, int parts)
{
	XF_DISPATCH_CFG(ns, cop, return visc_t<C>(d, th, vs, U, bc, s, launches, parts));
}
int launch_sweep_x(const XfDev &d, int ns, int cop, const double *U, const XfMarchArgs &a, cudaStream_t s)
{
	XF_DISPATCH_CFG(ns, cop, return sweep_x_t<C>(d, U, a, s));
}
int launch_march(const XfDev &d, int ns, int cop, const XfTma &tm, const double *UI, const XfMarchArgs &a, int dir, cudaStream_t s)
{
	if (dir == 1)
	{
		XF_DISPATCH_CFG(ns, cop, return march_t<C, 1>(d, tm, UI, a, s));
	}
	XF_DISPATCH_CFG(ns, cop, return march_t<C, 2>(d, tm, UI, a, s));
}
static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }
int launch_lu(const XfDev &d, int E_, double *LU, cudaStream_t s)
{
	const long long n = (long long)d.Xi * d.Yi * d.Zi;
	XF_DISPATCH_E(E_, k_lu<E><<<nblk(n, 256), 256, 0, s>>>(d, LU));
	XF_CHECK_LAUNCH();
	return 0;
}
// fused: ka, kb = inner z-planes [ka, kb) to update (ka < 0: all)
int launch_rk(const XfDev &d, int E_, double *U, double *U1, const double *LU, double dt, const double *dt_dev, int flag, int guard, int fused, cudaStream_t s,
			  int ka, int kb)
{
	const long long n = (long long)d.Xi * d.Yi * d.Zi;
	if (ka < 0)
		ka = 0, kb = d.Zi;
	const int nz = kb - ka;
	if (fused)
	{
		if (nz <= 0)
			return 0;
		long long nb = (long long)((d.Xi + 255) / 256) * d.Yi * nz; // one block per (x chunk, y, z), z fastest
		constexpr int TT = XF_RK_TILE > 0 ? XF_RK_TILE : 1;
		if (XF_RK_TILE > 0)
			nb = (long long)((d.Xi + 255) / 256) * ((d.Yi + TT - 1) / TT) * ((nz + TT - 1) / TT) * (TT * TT);
		XF_DISPATCH_E(E_, k_rk<E, true><<<(unsigned)nb, 256, 0, s>>>(d, U, U1, LU, dt, dt_dev, flag, guard, ka, nz));
	}
	else
	{
		XF_DISPATCH_E(E_, k_rk<E, false><<<nblk(n, 256), 256, 0, s>>>(d, U, U1, LU, dt, dt_dev, flag, guard, 0, d.Zi));
	}
	XF_CHECK_LAUNCH();
	return 0;
}
int launch_nan(const XfDev &d, int E_, const double *UI, const double *LU, cudaStream_t s)
{
	const long long n = (long long)d.Xi * d.Yi * d.Zi;
	XF_DISPATCH_E(E_, k_nan<E><<<nblk(n, 256), 256, 0, s>>>(d, UI, LU));
	XF_CHECK_LAUNCH();
	return 0;
}
// dirmask: bit d = fill direction d; k0, k1: z-plane range of the x and y fills (k0 < 0: all planes)
int launch_bc(const XfDev &d, int E_, int cop, double *U, const int bc[6], cudaStream_t s, long long *launches, int dirmask, int k0, int k1)
{
	if (k0 < 0)
		k0 = 0, k1 = d.Zmax;
	if (d.DimX && (dirmask & 1) && !(bc[0] == 0 && bc[1] == 0) && k1 > k0)
	{
		const long long n = (long long)d.Bx * d.Ymax * (k1 - k0);
		XF_DISPATCH_E(E_, k_bc<E, 0><<<nblk(n, 256), 256, 0, s>>>(d, U, bc[0], bc[1], cop, k0, k1));
		XF_CHECK_LAUNCH();
		++*launches;
	}
	if (d.DimY && (dirmask & 2) && k1 > k0)
	{
		const long long n = (long long)d.Xmax * d.By * (k1 - k0);
		XF_DISPATCH_E(E_, k_bc<E, 1><<<nblk(n, 256), 256, 0, s>>>(d, U, bc[2], bc[3], cop, k0, k1));
		XF_CHECK_LAUNCH();
		++*launches;
	}
	if (d.DimZ && (dirmask & 4))
	{
		const long long n = (long long)d.Xmax * d.Ymax * d.Bz;
		XF_DISPATCH_E(E_, k_bc<E, 2><<<nblk(n, 256), 256, 0, s>>>(d, U, bc[4], bc[5], cop, 0, 0));
		XF_CHECK_LAUNCH();
		++*launches;
	}
	return 0;
}
int launch_dt(const XfDev &d, const double *rho, cudaStream_t s)
{
	k_dt<<<nblk(d.N, 256), 256, 0, s>>>(d, rho);
	XF_CHECK_LAUNCH();
	return 0;
}
int launch_dt_final(const XfDev &d, double t_end, cudaStream_t s)
{
	k_dt_final<<<1, 1, 0, s>>>(d, t_end);
	XF_CHECK_LAUNCH();
	return 0;
}
// rows [row0, row0 + nrows) of the block (row = k * Ymax + j); nrows < 0: all of them.  `aos` is the whole block's AoS image.
int launch_layout(const XfDev &d, int E_, double *soa, double *aos, int to_soa, cudaStream_t s, long long row0, long long nrows)
{
	if (nrows < 0)
		row0 = 0, nrows = (long long)d.Ymax * d.Zmax;
	if (nrows == 0)
		return 0;
	dim3 g((unsigned)nrows, (d.Xmax + 127) / 128);
	if (to_soa)
	{
		XF_DISPATCH_E(E_, k_layout<E, true><<<g, 128, 0, s>>>(d, soa, aos, row0));
	}
	else
	{
		XF_DISPATCH_E(E_, k_layout<E, false><<<g, 128, 0, s>>>(d, soa, aos, row0));
	}
	XF_CHECK_LAUNCH();
	return 0;
}
int launch_scalar_pad(const XfDev &d, double *padded, double *flat, int to_padded, cudaStream_t s)
{
	const long long n = (long long)d.Xmax * d.Ymax * d.Zmax;
	k_scalar_pad<<<nblk(n, 256), 256, 0, s>>>(d, padded, flat, to_padded);
	XF_CHECK_LAUNCH();
	return 0;
}
int launch_halo(const XfDev &d, int E_, double *U, double *buf, int k0, int pack, cudaStream_t s)
{
	const long long n = (long long)d.Bz * d.sZ;
	XF_DISPATCH_E(E_, k_halo<E><<<nblk(n, 256), 256, 0, s>>>(d, U, buf, k0, pack));
	XF_CHECK_LAUNCH();
	return 0;
}

} // namespace XF_NS
