// xf_march.cuh -- the y / z sweeps as MARCHING pencils fed by a TMA + mbarrier ring, with the flux divergence and (in the last
// direction) the SSP-RK3 stage update fused in.  Included inside namespace XF_NS by xf_kernels.cu.
//
// Reference functions covered: ReconstructFluxY / ReconstructFluxZ (Reconstruction_kernels.hpp:73-199), GetLocalEigen,
// PositivityPreservingKernel, UpdateFluidLU (Reconstruction_kernels.hpp:201-234), EstimateFluidNANKernel (Fluids.cpp:47-87) and
// UpdateURK3rdKernel (Update_kernels.hpp:64-94).
//
// One block owns a column XF_MW cells wide in x at one transverse index and marches along the sweep direction, TF faces per
// iteration (thread (tx, ty) = face ty of the iteration at x offset tx):
//
//   TMA     the raw rows of the NEXT batch of TF cells (conserved variables, u v w p c, Y_i: 3 cp.async.bulk.tensor.4d per batch,
//           out-of-range rows / columns zero-filled by the hardware) land in shared memory while the block computes
//   ring    TF + NST - 1 rows of [U | 0.5 F | the 3 local wave speeds] per cell, the stencil of one iteration; after the faces of an
//           iteration are done the landed batch is converted in place of the TF rows that retire (every staged cell is loaded
//           and converted exactly once: the tiled kernel staged (TF + NST - 1) / TF = 1.6 rows per face row)
//   faces   Roe state, eigen-projection, WENO, back-projection in registers exactly as in k_sweep (same device functions)
//   update  the previous face's flux comes through shared memory (row ty - 1 of this iteration, or the last row of the previous
//           iteration), so  LU (+)= (F_{f-1} - F_f) * _dl  is formed in the reference's x -> y -> z association
//           (Reconstruction_kernels.hpp:219-226) and the wall flux never goes to HBM; in the last active direction the same
//           thread applies the NaN guard and the stage update, so LU does not go to HBM either.
//
// In-place safety of the stage-2 update (U1 is both the sweep input and the result): a cell is staged into the ring before its own
// face is computed and never re-read from HBM by this block; other columns never read it in this direction; the earlier
// directions have completed (stream order).  Segmented marches (2-D grids with few columns) would re-read rows of the previous
// segment, so the launcher never combines segments with an in-place update (it falls back to ACC + k_rk for that stage).
#pragma once

// stencil accessor on the ring: slot s of this thread's face lives at cell offset o[s] (row slot * XF_MW + tx)
template <int NST, int NCELL>
struct RingStencil
{
	const double *sU, *sF, *sL;
	int o[NST];
	__device__ __forceinline__ double U(int s, int n) const { return sU[n * NCELL + o[s]]; }
	__device__ __forceinline__ double F(int s, int n) const { return sF[n * NCELL + o[s]]; }
	__device__ __forceinline__ double lam(int s, int t) const { return sL[t * NCELL + o[s]]; }
};

template <class C, int WENO>
struct XfMarchGeom
{
	static constexpr int E = C::E, NC = C::NC, NST = XfStencil<WENO>::NST, P = XfStencil<WENO>::P;
	static constexpr int W = XF_MW, TF = XF_MTF(WENO), R = TF + NST - 1, NCELL = R * W, NRAW = E + 5 + NC, NT = W * TF;
	static constexpr size_t ring_doubles = size_t(2 * E + 3) * NCELL, land_doubles = size_t(NRAW) * TF * W;
	// exchange rows: [0] = the last face of the previous iteration (always its own storage), [1 .. TF] = this iteration's faces
	static constexpr size_t carry_doubles = size_t(E) * W, exch_doubles = XF_MARCH_ALIAS ? 0 : size_t(TF) * E * W;
	static_assert(!XF_MARCH_ALIAS || size_t(TF) * E * W <= land_doubles, "exchange rows must fit the landing buffer");
	static constexpr size_t smem_bytes = (ring_doubles + land_doubles + carry_doubles + exch_doubles) * sizeof(double) + 16;
};

template <class C, int DIR, int WENO, bool PP>
__global__ void __launch_bounds__(XF_MW *XF_MTF(WENO), XF_MARCH_MINB)
	k_march(XfDev d, const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmY,
			const double *__restrict__ UI, XfMarchArgs a)
{
	using G = XfMarchGeom<C, WENO>;
	constexpr int E = G::E, NC = G::NC, NST = G::NST, P = G::P, W = G::W, TF = G::TF, R = G::R, NCELL = G::NCELL;
	extern __shared__ __align__(128) double smem[];
	double *sU = smem, *sF = smem + E * NCELL, *sL = smem + 2 * E * NCELL;
	double *land = smem + G::ring_doubles;                 // [NRAW][TF][W]: U[E], u v w p c, y[NC]
	double *carry = land + G::land_doubles;                // [E][W]: the last face of the previous iteration
	double *exch = XF_MARCH_ALIAS ? land : carry + G::carry_doubles; // [TF][E][W]: this iteration's faces
	unsigned long long *bar = reinterpret_cast<unsigned long long *>(carry + G::carry_doubles + G::exch_doubles);

	const int tid = threadIdx.x, tx = tid % W, ty = tid / W;
	const int c0 = d.Bx + blockIdx.x * W, i = c0 + tx;
	const bool iok = i < d.Bx + d.Xi;
	const int tr = a.t0 + blockIdx.y; // DIR 1: z-plane k; DIR 2: row j
	// this block's segment of the march
	const int ca = a.ca + blockIdx.z * a.seglen;
	const int cb = min(a.cb, ca + a.seglen);
	if (ca >= cb)
		return;
	const int a0 = ca - 1 - P;         // row of stencil slot 0 of the first face
	const int nf = cb - ca + 1;        // faces ca - 1 .. cb - 1
	const int nit = (nf + TF - 1) / TF;
	const long long sS = DIR == 1 ? d.sY : d.sZ;
	const long long col = DIR == 1 ? (long long)tr * d.sZ + i : (long long)tr * d.sY + i; // cell index = col + q * sS

	if (tid == 0)
		xf_mbar_init(bar, 1);
	__syncthreads();

	auto issue = [&](int b)
	{ // batch b = rows a0 + NST - 1 + b TF ... + TF - 1 (TMA zero-fills whatever lies outside the arrays)
		if (tid == 0)
		{
			const int row0 = a0 + NST - 1 + b * TF;
			const int cy = DIR == 1 ? row0 : tr, cz = DIR == 1 ? tr : row0;
			xf_mbar_expect_tx(bar, unsigned(G::land_doubles * sizeof(double)));
			xf_tma_load_4d(land, &tmU, c0, cy, cz, 0, bar);
			xf_tma_load_4d(land + E * TF * W, &tmP, c0, cy, cz, 0, bar);
			if constexpr (NC > 0)
				xf_tma_load_4d(land + (E + 5) * TF * W, &tmY, c0, cy, cz, 0, bar);
		}
	};
	// raw row ty of the landed batch -> ring slot `slot` (GetPhysFlux, Update_device.hpp:81-110; GetLocalEigen): as stage_cell of k_sweep
	auto convert = [&](int slot)
	{
		const int l = ty * W + tx, c = slot * W + tx;
		double Uc[E];
#pragma unroll
		for (int n = 0; n < E; n++)
			Uc[n] = land[n * (TF * W) + l];
		const double u = land[(E + 0) * (TF * W) + l], v = land[(E + 1) * (TF * W) + l], w = land[(E + 2) * (TF * W) + l];
		const double p = land[(E + 3) * (TF * W) + l], cc = land[(E + 4) * (TF * W) + l];
		const double un = DIR == 1 ? v : w;
		const double m = Uc[1 + DIR];
		double Fc[E];
		Fc[0] = m;
		Fc[1] = m * u;
		Fc[2] = DIR == 1 ? m * v + p : m * v;
		Fc[3] = DIR == 2 ? m * w + p : m * w;
		Fc[4] = (Uc[4] + p) * un;
#pragma unroll
		for (int s = 0; s < NC; s++)
			Fc[5 + s] = m * land[(E + 5 + s) * (TF * W) + l];
#pragma unroll
		for (int n = 0; n < E; n++)
			sU[n * NCELL + c] = Uc[n], sF[n * NCELL + c] = 0.5 * Fc[n];
		sL[c] = fabs(un - cc), sL[NCELL + c] = fabs(un), sL[2 * NCELL + c] = fabs(un + cc);
	};

	// ---- prologue: the NST - 1 rows below the first batch, then batch 0 ----
	unsigned phase = 0;
	issue(-1);
	xf_mbar_wait(bar, phase), phase ^= 1;
	if (ty >= TF - (NST - 1))
		convert(ty - (TF - (NST - 1)));
	__syncthreads();
	issue(0);
	xf_mbar_wait(bar, phase), phase ^= 1;
	convert(NST - 1 + ty);
	__syncthreads();
	if (1 < nit)
		issue(1);

	const double dl = DIR == 1 ? d.dy : d.dz, _dl = DIR == 1 ? d._dy : d._dz;
	const double dt = (a.mode == XF_MODE_RK && a.dt_dev) ? *a.dt_dev : 0.0;
	// Global-load latency is kept off the critical path without spending shared memory on it: every warp issues the loads of a phase
	// as soon as IT is done with the previous one, ahead of the block-wide barrier -- the face-side scalars of iteration it + 1 after
	// its update of iteration it (before the second barrier), the update's operands right after its own flux (before the first).
	// Warps reach the barriers at different times, so one warp's load latency overlaps the others' arithmetic.  (First version: both
	// groups of loads were issued after the barriers, by all warps at once: long_scoreboard 4.8 warps per issue, 57 ms per z sweep.)
	XfSide<C> sl, sr;
	auto load_sides = [&](long long id)
	{ // rho comes from the ring (the staged copy of UI: no read of a field that is being updated in place)
		side_scalars<C>(d, id, sl), side_scalars<C>(d, id + sS, sr);
	};
	if (XF_MARCH_SIDEPF && iok && ca - 1 + ty <= cb - 1)
		load_sides(col + (long long)(ca - 1 + ty) * sS);
	int base = 0; // (it * TF) mod R
#pragma unroll 1
	for (int it = 0; it < nit; it++)
	{
		const int f = ca - 1 + it * TF + ty; // this thread's face: between cells f and f + 1
		const bool valid = iok && f <= cb - 1;
		const long long id_l = col + (long long)f * sS;
		const bool upd = valid && f >= ca && a.mode != XF_MODE_FW;
		double F[E], own[E], lu_in[E], oth[E];
		if (valid)
		{
			RingStencil<NST, NCELL> st;
			st.sU = sU, st.sF = sF, st.sL = sL;
#pragma unroll
			for (int s = 0; s < NST; s++)
			{
				int r = base + ty + s;
				r = r >= R ? r - R : r;
				st.o[s] = r * W + tx;
			}
			XfRoe<C> Rs;
			if (!XF_MARCH_SIDEPF)
				load_sides(id_l);
			sl.rho = st.U(P, 0), sr.rho = st.U(P + 1, 0);
			xf_roe_state<C>(sl, sr, d.gamma0, Rs);
			double glf[3] = {d.red[XF_RED_GLF + DIR * 3 + 0], d.red[XF_RED_GLF + DIR * 3 + 1], d.red[XF_RED_GLF + DIR * 3 + 2]};
			xf_face_flux<C, DIR, WENO>(st, Rs, d.alpha, glf, dl, F);
			if constexpr (PP)
			{ // the face below the first inner cell is never limited (ConVenction_block.hpp:330-410)
				if (f >= (DIR == 1 ? d.By : d.Bz))
					xf_positivity<C, WENO>(st, d.red[XF_RED_PPL + DIR], d.CFL, F);
			}
			if (a.mode == XF_MODE_FW)
			{
#pragma unroll
				for (int n = 0; n < E; n++)
					a.Fw[n * d.N + id_l] = F[n];
			}
			else
			{
				if (!XF_MARCH_ALIAS)
				{
#pragma unroll
					for (int n = 0; n < E; n++)
						exch[(ty * E + n) * W + tx] = F[n];
				}
				if (upd)
				{ // operands of this cell's update: the sweep input's own value from the ring (its row may retire at the barrier), the rest from HBM
#pragma unroll
					for (int n = 0; n < E; n++)
					{
						const long long o = n * d.N + id_l;
						own[n] = st.U(P, n);
						lu_in[n] = a.first ? 0.0 : a.LU[o];
						if (a.mode == XF_MODE_RK)
							oth[n] = (a.flag == 1) ? (a.guard ? a.U1[o] : 0.0) : a.U[o]; // stage 1 reads U1 for the guard only (EstimateFluidNAN checks U1, U1, U)
					}
				}
			}
		}
		__syncthreads(); // every face of this iteration is done with the ring; the fluxes are in exch
		if (it + 1 < nit)
		{ // batch it + 1 replaces the TF rows that have just retired
			xf_mbar_wait(bar, phase), phase ^= 1;
			int r = base + ty; // (NST - 1 + (it + 1) TF + ty) mod R == (it TF + ty) mod R: the slot of the row that has just retired
			r = r >= R ? r - R : r;
			convert(r);
		}
		if (XF_MARCH_ALIAS && a.mode != XF_MODE_FW)
		{ // the landing buffer has been consumed: it now carries the fluxes from thread row ty to thread row ty + 1
			__syncthreads();
			if (valid)
			{
#pragma unroll
				for (int n = 0; n < E; n++)
					exch[(ty * E + n) * W + tx] = F[n];
			}
			__syncthreads();
		}
		if (upd)
		{
			bool bad = false;
#pragma unroll
			for (int n = 0; n < E; n++)
			{
				const long long o = n * d.N + id_l;
				const double Fp = ty == 0 ? carry[n * W + tx] : exch[((ty - 1) * E + n) * W + tx];
				const double lu = lu_in[n] + (Fp - F[n]) * _dl;
				if (a.mode == XF_MODE_ACC)
					a.LU[o] = lu;
				else
				{
					const double u0 = (a.flag == 1) ? own[n] : oth[n], u1 = (a.flag == 1) ? oth[n] : own[n];
					const double v1 = (a.flag == 1) ? 0.0 : u1;
					if (a.guard)
					{
						const double ui = (a.flag == 3) ? u0 : u1;
						bad = bad || isnan(ui) || isinf(ui) || isnan(lu) || isinf(lu) || (n == 0 && ui < 0);
					}
					const double r = rk_of(u0, v1, lu, dt, a.flag);
					if (a.flag == 3)
						a.U[o] = r;
					else
						a.U1[o] = r;
				}
			}
			if (bad)
				d.err[2] = 1;
		}
		if (XF_MARCH_SIDEPF && it + 1 < nit && iok && f + TF <= cb - 1)
			load_sides(id_l + (long long)TF * sS); // face-side scalars of this thread's next face
		__syncthreads(); // ring holds the next iteration's stencil; landing buffer and exch rows 1.. are free
		if (it + 2 < nit)
			issue(it + 2);
		if (ty == TF - 1 && valid && a.mode != XF_MODE_FW)
		{
#pragma unroll
			for (int n = 0; n < E; n++)
				carry[n * W + tx] = F[n];
		}
		base += TF;
		base = base >= R ? base - R : base;
	}
}
