// xf_visc.cuh -- viscous, heat-conduction and species-diffusion wall fluxes.  Included inside namespace XF_NS.
//
// Reference: GetCellCenterDerivative + CenterDerivativeBCKernel{X,Y,Z} (viscosity/Visc_block.hpp:8-84, Visc_kernels.hpp:7-217),
// GetInnerCellCenterDerivativeKernel / GetWallViscousFlux{X,Y,Z} (viscosity/Fourth_Order/Visc_Order_kernels.hpp:16-340 with the macros of
// Fourth_Order/Flux_discrete.h), Gettransport_coeff_aver / Get_transport_coeff_aver (Visc_kernels.hpp:219-241, Visc_device.h:10-177) and
// the viscous block of GetLU that sequences them (FDM_Method/ConVenction_block.hpp:424-575).  Every expression keeps the reference's
// association order; log() is the bit-exact xf_log; exp() and pow() are CUDA's (<= 1-2 ulp from glibc's: the viscous terms carry the
// north_star tolerance, not a bitwise claim).
//
// Pure per-cell functions of (T, p, X) that the reference re-evaluates many times per cell -- ln T, the NS species viscosities, the
// NS x NS molar-mass factors of PHI -- are evaluated once; the calls themselves are the same, so are the results.
//
// Quirks kept: the species-diffusion limiter Diffu_limiter = Dim_max * Dim_limiter * yi_max with Dim_max left at 0.0 by the reference's
// single-process build (its reduction over Dkm is commented out, ConVenction_block.hpp:478-479; the MPI build sets 1.0,:497): with
// XfVisc::dim_max0 = 0 the diffusion fluxes are clipped to zero exactly as the reference's are; F_wall_v[p] of the species equations uses
// Dim_Yil of the LAST species of the loop for every p (Flux_discrete.h:69).  Not reproduced: the out-of-range reads of the derivative
// kernel in inactive dimensions (multiplied by DimX_t = 0 there; taken as +0 here) and CenterDerivativeBCKernelZ's nslipWall branch,
// whose mirror index (k - offset) is negative (Visc_kernels.hpp:197).
#pragma once

// GetInnerCellCenterDerivativeKernel: cells [B-2, B+inner+2) of every active direction
__global__ void __launch_bounds__(256) k_vde(XfDev d, XfVisc vs)
{
	const int nx = d.DimX ? d.Xi + 4 : 1, ny = d.DimY ? d.Yi + 4 : 1, nz = d.DimZ ? d.Zi + 4 : 1;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int ii = int(t % nx);
	const long long r = t / nx;
	const int jj = int(r % ny), kk = int(r / ny);
	if (kk >= nz)
		return;
	const int i = ii + (d.DimX ? d.Bx - 2 : 0), j = jj + (d.DimY ? d.By - 2 : 0), k = kk + (d.DimZ ? d.Bz - 2 : 0);
	const long long id = ((long long)k * d.Ymax + j) * d.Xp + i;
	const double _twle = 1.0 / 12.0;
	const double tx = d.DimX ? 1.0 : 0.0, ty = d.DimY ? 1.0 : 0.0, tz = d.DimZ ? 1.0 : 0.0;
	auto diff = [&](const double *q, long long s, double _dl) { return (8.0 * (q[id + s] - q[id - s]) - (q[id + 2 * s] - q[id - 2 * s])) * _dl * _twle; };
	double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
	if (d.DimX)
		D[0] = diff(d.u, 1, d._dx) * tx, D[1] = diff(d.v, 1, d._dx) * tx * ty, D[2] = diff(d.w, 1, d._dx) * tx * tz;
	if (d.DimY)
		D[3] = diff(d.u, d.sY, d._dy) * ty * tx, D[4] = diff(d.v, d.sY, d._dy) * ty, D[5] = diff(d.w, d.sY, d._dy) * ty * tz;
	if (d.DimZ)
		D[6] = diff(d.u, d.sZ, d._dz) * tz * tx, D[7] = diff(d.v, d.sZ, d._dz) * tz * ty, D[8] = diff(d.w, d.sZ, d._dz) * tz;
#pragma unroll
	for (int m = 0; m < 9; m++)
		vs.Vde[m * d.N + id] = D[m];
}

// CenterDerivativeBCKernel{X,Y,Z}: the four derivatives a direction's wall flux reads, in the ghost cells of that direction
template <int DIR>
__device__ __forceinline__ void vde_bc_apply(const XfDev &d, const XfVisc &vs, int BC, int i, int j, int k, int mirror_offset, int index_inner, int sign)
{
	// X: ducy ducz dvcy dwcz ; Y: dvcx dvcz ducx dwcz ; Z: dwcx dwcy ducx dvcy
	constexpr int sel[3][4] = {{3, 6, 4, 8}, {1, 7, 0, 8}, {2, 5, 0, 4}};
	const int Bw = DIR == 0 ? d.Bx : (DIR == 1 ? d.By : d.Bz);
	const int inner = DIR == 0 ? d.Xi : (DIR == 1 ? d.Yi : d.Zi);
	const int g = DIR == 0 ? i : (DIR == 1 ? j : k);
	const long long id = ((long long)k * d.Ymax + j) * d.Xp + i;
	auto tid = [&](int t) -> long long
	{ return DIR == 0 ? ((long long)k * d.Ymax + j) * d.Xp + t : (DIR == 1 ? ((long long)k * d.Ymax + t) * d.Xp + i : ((long long)t * d.Ymax + j) * d.Xp + i); };
	long long src = -1;
	double sg = 1.0;
	switch (BC)
	{
	case 2: src = tid(2 * (Bw + mirror_offset) - 1 - g); break;          // Symmetry
	case 3: src = tid(g + sign * inner); break;                           // Periodic
	case 1: src = tid(index_inner); break;                                // Outflow
	case 4:                                                               // nslipWall
		if (DIR == 2)
			return; // the reference's mirror index is out of range here (see the header)
		src = tid(2 * (Bw + mirror_offset) - 1 - g), sg = -1.0;
		break;
	case 0:                                                               // Inflow: zero
#pragma unroll
		for (int n = 0; n < 4; n++)
			vs.Vde[sel[DIR][n] * d.N + id] = 0.0;
		return;
	default: return;                                                      // viscWall, slipWall, innerBlock, BC_COPY: untouched
	}
#pragma unroll
	for (int n = 0; n < 4; n++)
	{
		const double v = vs.Vde[sel[DIR][n] * d.N + src];
		vs.Vde[sel[DIR][n] * d.N + id] = sg < 0 ? -v : v;
	}
}
template <int DIR>
__global__ void __launch_bounds__(256) k_vde_bc(XfDev d, XfVisc vs, int bc_min, int bc_max)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	int i, j, k, g;
	if constexpr (DIR == 0)
	{
		g = int(t % d.Bx);
		const long long r = t / d.Bx;
		j = int(r % d.Ymax), k = int(r / d.Ymax);
		if (k >= d.Zmax)
			return;
		vde_bc_apply<0>(d, vs, bc_min, g, j, k, 0, d.Bx, 1);
		vde_bc_apply<0>(d, vs, bc_max, g + d.Xmax - d.Bx, j, k, d.Xi, d.Xmax - d.Bx - 1, -1);
	}
	else if constexpr (DIR == 1)
	{
		i = int(t % d.Xmax);
		const long long r = t / d.Xmax;
		g = int(r % d.By), k = int(r / d.By);
		if (k >= d.Zmax)
			return;
		vde_bc_apply<1>(d, vs, bc_min, i, g, k, 0, d.By, 1);
		vde_bc_apply<1>(d, vs, bc_max, i, g + d.Ymax - d.By, k, d.Yi, d.Ymax - d.By - 1, -1);
	}
	else
	{
		i = int(t % d.Xmax);
		const long long r = t / d.Xmax;
		j = int(r % d.Ymax), g = int(r / d.Ymax);
		if (g >= d.Bz)
			return;
		vde_bc_apply<2>(d, vs, bc_min, i, j, g, 0, d.Bz, 1);
		vde_bc_apply<2>(d, vs, bc_max, i, j, g + d.Zmax - d.Bz, d.Zi, d.Zmax - d.Bz - 1, -1);
	}
}

// ln(coefficient) = ((c3 L + c2) L + c1) L + c0 with L = ln T  (Viscosity / Thermal_conductivity / GetDkj, Visc_device.h:10-101)
// Out of line (thirty inlined copies of exp() made k_transport stall on instruction fetch) and two arguments per call: exp and pow are
// single dependency chains of ~15 / ~40 FP64 instructions, two independent ones interleave in the FP64 pipe (the kernel's top stall was
// `wait`).  glibc's exp / pow, bit for bit (xf_exp.cuh).
static __device__ __noinline__ double2 xf_exp_pair(double a, double b) { return make_double2(xf_exp(a), xf_exp(b)); }
static __device__ __noinline__ double2 xf_pow_half_pair(double a, double b) { return make_double2(xf_pow(a, 0.5), xf_pow(b, 0.5)); }
#define xf_fit_arg(c, L) ((((c)[3] * (L) + (c)[2]) * (L) + (c)[1]) * (L) + (c)[0]) // by value: the coefficients live in the kernel-parameter bank

// Gettransport_coeff_aver over ALL cells (Visc_kernels.hpp:219-241)
template <class C>
__global__ void __launch_bounds__(128, 5) k_transport(XfDev d, XfThermo th, XfVisc vs, const double *__restrict__ U, int k0, int k1)
{
	constexpr int NS = C::NS;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int i = int(t % d.Xp);
	const long long row = t / d.Xp;
	const int j = int(row % d.Ymax), k = k0 + int(row / d.Ymax);
	if (k >= k1 || i >= d.Xmax)
		return;
	const long long id = ((long long)k * d.Ymax + j) * d.Xp + i;
	const double T = d.T[id], p = d.p[id], rho = U[id];
	double yi[NS], X[NS];
	if constexpr (C::COP)
	{
#pragma unroll
		for (int n = 0; n < NS; n++)
			yi[n] = d.y[n * d.N + id];
	}
	else
		yi[0] = 1.0;
	if (vs.diffu)
	{ // hi = get_Enthalpy(T) per species (Visc_kernels.hpp:232-233)
		double hi[NS];
		if constexpr (C::COP)
			xf_species_h<C>(th, T, hi);
		else
			hi[0] = 0.0;
#pragma unroll
		for (int n = 0; n < NS; n++)
			vs.hi[n * d.N + id] = hi[n];
	}
	// get_xi (Mixing_device.h:31-44): mole fractions, total concentration
	double C_total = 0.0;
#pragma unroll
	for (int n = 0; n < NS; n++)
	{
		X[n] = yi[n] * th._Wi[n] * 1e-3 * rho;
		C_total = C_total + X[n];
	}
	const double _C_total = 1.0 / C_total;
#pragma unroll
	for (int n = 0; n < NS; n++)
		X[n] = X[n] * _C_total;
	// Get_transport_coeff_aver (Visc_device.h:107-177).  The reference evaluates Viscosity(k), Viscosity(i) inside the pair loop and three
	// pow() per pair; here every fit is evaluated once per cell, the two molar-mass powers are the host's (xf_set_transport), and log, exp
	// and pow replay glibc's algorithms (xf_log.cuh, xf_exp.cuh): the same bits as the reference CPU path.
	const double L = xf_log(T);
	double mu[NS], lam[NS];
#pragma unroll
	for (int n = 0; n < NS; n++)
	{
		const double2 e = xf_exp_pair(xf_fit_arg(vs.fit_visc[n], L), vs.heat ? xf_fit_arg(vs.fit_therm[n], L) : 0.0);
		mu[n] = e.x, lam[n] = vs.heat ? e.y : 0.0;
	}
	// quotients by one denominator share its IEEE reciprocal (xf_div_shared, xf_math.cuh: the correctly rounded quotient in 3 instructions)
	const double sqrt2 = sqrt(2.0), _sqrt2 = 1.0 / sqrt2, _p = 1.0 / p;
	// den_k = sum_i X_i PHI(k, i), i ascending (Visc_device.h:34-41, 120-140).  The two PHI of an unordered pair are evaluated together; the
	// loop order below still adds the terms of every den_k in ascending i: (j, k) pairs with j < k arrive at outer index j, the diagonal
	// term at the start of outer index k, the (k, i > k) terms after it.
	double den[NS], _mu[NS];
#pragma unroll
	for (int n = 0; n < NS; n++)
		den[n] = 0.0, _mu[n] = 1.0 / mu[n]; // every mu_n is the denominator of NS - 1 ratios
	auto phi_of = [&](int kk, int ii, double root) { // root = pow(mu_k / mu_i, 0.5)
		double phi = vs.phiW[kk * NS + ii] * root;
		phi = xf_div_shared((phi + 1.0) * (phi + 1.0) * 0.5, sqrt2, _sqrt2);
		return phi * vs.phiS[kk * NS + ii];
	};
#pragma unroll
	for (int a = 0; a < NS; a++)
	{
		den[a] = den[a] + X[a] * phi_of(a, a, 1.0); // pow(1, 0.5) == 1
#pragma unroll
		for (int b = a + 1; b < NS; b++)
		{
			const double2 r = xf_pow_half_pair(xf_div_shared(mu[a], mu[b], _mu[b]), xf_div_shared(mu[b], mu[a], _mu[a]));
			den[a] = den[a] + X[b] * phi_of(a, b, r.x);
			den[b] = den[b] + X[a] * phi_of(b, a, r.y);
		}
	}
	double va = 0.0, tca = 0.0;
#pragma unroll
	for (int kk = 0; kk < NS; kk++)
	{
		const double _den = 1.0 / den[kk];
		va = va + X[kk] * mu[kk] * _den;
		if (vs.heat)
			tca = tca + X[kk] * lam[kk] * _den;
	}
	vs.va[id] = va;
	if (vs.heat)
		vs.tca[id] = tca;
	if (vs.diffu)
	{
		double Dk[NS];
		if constexpr (NS > 1)
		{
			// (X_i + 1e-40) / (D_ik / p + 1e-40) per ordered pair; with bitwise symmetric fits (GetFitCoefficient's are) D_ik is D_ki and the
			// NS (NS - 1) / 2 coefficients of the upper triangle are evaluated two per call
			constexpr int NP = NS * (NS - 1) / 2;
			double Dp[NS][NS], _Dp[NS][NS]; // D_ik / p + 1e-40 and its reciprocal (symmetric fits: one reciprocal serves both quotients of a pair)
			for (int pass = 0; pass < (vs.dkj_sym ? 1 : 2); pass++)
			{
				double arg[NP + 1];
				{
					int q = 0;
#pragma unroll
					for (int kk = 0; kk < NS; kk++)
#pragma unroll
						for (int ii = kk + 1; ii < NS; ii++, q++)
							arg[q] = pass == 0 ? xf_fit_arg(vs.fit_Dkj[ii * NS + kk], L) : xf_fit_arg(vs.fit_Dkj[kk * NS + ii], L);
					arg[NP] = 0.0;
				}
#pragma unroll
				for (int q = 0; q < NP; q += 2)
				{
					const double2 e = xf_exp_pair(arg[q], arg[q + 1]);
					arg[q] = e.x, arg[q + 1] = e.y;
				}
				{
					int q = 0;
#pragma unroll
					for (int kk = 0; kk < NS; kk++)
#pragma unroll
						for (int ii = kk + 1; ii < NS; ii++, q++)
						{
							const double v = xf_div_shared(arg[q], p, _p) + 1.0e-40;
							const double _v = 1.0 / v;
							if (pass == 0)
								Dp[kk][ii] = v, Dp[ii][kk] = v, _Dp[kk][ii] = _v, _Dp[ii][kk] = _v; // Dp[kk][ii] uses fit_Dkj[ii * NS + kk]; symmetric: the same for [ii][kk]
							else
								Dp[ii][kk] = v, _Dp[ii][kk] = _v; // not symmetric: Dp[ii][kk] uses fit_Dkj[kk * NS + ii]
						}
				}
			}
#pragma unroll
			for (int kk = 0; kk < NS; kk++)
			{
				double temp1 = 0.0, temp2 = 1.0e-20;
#pragma unroll
				for (int ii = 0; ii < NS; ii++)
					if (ii != kk)
					{
						temp1 += (X[ii] + 1.0e-40) * vs.Wi[ii];
						temp2 += xf_div_shared(X[ii] + 1.0e-40, Dp[kk][ii], _Dp[kk][ii]);
					}
				// sycl::step(ceil(temp1), 0.0) == 1 <=> 0.0 >= ceil(temp1)
				if (!(0.0 < ceil(temp1)))
					Dk[kk] = xf_exp(xf_fit_arg(vs.fit_Dkj[kk * NS + kk], L)) / p;
				else
					Dk[kk] = temp1 / temp2 / rho * C_total;
				Dk[kk] *= 1.0e-1;
			}
		}
		else
		{
			Dk[0] = xf_exp(xf_fit_arg(vs.fit_Dkj[0], L)) / p;
			Dk[0] *= 1.0e-1;
		}
#pragma unroll
		for (int kk = 0; kk < NS; kk++)
			vs.Dkm[kk * d.N + id] = xf_max(Dk[kk], 1.0e-10);
	}
}

// min / max of every mass fraction over the inner cells, both reductions starting from 0.0 (ConVenction_block.hpp:460-487)
__device__ __forceinline__ void atomic_min_max(double *pmin, double *pmax, double lo, double hi)
{
	unsigned long long *a = reinterpret_cast<unsigned long long *>(pmin);
	unsigned long long old = *a, assumed;
	while (lo < __longlong_as_double((long long)old))
	{
		assumed = old;
		old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(lo));
		if (old == assumed)
			break;
	}
	a = reinterpret_cast<unsigned long long *>(pmax);
	old = *a;
	while (__longlong_as_double((long long)old) < hi)
	{
		assumed = old;
		old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(hi));
		if (old == assumed)
			break;
	}
}
template <int NS>
__global__ void __launch_bounds__(256) k_yi_minmax(XfDev d, XfVisc vs)
{
	long long id;
	const bool in = inner_cell(d, id);
	double lo[NS], hi[NS];
#pragma unroll
	for (int n = 0; n < NS; n++)
	{
		const double y = in ? d.y[n * d.N + id] : 0.0;
		lo[n] = y, hi[n] = y;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			lo[n] = fmin(lo[n], __shfl_xor_sync(0xffffffffu, lo[n], o));
			hi[n] = fmax(hi[n], __shfl_xor_sync(0xffffffffu, hi[n], o));
		}
	}
	if ((threadIdx.x & 31) == 0)
#pragma unroll
		for (int n = 0; n < NS; n++)
			atomic_min_max(vs.lim + n, vs.lim + NS + n, lo[n], hi[n]);
}
// yi_max -= yi_min; yi_max *= Yil_limiter; Dim_max *= Dim_limiter * yi_max  (ConVenction_block.hpp:489-503)
__global__ void k_visc_limits(XfVisc vs, int NS)
{
	const int n = threadIdx.x;
	if (n >= NS)
		return;
	double ymin = vs.lim[n], ymax = vs.lim[NS + n], dmax = vs.dim_max0;
	if (vs.dim_max0 != 0.0)
		ymin = xf_max(ymin, 0.0), ymax = xf_min(ymax, 1.0); // the MPI build clamps the reduced extrema (:495)
	ymax -= ymin;
	ymax *= vs.Yil_limiter;
	dmax *= vs.Dim_limiter * ymax;
	vs.lim[2 * NS + n] = ymax, vs.lim[3 * NS + n] = dmax;
}

// GetWallViscousFlux{X,Y,Z}: Flux_wall -= F_wall_v at the face of direction DIR above cell id (the flux itself: xf_visc_face.cuh)
template <class C, int DIR>
__device__ __forceinline__ void visc_face(const XfDev &d, const XfVisc &vs, const double *__restrict__ U, double *__restrict__ Fw, const long long id)
{
	double Fv[C::E];
	visc_face_flux<C, DIR>(d, vs, U, id, Fv);
#pragma unroll
	for (int n = 0; n < C::E; n++)
		Fw[n * d.N + id] -= Fv[n];
}

// All three directions in one pass: one thread per cell of the box [B - 1, B + inner) of every active direction forms the viscous flux at
// its upper x, y and z face (where that face exists: the other two indices inner).  The ~26 per-cell arrays the three stencils share
// come from DRAM once instead of three times; blocks walk the (y, z) plane in tiles, z fastest (k_rk's order), so that the rows / planes a
// stencil shares with its neighbours are in L1 / L2.
template <class C>
__global__ void __launch_bounds__(128) k_visc_flux3(XfDev d, XfVisc vs, const double *__restrict__ U)
{
	constexpr int TT = XF_RK_TILE > 0 ? XF_RK_TILE : 8;
	const int ny = d.Yi + (d.DimY ? 1 : 0), nz = d.Zi + (d.DimZ ? 1 : 0);
	const int nty = (ny + TT - 1) / TT, ntz = (nz + TT - 1) / TT;
	const unsigned per_x = unsigned(nty) * unsigned(ntz) * (TT * TT);
	const int xc = int(blockIdx.x / per_x);
	const unsigned r = unsigned(blockIdx.x - (long long)xc * per_x);
	const unsigned tile = r / (TT * TT), w = r % (TT * TT);
	const int kk = int(tile / nty) * TT + int(w % TT), jj = int(tile % nty) * TT + int(w / TT);
	const int ii = xc * 128 + int(threadIdx.x);
	if (jj >= ny || kk >= nz || ii >= d.Xi + (d.DimX ? 1 : 0))
		return;
	const int i = ii + d.Bx - (d.DimX ? 1 : 0), j = jj + d.By - (d.DimY ? 1 : 0), k = kk + d.Bz - (d.DimZ ? 1 : 0);
	const long long id = ((long long)k * d.Ymax + j) * d.Xp + i;
	const bool xin = i >= d.Bx, yin = j >= d.By, zin = k >= d.Bz;
	if (d.DimX && yin && zin)
		visc_face<C, 0>(d, vs, U, d.Fw[0], id);
	if (d.DimY && xin && zin)
		visc_face<C, 1>(d, vs, U, d.Fw[1], id);
	if (d.DimZ && xin && yin)
		visc_face<C, 2>(d, vs, U, d.Fw[2], id);
}

// the viscous block of GetLU (ConVenction_block.hpp:424-575) after the inviscid wall fluxes (and their limiter) are in Fw; U = the sweep input
template <class C>
// parts: 1 = everything the wall fluxes read (derivatives + their ghost fill, transport coefficients, limiter extrema), 2 = the wall fluxes
// (stand-alone kernel; the sweeps' tails do the same on the fly), 3 = both
static int visc_t(const XfDev &d, const XfThermo &th, const XfVisc &vs, const double *U, const int bc[6], cudaStream_t s, long long *launches, int parts)
{
	if (parts & 1)
	{
	const long long nv = (long long)(d.DimX ? d.Xi + 4 : 1) * (d.DimY ? d.Yi + 4 : 1) * (d.DimZ ? d.Zi + 4 : 1);
	k_vde<<<(unsigned)((nv + 255) / 256), 256, 0, s>>>(d, vs);
	++*launches;
	if (d.DimX)
		k_vde_bc<0><<<(unsigned)(((long long)d.Bx * d.Ymax * d.Zmax + 255) / 256), 256, 0, s>>>(d, vs, bc[0], bc[1]), ++*launches;
	if (d.DimY)
		k_vde_bc<1><<<(unsigned)(((long long)d.Xmax * d.By * d.Zmax + 255) / 256), 256, 0, s>>>(d, vs, bc[2], bc[3]), ++*launches;
	if (d.DimZ)
		k_vde_bc<2><<<(unsigned)(((long long)d.Xmax * d.Ymax * d.Bz + 255) / 256), 256, 0, s>>>(d, vs, bc[4], bc[5]), ++*launches;
	const long long nall = (long long)d.sZ * d.Zmax;
	k_transport<C><<<(unsigned)((nall + 127) / 128), 128, 0, s>>>(d, th, vs, U, 0, d.Zmax);
	++*launches;
	if (vs.diffu && C::COP)
	{
		cudaMemsetAsync(vs.lim, 0, 4 * C::NS * sizeof(double), s);
		const long long ni = (long long)d.Xi * d.Yi * d.Zi;
		k_yi_minmax<C::NS><<<(unsigned)((ni + 255) / 256), 256, 0, s>>>(d, vs);
		k_visc_limits<<<1, 32, 0, s>>>(vs, C::NS);
		*launches += 2;
	}
	}
	if (parts & 2)
	{
		constexpr int TT = XF_RK_TILE > 0 ? XF_RK_TILE : 8;
		const int nx = d.Xi + (d.DimX ? 1 : 0), ny = d.Yi + (d.DimY ? 1 : 0), nz = d.Zi + (d.DimZ ? 1 : 0);
		const long long nb = (long long)((nx + 127) / 128) * ((ny + TT - 1) / TT) * ((nz + TT - 1) / TT) * (TT * TT);
		k_visc_flux3<C><<<(unsigned)nb, 128, 0, s>>>(d, vs, U);
		++*launches;
	}
	XF_CHECK_LAUNCH();
	return 0;
}
