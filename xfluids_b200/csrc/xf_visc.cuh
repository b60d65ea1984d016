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
static __device__ __noinline__ double xf_pow_half(double x) { return xf_pow(x, 0.5); }
// (out of line: thirty inlined copies of exp() made k_transport stall on instruction fetch)
static __device__ __noinline__ double xf_fit4(double c0, double c1, double c2, double c3, double L) { return xf_exp(((c3 * L + c2) * L + c1) * L + c0); } // glibc's exp, bit for bit (xf_exp.cuh)
#define xf_fit(c, L) xf_fit4((c)[0], (c)[1], (c)[2], (c)[3], (L)) // by value: the coefficients live in the kernel-parameter bank

// Gettransport_coeff_aver over ALL cells (Visc_kernels.hpp:219-241)
template <class C>
__global__ void __launch_bounds__(128) k_transport(XfDev d, XfThermo th, XfVisc vs, const double *__restrict__ U, int k0, int k1)
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
		mu[n] = xf_fit(vs.fit_visc[n], L), lam[n] = vs.heat ? xf_fit(vs.fit_therm[n], L) : 0.0;
	double va = 0.0, tca = 0.0;
	const double sqrt2 = sqrt(2.0);
#pragma unroll
	for (int kk = 0; kk < NS; kk++)
	{
		double den = 0.0;
#pragma unroll
		for (int ii = 0; ii < NS; ii++)
		{ // PHI(specie_k, specie_i) (Visc_device.h:34-41)
			double phi = vs.phiW[kk * NS + ii] * (ii == kk ? 1.0 : xf_pow_half(mu[kk] / mu[ii])); // pow(1, 0.5) == 1
			phi = (phi + 1.0) * (phi + 1.0) * 0.5 / sqrt2;
			phi = phi * vs.phiS[kk * NS + ii];
			den = den + X[ii] * phi;
		}
		const double _den = 1.0 / den;
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
			// (X_i + 1e-40) / (D_ik / p + 1e-40) per ordered pair; with bitwise symmetric fits (GetFitCoefficient's are) D_ik is D_ki
			double Dp[NS][NS];
#pragma unroll
			for (int kk = 0; kk < NS; kk++)
#pragma unroll
				for (int ii = 0; ii < NS; ii++)
					if (ii != kk)
					{
						if (ii > kk || !vs.dkj_sym)
							Dp[kk][ii] = xf_fit(vs.fit_Dkj[ii * NS + kk], L) / p + 1.0e-40;
						else
							Dp[kk][ii] = Dp[ii][kk];
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
						temp2 += (X[ii] + 1.0e-40) / Dp[kk][ii];
					}
				// sycl::step(ceil(temp1), 0.0) == 1 <=> 0.0 >= ceil(temp1)
				if (!(0.0 < ceil(temp1)))
					Dk[kk] = xf_fit(vs.fit_Dkj[kk * NS + kk], L) / p;
				else
					Dk[kk] = temp1 / temp2 / rho * C_total;
				Dk[kk] *= 1.0e-1;
			}
		}
		else
		{
			Dk[0] = xf_fit(vs.fit_Dkj[0], L) / p;
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

// GetWallViscousFlux{X,Y,Z}: Flux_wall -= F_wall_v at the face of direction DIR above cell id
template <class C, int DIR>
__device__ __forceinline__ void visc_face(const XfDev &d, const XfVisc &vs, const double *__restrict__ U, double *__restrict__ Fw, const long long id)
{
	constexpr int NS = C::NS, E = C::E;
	const long long s = DIR == 0 ? 1 : (DIR == 1 ? d.sY : d.sZ);
	const long long id_m1 = id - s, id_p1 = id + s, id_p2 = id + 2 * s;
	const double _sxtn = 1.0 / 16.0, _twfr = 1.0 / 24.0, _OT = 1.0 / 3.0;
	const double _dl = DIR == 0 ? d._dx : (DIR == 1 ? d._dy : d._dz);
	const double tX = d.DimX ? 1.0 : 0.0, tY = d.DimY ? 1.0 : 0.0, tZ = d.DimZ ? 1.0 : 0.0;
	auto avg = [&](const double *q) { return (9.0 * (q[id_p1] + q[id]) - (q[id_p2] + q[id_m1])) * _sxtn; };
	auto grad = [&](const double *q) { return (27.0 * (q[id_p1] - q[id]) - (q[id_p2] - q[id_m1])) * _dl * _twfr; };
	const double *Vd = vs.Vde;
	const long long N = d.N;
	const double mue = avg(vs.va);
	const double lamada = -2.0 * _OT * mue;
	double f_x, f_y, f_z, u_hlf, v_hlf, w_hlf;
	if constexpr (DIR == 0)
	{ // Ducy 3, Ducz 6, Dvcy 4, Dwcz 8
		f_x = (2.0 * mue + lamada) * (27.0 * (d.u[id_p1] - d.u[id]) - (d.u[id_p2] - d.u[id_m1])) * _dl * _twfr;
		f_x += lamada * (9.0 * (Vd[4 * N + id_p1] + Vd[4 * N + id]) - (Vd[4 * N + id_p2] + Vd[4 * N + id_m1]) + 9.0 * (Vd[8 * N + id_p1] + Vd[8 * N + id]) - (Vd[8 * N + id_p2] + Vd[8 * N + id_m1])) * _sxtn;
		f_y = mue * (27.0 * (d.v[id_p1] - d.v[id]) - (d.v[id_p2] - d.v[id_m1])) * _dl * _twfr * tY;
		f_y += mue * (9.0 * (Vd[3 * N + id_p1] + Vd[3 * N + id]) - (Vd[3 * N + id_p2] + Vd[3 * N + id_m1])) * _sxtn * tY;
		f_z = mue * (27.0 * (d.w[id_p1] - d.w[id]) - (d.w[id_p2] - d.w[id_m1])) * _dl * _twfr * tZ;
		f_z += mue * (9.0 * (Vd[6 * N + id_p1] + Vd[6 * N + id]) - (Vd[6 * N + id_p2] + Vd[6 * N + id_m1])) * _sxtn * tZ;
		u_hlf = avg(d.u), v_hlf = avg(d.v) * tY, w_hlf = avg(d.w) * tZ;
	}
	else if constexpr (DIR == 1)
	{ // Dvcx 1, Dvcz 7, Ducx 0, Dwcz 8
		f_x = mue * (27.0 * (d.u[id_p1] - d.u[id]) - (d.u[id_p2] - d.u[id_m1])) * _dl * _twfr * tX;
		f_x += mue * (9.0 * (Vd[1 * N + id_p1] + Vd[1 * N + id]) - (Vd[1 * N + id_p2] + Vd[1 * N + id_m1])) * _sxtn * tX;
		f_y = (2.0 * mue + lamada) * (27.0 * (d.v[id_p1] - d.v[id]) - (d.v[id_p2] - d.v[id_m1])) * _dl * _twfr;
		f_y += lamada * (9.0 * (Vd[0 * N + id_p1] + Vd[0 * N + id]) - (Vd[0 * N + id_p2] + Vd[0 * N + id_m1]) + 9.0 * (Vd[8 * N + id_p1] + Vd[8 * N + id]) - (Vd[8 * N + id_p2] + Vd[8 * N + id_m1])) * _sxtn;
		f_z = mue * (27.0 * (d.w[id_p1] - d.w[id]) - (d.w[id_p2] - d.w[id_m1])) * _dl * _twfr * tZ;
		f_z += mue * (9.0 * (Vd[7 * N + id_p1] + Vd[7 * N + id]) - (Vd[7 * N + id_p2] + Vd[7 * N + id_m1])) * _sxtn * tZ;
		u_hlf = avg(d.u) * tX, v_hlf = avg(d.v), w_hlf = avg(d.w) * tZ;
	}
	else
	{ // Dwcx 2, Dwcy 5, Ducx 0, Dvcy 4
		f_x = mue * (27.0 * (d.u[id_p1] - d.u[id]) - (d.u[id_p2] - d.u[id_m1])) * _dl * _twfr * tX;
		f_x += mue * (9.0 * (Vd[2 * N + id_p1] + Vd[2 * N + id]) - (Vd[2 * N + id_p2] + Vd[2 * N + id_m1])) * _sxtn * tX;
		f_y = mue * (27.0 * (d.v[id_p1] - d.v[id]) - (d.v[id_p2] - d.v[id_m1])) * _dl * _twfr * tY;
		f_y += mue * (9.0 * (Vd[5 * N + id_p1] + Vd[5 * N + id]) - (Vd[5 * N + id_p2] + Vd[5 * N + id_m1])) * _sxtn * tY;
		f_z = (2.0 * mue + lamada) * (27.0 * (d.w[id_p1] - d.w[id]) - (d.w[id_p2] - d.w[id_m1])) * _dl * _twfr;
		f_z += lamada * (9.0 * (Vd[0 * N + id_p1] + Vd[0 * N + id]) - (Vd[0 * N + id_p2] + Vd[0 * N + id_m1]) + 9.0 * (Vd[4 * N + id_p1] + Vd[4 * N + id]) - (Vd[4 * N + id_p2] + Vd[4 * N + id_m1])) * _sxtn;
		u_hlf = avg(d.u) * tX, v_hlf = avg(d.v) * tY, w_hlf = avg(d.w);
	}
	double Fv[E];
	Fv[0] = 0.0, Fv[1] = f_x, Fv[2] = f_y, Fv[3] = f_z;
	Fv[4] = f_x * u_hlf + f_y * v_hlf + f_z * w_hlf;
	if (vs.heat)
	{ // MARCO_VIS_HEAT
		double kk_ = avg(vs.tca);
		kk_ *= grad(d.T);
		Fv[4] += kk_;
	}
	if (vs.diffu)
	{ // MARCO_VIS_Diffu
		const double rho_wall = avg(U);
		double CorrectTerm = 0.0, Dim_Yil = 1.0E-20;
		double Yi_wall[NS];
#pragma unroll
		for (int l = 0; l < NS; l++)
		{
			const double hi_wall = avg(vs.hi + l * N), Dim_wall = avg(vs.Dkm + l * N);
			double Yil_wall = 0.0;
			if constexpr (C::COP)
			{
				const double *Y = d.y + l * N;
				const double yl = vs.lim[2 * NS + l], dlm = vs.lim[3 * NS + l];
				Yil_wall = xf_min(xf_max(grad(Y), -yl), yl);
				Yi_wall[l] = xf_min(xf_max(avg(Y), 1.0E-20), 1.0);
				Dim_Yil = xf_min(xf_max(Dim_wall * Yil_wall, -dlm), dlm);
				CorrectTerm += Dim_Yil;
			}
			(void)Yil_wall;
			Fv[4] += rho_wall * hi_wall * Dim_Yil;
		}
		CorrectTerm *= rho_wall;
#pragma unroll
		for (int p = 5; p < E; p++)
			Fv[p] = rho_wall * Dim_Yil - Yi_wall[p - 5] * CorrectTerm;
	}
	else
	{
#pragma unroll
		for (int p = 5; p < E; p++)
			Fv[p] = 0.0;
	}
#pragma unroll
	for (int n = 0; n < E; n++)
		Fw[n * N + id] -= Fv[n];
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
static int visc_t(const XfDev &d, const XfThermo &th, const XfVisc &vs, const double *U, const int bc[6], cudaStream_t s, long long *launches)
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
