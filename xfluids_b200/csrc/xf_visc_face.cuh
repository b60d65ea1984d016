// xf_visc_face.cuh -- the viscous / heat-conduction / species-diffusion wall flux of one face, shared by the stand-alone
// kernel of the block-level API (k_visc_flux3, xf_visc.cuh) and by the tails of the sweeps (k_sweep<..., VISC = true>), where the flux is
// subtracted from the inviscid wall flux while that is still in registers: Fw is written once and never read back.
#pragma once
#include "xf_types.h"

// F_wall_v of GetWallViscousFlux{X,Y,Z} (Fourth_Order/Visc_Order_kernels.hpp:76-340 with the macros of Flux_discrete.h) at the face of
// direction DIR above cell id; VS = XfVisc (the stand-alone kernel) or XfViscF (the tail of a sweep)
template <class C, int DIR, class VS>
__device__ __forceinline__ void visc_face_flux(const XfDev &d, const VS &vs, const double *__restrict__ U, const long long id, double *__restrict__ Fv)
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
}

