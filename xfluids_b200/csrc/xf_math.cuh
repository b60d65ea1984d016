// xf_math.cuh -- device math of the inviscid path: NASA-9 thermo + limited Newton T(e,Y), WENO5-JS / WENO7-JS,
// Roe-averaged multi-species sound speed and the characteristic flux at one face.
//
// Written for FP64 on sm_100a.  Every expression keeps the reference's association order, so that the
// strict build (-fmad=false) performs the same sequence of IEEE-754 roundings as the reference CPU path;
// the only places where bits can differ are libdevice log() vs glibc log().  Structural zeros / +-1 of
// the eigenvector matrices are exploited (x*0 and +0 are exact no-ops, *1 and *-1 are exact), which is
// where this differs from the reference's dense Emax x Emax loops (Eigen_callback.h:194-230).
//
// Reference files followed: solver_Ini/Thermo_device.h, solver_Ini/Mixing_device.h,
// solver_Reconstruction/schemes/WENO{5,7}s_schemes.hpp, FDM_Method/positive-definite_eigen/
// {Utils_device.hpp,Eigen_matrix.hpp,Eigen_callback.h}, include/global_marco.h.
#pragma once
#include "xf_types.h"
#include "xf_log.cuh"
#include "xf_exp.cuh"

#define XF_DEV __device__ __forceinline__
#if !defined(XF_THERMO_STATIC) && !defined(XF_THERMO_DYN)
#define XF_THERMO_DYN // NASA-9 range as a run-time constant-bank index (one code copy, 94 registers, no spills); XF_THERMO_STATIC: 3 copies
#endif

template <int NS_, bool COP_>
struct XfCfg
{
	static constexpr int NS = NS_;
	static constexpr bool COP = COP_;
	static constexpr int E = COP_ ? NS_ + 4 : 5;   // Emax
	static constexpr int NC = COP_ ? NS_ - 1 : 0;  // NUM_COP
};

// sycl::min / sycl::max semantics of the host path: (b < a) ? b : a  /  (a < b) ? b : a
XF_DEV double xf_min(double a, double b) { return (b < a) ? b : a; }
XF_DEV double xf_max(double a, double b) { return (a < b) ? b : a; }

// ------------------------------------------------------------------------------------------------
// NASA-9 species thermo (Thermo_device.h:10-23, 62-80)
// ------------------------------------------------------------------------------------------------
template <int R>
XF_DEV double xf_cp_r(const XfThermo &th, int n, double T, double _T)
{
	const double *a = th.ccoef[R][n];
	return th.Ri[n] * ((a[0] * _T + a[1]) * _T + a[2] + (a[3] + (a[4] + (a[5] + a[6] * T) * T) * T) * T);
}
// a / b given y = RN(1 / b): q0 = RN(a y), r = a - b q0 (exact in one fma), q = RN(q0 + r y) is the correctly rounded quotient
// (Markstein's theorem: y within 1/2 ulp of 1/b, q0 within 1 ulp of a/b; no over/underflow for NASA-9 coefficients over 200 K <= b),
// i.e. bit-identical to the IEEE division the reference performs -- the NS species of one evaluation share the reciprocal of T
// instead of running NS full division sequences.  (tools/check_shared_reciprocal.c: 3e8 random pairs, 0 mismatches.)
XF_DEV double xf_div_shared(double a, double b, double y)
{
#ifdef XF_NO_DIVSHARED
	return a / b;
#else
	const double q0 = a * y;
	const double r = fma(-q0, b, a);
	return fma(r, y, q0);
#endif
}
template <int R>
XF_DEV double xf_h_r(const XfThermo &th, int n, double T, double _T, double lnT)
{
	const double *h = th.hcoef[R][n];
	return th.Ri[n] * (xf_div_shared(h[0], T, _T) + h[1] * lnT + (h[2] + (h[3] + (h[4] + (h[5] + h[6] * T) * T) * T) * T) * T + h[7]);
}
XF_DEV int xf_range(double T) { return (T >= 1000.0 && T < 6000.0) ? 1 : ((T < 1000.0) ? 0 : 2); }
// the same polynomials with the temperature range as a run-time index into the constant-bank tables: one copy of the code
// instead of three (XF_THERMO_DYN)
XF_DEV double xf_cp_dyn(const XfThermo &th, int r, int n, double T, double _T)
{
	const double *a = th.ccoef[r][n];
	return th.Ri[n] * ((a[0] * _T + a[1]) * _T + a[2] + (a[3] + (a[4] + (a[5] + a[6] * T) * T) * T) * T);
}

// mixture Cp at T0 (get_CopCp, Mixing_device.h:61-68)
template <class C>
XF_DEV double xf_mix_cp(const XfThermo &th, const double *yi, double T0)
{
	const double T = xf_max(T0, 200.0), _T = 1.0 / T;
	const int r = xf_range(T);
	double cp = 0.0;
#ifdef XF_THERMO_DYN
#pragma unroll
	for (int n = 0; n < C::NS; n++)
		cp += yi[n] * xf_cp_dyn(th, r, n, T, _T);
	return cp;
#endif
	if (r == 0)
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			cp += yi[n] * xf_cp_r<0>(th, n, T, _T);
	}
	else if (r == 1)
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			cp += yi[n] * xf_cp_r<1>(th, n, T, _T);
	}
	else
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			cp += yi[n] * xf_cp_r<2>(th, n, T, _T);
	}
	return cp;
}
// species enthalpies at T0 into hi[] (get_Enthalpy_NASA incl. the linear extension below 200 K)
template <class C>
XF_DEV void xf_species_h(const XfThermo &th, double T0, double *hi)
{
	const double T = xf_max(T0, 200.0), lnT = xf_log(T), _T = 1.0 / T; // glibc's log, bit for bit (xf_log.cuh)
	const int r = xf_range(T);
#ifdef XF_THERMO_DYN
#pragma unroll
	for (int n = 0; n < C::NS; n++)
	{
		const double *h = th.hcoef[r][n];
		hi[n] = th.Ri[n] * (xf_div_shared(h[0], T, _T) + h[1] * lnT + (h[2] + (h[3] + (h[4] + (h[5] + h[6] * T) * T) * T) * T) * T + h[7]);
	}
#else
	if (r == 0)
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			hi[n] = xf_h_r<0>(th, n, T, _T, lnT);
	}
	else if (r == 1)
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			hi[n] = xf_h_r<1>(th, n, T, _T, lnT);
	}
	else
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			hi[n] = xf_h_r<2>(th, n, T, _T, lnT);
	}
#endif
	if (T0 < 200.0)
	{
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			hi[n] += xf_cp_r<0>(th, n, 200.0, 1.0 / 200.0) * (T0 - 200.0);
	}
}

// limited Newton iteration for T from internal energy (get_T / sub_FuncT, Mixing_device.h:159-193).
// R = get_CopR(yi) is loop-invariant and passed in.
template <class C>
XF_DEV double xf_get_T(const XfThermo &th, const double *yi, double e, double T0, double R)
{
	double T = T0;
	for (int it = 1; it < 101; it++)
	{
		double hi[C::NS];
		xf_species_h<C>(th, T, hi);
		double h = 0.0;
#pragma unroll
		for (int n = 0; n < C::NS; n++)
			h += hi[n] * yi[n];
		const double Cp = xf_mix_cp<C>(th, yi, T);
		const double func_T = h - R * T - e;
		const double dfunc_T = Cp - R;
		double df = xf_min(func_T / (dfunc_T + 1.0e-30), 1e-3 * T);
		df = xf_max(df, -1e-2 * T);
		T = T - df;
		if (fabs(df) <= 1.0e-6)
			break;
	}
	return T;
}

// ------------------------------------------------------------------------------------------------
// WENO5-JS as written in weno5old_BODY (WENO5s_schemes.hpp:12-69) and WENO7-JS (WENO7s_schemes.hpp:8-128)
// ------------------------------------------------------------------------------------------------
// Multiplications by 2 and 4 are exact in binary floating point, so  a - 2.0*b  and  fma(-2.0, b, a)  round the same real
// number once and are bit-identical; those (and only those) products are fused explicitly here, in both build flavours.
XF_DEV double xf_weno5_body(double v1, double v2, double v3, double v4, double v5)
{
	double a1, a2, a3;
	a1 = fma(-2.0, v2, v1) + v3;                 // v1 - 2.0 * v2 + v3
	double s1 = 13.0 * a1 * a1;
	a1 = fma(-4.0, v2, v1) + 3.0 * v3;           // v1 - 4.0 * v2 + 3.0 * v3
	s1 += 3.0 * a1 * a1;
	a1 = fma(-2.0, v3, v2) + v4;                 // v2 - 2.0 * v3 + v4
	double s2 = 13.0 * a1 * a1;
	a1 = v2 - v4;
	s2 += 3.0 * a1 * a1;
	a1 = fma(-2.0, v4, v3) + v5;                 // v3 - 2.0 * v4 + v5
	double s3 = 13.0 * a1 * a1;
	a1 = fma(-4.0, v4, 3.0 * v3) + v5;           // 3.0 * v3 - 4.0 * v4 + v5
	s3 += 3.0 * a1 * a1;
	s1 += 1.0E-6, s2 += 1.0E-6, s3 += 1.0E-6;
	a1 = 0.1 * s2 * s2 * s3 * s3;
	a2 = 0.6 * s1 * s1 * s3 * s3;
	a3 = 0.3 * s1 * s1 * s2 * s2;
	const double tw1 = 1.0 / (a1 + a2 + a3);
	a1 = a1 * tw1, a2 = a2 * tw1, a3 = a3 * tw1;
	s1 = a1 * (fma(2.0, v1, -(7.0 * v2)) + 11.0 * v3);   // 2.0 * v1 - 7.0 * v2 + 11.0 * v3
	s2 = a2 * fma(2.0, v4, 5.0 * v3 - v2);                // -v2 + 5.0 * v3 + 2.0 * v4
	s3 = a3 * (fma(2.0, v3, 5.0 * v4) - v5);              // 2.0 * v3 + 5.0 * v4 - v5
	return (s1 + s2 + s3);
}
// WENO-CU6 (WENOCU6_BODYGPU, WENO6s_schemes.hpp:5-50; constants Utils_schemes.hpp:5-15, global_setup.h:36-37), expression order
// as written.  The same exact-product fusions as in xf_weno5_body ( *2, *4 ).
XF_DEV double xf_wenocu6_body(double v1, double v2, double v3, double v4, double v5, double v6, double epsilon)
{
	const double _six = 1.0 / 6.0, _sxtn = 1.0 / 16.0, _twfr = 1.0 / 24.0, _ohtz = 1.0 / 120.0, _ohff = 1.0 / 144.0, _ftss = 1.0 / 5760.0;
	const double a2a2 = 13.0 / 3.0, a3a3 = 3129.0 / 80.0, a4a4 = 87617.0 / 140.0, a3a5 = 14127.0 / 224.0, a5a5 = 252337135.0 / 16128.0;
	const double s11 = fma(-2.0, v2, v1) + v3;
	const double s12 = fma(-4.0, v2, v1) + 3.0 * v3;
	const double s1 = 13.0 * s11 * s11 + 3.0 * s12 * s12;
	const double s21 = fma(-2.0, v3, v2) + v4;
	const double s22 = v2 - v4;
	const double s2 = 13.0 * s21 * s21 + 3.0 * s22 * s22;
	const double s31 = fma(-2.0, v4, v3) + v5;
	const double s32 = fma(-4.0, v4, 3.0 * v3) + v5;
	const double s3 = 13.0 * s31 * s31 + 3.0 * s32 * s32;
	const double tau61 = (259.0 * v6 - 1895.0 * v5 + 6670.0 * v4 - 2590.0 * v3 - 2785.0 * v2 + 341.0 * v1) * _ftss;
	const double tau62 = -(v5 - 12.0 * v4 + 22.0 * v3 - 12.0 * v2 + v1) * _sxtn;
	const double tau63 = -(7.0 * v6 - 47.0 * v5 + 94.0 * v4 - 70.0 * v3 + 11.0 * v2 + 5.0 * v1) * _ohff;
	const double tau64 = (fma(-4.0, v2, fma(-4.0, v4, v5) + 6.0 * v3) + v1) * _twfr;           // v5 - 4 v4 + 6 v3 - 4 v2 + v1
	const double tau65 = -(-v6 + 5.0 * v5 - 10.0 * v4 + 10.0 * v3 - 5.0 * v2 + v1) * _ohtz;
	// a1a1 = 1 (exact, skipped), a1a3 = 0.5, a2a4 = 4.2, a1a5 = 0.125
	const double s6 = (tau61 * tau61 + tau62 * tau62 * a2a2 + tau61 * tau63 * 0.5 + tau63 * tau63 * a3a3 + tau62 * tau64 * 4.2 + tau61 * tau65 * 0.125 +
					   tau64 * tau64 * a4a4 + tau63 * tau65 * a3a5 + tau65 * tau65 * a5a5) * 12.0;
	const double s55 = (s1 + s3 + 4.0 * s2) * _six;
	const double s5 = fabs(s6 - s55);
#ifndef XF_CU6_NOZSKIP // measured (512x256x256, SBI): 52 ms -> 33 ms per sweep direction
	// s5 == 0 (exactly uniform stencil): 0 / (s + epsilon) = +0 for every positive denominator; skips four divisions whose zero
	// numerator sends libdevice's division to its slow path
	double q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
	if (s5 != 0.0)
		q1 = s5 / (s1 + epsilon), q2 = s5 / (s2 + epsilon), q3 = s5 / (s3 + epsilon), q4 = s5 / (s6 + epsilon);
	const double r1 = 20.0 + q1, r2 = 20.0 + q2, r3 = 20.0 + q3, r4 = 20.0 + q4;
#else
	const double r1 = 20.0 + s5 / (s1 + epsilon);
	const double r2 = 20.0 + s5 / (s2 + epsilon);
	const double r3 = 20.0 + s5 / (s3 + epsilon);
	const double r4 = 20.0 + s5 / (s6 + epsilon);
#endif
	const double a1 = 0.05 * r1, a2 = 0.45 * r2, a3 = 0.45 * r3, a4 = 0.05 * r4;
	const double tw1 = 1.0 / (a1 + a2 + a3 + a4);
	const double w1 = a1 * tw1, w2 = a2 * tw1, w3 = a3 * tw1, w4 = a4 * tw1;
	double temp = 0.0;
	temp += w1 * (fma(2.0, v1, -(7.0 * v2)) + 11.0 * v3);   // 2 v1 - 7 v2 + 11 v3
	temp += w2 * fma(2.0, v4, 5.0 * v3 - v2);                // -v2 + 5 v3 + 2 v4
	temp += w3 * (fma(2.0, v3, 5.0 * v4) - v5);              // 2 v3 + 5 v4 - v5
	temp += w4 * fma(2.0, v6, 11.0 * v4 - 7.0 * v5);         // 11 v4 - 7 v5 + 2 v6
	return temp;
}
XF_DEV double xf_weno7_body(double v1, double v2, double v3, double v4, double v5, double v6, double v7)
{
	const double ep = 1.0e-7;
	const double C0 = 1.0 / 35.0, C1 = 12.0 / 35.0, C2 = 18.0 / 35.0, C3 = 4.0 / 35.0;
	const double S10 = -2.0 / 6.0 * v1 + 9.0 / 6.0 * v2 - 18.0 / 6.0 * v3 + 11.0 / 6.0 * v4;
	const double S11 = 1.0 / 6.0 * v2 - 6.0 / 6.0 * v3 + 3.0 / 6.0 * v4 + 2.0 / 6.0 * v5;
	const double S12 = -2.0 / 6.0 * v3 - 3.0 / 6.0 * v4 + 6.0 / 6.0 * v5 - 1.0 / 6.0 * v6;
	const double S13 = -11.0 / 6.0 * v4 + 18.0 / 6.0 * v5 - 9.0 / 6.0 * v6 + 2.0 / 6.0 * v7;
	const double S20 = fma(2.0, v4, fma(4.0, v2, -v1) - 5.0 * v3);          // -v1 + 4.0 * v2 - 5.0 * v3 + 2.0 * v4
	const double S21 = fma(-2.0, v4, v3) + v5;                               // v3 - 2.0 * v4 + v5
	const double S22 = fma(-2.0, v5, v4) + v6;                               // v4 - 2.0 * v5 + v6
	const double S23 = fma(4.0, v6, fma(2.0, v4, -(5.0 * v5))) - v7;         // 2.0 * v4 - 5.0 * v5 + 4.0 * v6 - 1.0 * v7
	const double S30 = -v1 + 3.0 * v2 - 3.0 * v3 + v4;
	const double S31 = -v2 + 3.0 * v3 - 3.0 * v4 + v5;
	const double S32 = -v3 + 3.0 * v4 - 3.0 * v5 + v6;
	const double S33 = -v4 + 3.0 * v5 - 3.0 * v6 + v7;
	const double S0 = S10 * S10 + 13.0 / 12.0 * S20 * S20 + 1043.0 / 960.0 * S30 * S30 + 1.0 / 12.0 * S10 * S30;
	const double S1 = S11 * S11 + 13.0 / 12.0 * S21 * S21 + 1043.0 / 960.0 * S31 * S31 + 1.0 / 12.0 * S11 * S31;
	const double S2 = S12 * S12 + 13.0 / 12.0 * S22 * S22 + 1043.0 / 960.0 * S32 * S32 + 1.0 / 12.0 * S12 * S32;
	const double S3 = S13 * S13 + 13.0 / 12.0 * S23 * S23 + 1043.0 / 960.0 * S33 * S33 + 1.0 / 12.0 * S13 * S33;
	const double a0 = C0 / ((ep + S0) * (ep + S0));
	const double a1 = C1 / ((ep + S1) * (ep + S1));
	const double a2 = C2 / ((ep + S2) * (ep + S2));
	const double a3 = C3 / ((ep + S3) * (ep + S3));
	const double sum = a0 + a1 + a2 + a3;
	const double _sum = 1.0 / sum; // four quotients by the same denominator: one division + exact residual corrections (xf_div_shared)
	const double W0 = xf_div_shared(a0, sum, _sum), W1 = xf_div_shared(a1, sum, _sum), W2 = xf_div_shared(a2, sum, _sum), W3 = xf_div_shared(a3, sum, _sum);
	const double q0 = -3.0 / 12.0 * v1 + 13.0 / 12.0 * v2 - 23.0 / 12.0 * v3 + 25.0 / 12.0 * v4;
	const double q1 = 1.0 / 12.0 * v2 - 5.0 / 12.0 * v3 + 13.0 / 12.0 * v4 + 3.0 / 12.0 * v5;
	const double q2 = -1.0 / 12.0 * v3 + 7.0 / 12.0 * v4 + 7.0 / 12.0 * v5 - 1.0 / 12.0 * v6;
	const double q3 = 3.0 / 12.0 * v4 + 13.0 / 12.0 * v5 - 5.0 / 12.0 * v6 + 1.0 / 12.0 * v7;
	return W0 * q0 + W1 * q1 + W2 * q2 + W3 * q3;
}

// Lax-Friedrichs splitting pp = 0.5 (ff + av uf), mm = 0.5 (ff - av uf) and the +/- reconstructions of one
// characteristic field (Eigen_callback.h:194-208 / 146-163).  Deliberately NOT inlined: the field loop of xf_face_flux is
// fully unrolled (every field has its own sparse projection), and nine inlined copies of this body push the sweep kernel
// to ~78 KB of SASS, which showed up in ncu as the top stall reason ("no instruction").
static __device__ __noinline__ double xf_split_weno5(double av, double u0, double u1, double u2, double u3, double u4, double u5,
													 double f0, double f1, double f2, double f3, double f4, double f5)
{
	// stencil s=0..5 <-> cells i-2..i+3 ; weno5old_GPU(&pp[3],&mm[3]): plus uses i-2..i+2, minus i+3..i-1
	// 0.5 * (f + av * u) == 0.5 * f + (0.5 * av) * u bit for bit (scaling by a power of two commutes with rounding), which
	// needs one multiplication less per stencil point than halving pp and mm separately
	// f0..f5 arrive halved (the stencil stages 0.5 * F, see stage_cell)
	const double hv = 0.5 * av;
	const double a0 = hv * u0, a1 = hv * u1, a2 = hv * u2, a3 = hv * u3, a4 = hv * u4, a5 = hv * u5;
	const double h0 = f0, h1 = f1, h2 = f2, h3 = f3, h4 = f4, h5 = f5;
	const double p1 = h0 + a0, p2 = h1 + a1, p3 = h2 + a2, p4 = h3 + a3, p5 = h4 + a4;
	const double m1 = h5 - a5, m2 = h4 - a4, m3 = h3 - a3, m4 = h2 - a2, m5 = h1 - a1;
	return (xf_weno5_body(p1, p2, p3, p4, p5) + xf_weno5_body(m1, m2, m3, m4, m5)) * (1.0 / 6.0);
}
// WENOCU6_GPU(&pp[3], &mm[3], dl) (WENO6s_schemes.hpp:53-78): plus side cells i-2..i+3, minus side i+3..i-2; epsilon = 1e-8 dl dl
static __device__ __noinline__ double xf_split_wenocu6(double av, double u0, double u1, double u2, double u3, double u4, double u5,
													   double f0, double f1, double f2, double f3, double f4, double f5, double epsilon)
{
	const double hv = 0.5 * av;
	const double a0 = hv * u0, a1 = hv * u1, a2 = hv * u2, a3 = hv * u3, a4 = hv * u4, a5 = hv * u5;
	const double h0 = f0, h1 = f1, h2 = f2, h3 = f3, h4 = f4, h5 = f5; // halved by the staging
	return (xf_wenocu6_body(h0 + a0, h1 + a1, h2 + a2, h3 + a3, h4 + a4, h5 + a5, epsilon) +
			xf_wenocu6_body(h5 - a5, h4 - a4, h3 - a3, h2 - a2, h1 - a1, h0 - a0, epsilon)) * (1.0 / 6.0);
}
static __device__ __noinline__ double xf_split_weno7(double av, double u0, double u1, double u2, double u3, double u4, double u5, double u6, double u7,
											  double f0, double f1, double f2, double f3, double f4, double f5, double f6, double f7)
{
	const double uf[8] = {u0, u1, u2, u3, u4, u5, u6, u7}, ff[8] = {f0, f1, f2, f3, f4, f5, f6, f7};
	double pp[8], mm[8];
	// ff arrives halved: 0.5 * (ff +- av uf) == 0.5 ff +- (0.5 av) uf bit for bit (scaling by 0.5 commutes with rounding)
	const double hv = 0.5 * av;
#pragma unroll
	for (int s = 0; s < 8; s++)
	{
		const double au = hv * uf[s];
		pp[s] = ff[s] + au;
		mm[s] = ff[s] - au;
	}
	// weno7_P(&pp[3]): f[-3..3]; weno7_M(&mm[3]): k=1, v1=f[4] ... v7=f[-2]
	return xf_weno7_body(pp[0], pp[1], pp[2], pp[3], pp[4], pp[5], pp[6]) + xf_weno7_body(mm[7], mm[6], mm[5], mm[4], mm[3], mm[2], mm[1]);
}

// ------------------------------------------------------------------------------------------------
// Roe state at a face (MARCO_ROE global_marco.h:26-34, ReconstructSoundSpeed Utils_device.hpp:102-140,
// SoundSpeedMultiSpecies :42-79, MARCO_NOCOPC2 / MARCO_PREEIGEN Eigen_callback.h:87-106)
// ------------------------------------------------------------------------------------------------
template <class C>
struct XfRoe
{
	double u, v, w, H, c, c1, b1, b2, b3;
	double z[C::NC > 0 ? C::NC : 1], y[C::NC > 0 ? C::NC : 1];
};

// per-cell face-side inputs gathered from the work arrays
template <class C>
struct XfSide
{
	double rho, u, v, w, H, p;
	double g3, dpdrho, e, prho;                                    // COP only
	double y[C::NC > 0 ? C::NC : 1], dpdrhoi[C::NC > 0 ? C::NC : 1]; // COP only
};

template <class C>
XF_DEV void xf_roe_state(const XfSide<C> &l, const XfSide<C> &r, double gamma0, XfRoe<C> &R)
{
	const double D = sqrt(r.rho / l.rho);
	const double D1 = 1.0 / (D + 1.0);
	R.u = (l.u + D * r.u) * D1;
	R.v = (l.v + D * r.v) * D1;
	R.w = (l.w + D * r.w) * D1;
	R.H = (l.H + D * r.H) * D1;
	const double _P = (l.p + D * r.p) * D1;
	const double _rho = sqrt(r.rho * l.rho);
	double c2, b1, b3;
	if constexpr (C::COP)
	{
		constexpr int NC = C::NC;
		double dpi[NC > 0 ? NC : 1], drhoi[NC > 0 ? NC : 1];
#pragma unroll
		for (int n = 0; n < NC; n++)
			R.y[n] = (l.y[n] + D * r.y[n]) * D1;
		const double Gamma = (l.g3 + D * r.g3) * D1;
		const double _dpdrho = (l.dpdrho + D * r.dpdrho) * D1;
#pragma unroll
		for (int n = 0; n < NC; n++)
		{
			dpi[n] = (l.dpdrhoi[n] + D * r.dpdrhoi[n]) * D1;
			drhoi[n] = r.rho * r.y[n] - l.rho * l.y[n];
		}
		const double du = r.u - l.u, dv = r.v - l.v, dw = r.w - l.w;
		const double _prho = (l.prho + D * r.prho) * D1 + 0.5 * D * D1 * D1 * (du * du + dv * dv + dw * dw);
		const double _dpdE = ((l.g3 - 1.0) + D * (r.g3 - 1.0)) * D1;
		const double _dpde = ((l.g3 - 1.0) * l.rho + D * ((r.g3 - 1.0) * r.rho)) * D1;
		const double dp = r.p - l.p, drho = r.rho - l.rho, de = r.e - l.e;
		// SoundSpeedMultiSpecies
		double Sum_dpdrhoi = 0.0, Sum_dpdrhoi2 = 0.0, Sum_Yidpdrhoi = 0.0;
#pragma unroll
		for (int n = 0; n < NC; n++)
		{
			Sum_dpdrhoi += dpi[n] * drhoi[n];
			Sum_dpdrhoi2 += dpi[n] * drhoi[n] * dpi[n] * drhoi[n];
		}
		const double temp1 = dp - (_dpdrho * drho + _dpde * de + Sum_dpdrhoi);
		const double temp = temp1 / (_dpdrho * _dpdrho * drho * drho + _dpde * de * _dpde * de + Sum_dpdrhoi2 + 1e-19);
#pragma unroll
		for (int n = 0; n < NC; n++)
			Sum_Yidpdrhoi += R.y[n] * dpi[n];
		const double _dpdE_new = _dpdE + _dpdE * _dpdE * de * _rho * temp;
		const double _dpdrho_new = _dpdrho + _dpdrho * _dpdrho * drho * temp;
		const double csqr = _dpdrho_new + _dpdE_new * _prho + Sum_Yidpdrhoi;
		b1 = _dpdE_new / csqr;
		b3 = 0.0;
#pragma unroll
		for (int n = 0; n < NC; n++)
		{
			const double dpn = dpi[n] + dpi[n] * dpi[n] * drhoi[n] * temp;
			R.z[n] = -dpn / _dpdE_new;
			b3 += R.y[n] * R.z[n];
		}
		b3 *= b1;
		// c2 <= 0 fallback: Gamma * P * rho (a product, as written; Utils_device.hpp:136-137)
		const double c2w = (0.0 < csqr) ? 0.0 : 1.0;
		c2 = Gamma * _P * _rho * c2w + (1.0 - c2w) * csqr;
	}
	else
	{
		c2 = gamma0 * _P / _rho;
		b1 = (gamma0 - 1.0) / c2;
		b3 = 0.0;
	}
	const double q2 = R.u * R.u + R.v * R.v + R.w * R.w;
	R.c = sqrt(c2);
	R.b1 = b1, R.b3 = b3;
	R.b2 = 1.0 + b1 * q2 - b1 * R.H;
	R.c1 = 1.0 / R.c;
}

// ------------------------------------------------------------------------------------------------
// Characteristic-wise split flux at one face (MARCO_FLUXWALL_WENO5/7, Eigen_callback.h:127-230, with the
// rows of L / columns of R from Eigen_matrix.hpp:7-455).
//
// ST is a stencil accessor: ST::U(s,n), ST::F(s,n) conserved variable / HALF the physical flux component n of
// stencil cell s (s = 0..NST-1 <-> offset m = s-P along the sweep), ST::lam(s,t) = |u_d - c|, |u_d|,
// |u_d + c| (t = 0,1,2) of that cell.  Face lies between s = P and s = P+1.
// ------------------------------------------------------------------------------------------------
template <int WENO>
struct XfStencil
{
	static constexpr int P = WENO == 7 ? 3 : 2;    // cells left of the face's left cell
	static constexpr int NST = WENO == 7 ? 8 : 6;  // cells actually used by the reconstruction (WENO5-JS and WENO-CU6: i-2..i+3)
};

template <class C, int DIR, int WENO, class ST>
XF_DEV void xf_face_flux(const ST &st, const XfRoe<C> &R, int alpha, const double *glf /*3*/, double dl, double *Fw /*E*/)
{
	constexpr int E = C::E, NC = C::NC, NST = XfStencil<WENO>::NST;
	constexpr int ENT = DIR + 1; // row/column index of the entropy wave: x 1, y 2, z 3
	const double un = DIR == 0 ? R.u : (DIR == 1 ? R.v : R.w);
	const double un_c = un * R.c1;
	const double b1 = R.b1, b2 = R.b2, b3 = R.b3, c1 = R.c1;

	// local Lax-Friedrichs wave speeds: max over the stencil of |u_d - c|, |u_d|, |u_d + c| -- the same for every field of a
	// family, so formed once per face (the reference re-forms it per field; max is exact, the order is kept)
	double lmax[3] = {0.0, 0.0, 0.0};
	if (alpha == 2 && WENO != 7)
	{
#pragma unroll
		for (int t = 0; t < 3; t++)
#pragma unroll
			for (int s = 0; s < NST; s++)
				lmax[t] = xf_max(lmax[t], st.lam(s, t));
	}

	double f[E];
	// The two acoustic rows (n = 0, E-1) of L differ only in their first entry and in the entry of the normal momentum: the other
	// E - 2 products U_k l_k / F_k l_k are the same numbers in both rows (-0.5 (b1 u) == 0.5 (-b1 u) exactly), and both rows read the
	// same stencil.  Projected together: one pass of shared-memory loads and E - 2 shared products per stencil value; each row keeps
	// its own k-ascending summation, so every partial sum is the one the reference forms.  (Measured, 512x256x256: sweeps
	// 19.71 / 20.57 / 20.27 -> 18.22 / 19.68 / 19.20 ms with the component-0 hoist below; also carrying the entropy row along made
	// it slower again, 19.74 / 21.00 / 20.58 -- profiles/r01_tuning.md.)
	constexpr bool PAIR = true; // WENO7 (32 accumulators) still fits 128 registers without spilling: -1..2 %
	if constexpr (PAIR)
	{
		const double lA0 = 0.5 * (b2 + un_c + b3), lB0 = 0.5 * (b2 - un_c + b3);
		double ufA[NST], ffA[NST], ufB[NST], ffB[NST];
#pragma unroll
		for (int s = 0; s < NST; s++)
		{
			const double u0 = st.U(s, 0), f0 = st.F(s, 0);
			ufA[s] = u0 * lA0, ffA[s] = f0 * lA0, ufB[s] = u0 * lB0, ffB[s] = f0 * lB0;
		}
#pragma unroll
		for (int k = 1; k < E; k++)
		{
			if (k == 1 + DIR)
			{
				const double lAn = -0.5 * (b1 * un + c1), lBn = 0.5 * (-b1 * un + c1);
#pragma unroll
				for (int s = 0; s < NST; s++)
				{
					const double uk = st.U(s, k), fk = st.F(s, k);
					ufA[s] = ufA[s] + uk * lAn, ffA[s] = ffA[s] + fk * lAn;
					ufB[s] = ufB[s] + uk * lBn, ffB[s] = ffB[s] + fk * lBn;
				}
			}
			else
			{
				const double lk = (k <= 3) ? -0.5 * (b1 * (k == 1 ? R.u : (k == 2 ? R.v : R.w))) : ((k == 4) ? 0.5 * b1 : -0.5 * b1 * R.z[(k - 5) < NC ? (k - 5) : 0]);
#pragma unroll
				for (int s = 0; s < NST; s++)
				{
					const double pu = st.U(s, k) * lk, pf = st.F(s, k) * lk;
					ufA[s] = ufA[s] + pu, ffA[s] = ffA[s] + pf;
					ufB[s] = ufB[s] + pu, ffB[s] = ffB[s] + pf;
				}
			}
		}
		// WENO7: eigen_local == 0 (Eigen_value.hpp:29-32) -> LLF == ROE
		const double avA = (alpha == 1 || (alpha == 2 && WENO == 7)) ? fabs(un - R.c) : ((alpha == 2) ? lmax[0] : glf[0]);
		const double avB = (alpha == 1 || (alpha == 2 && WENO == 7)) ? fabs(un + R.c) : ((alpha == 2) ? lmax[2] : glf[2]);
		if constexpr (WENO == 6)
		{
			f[0] = xf_split_wenocu6(avA, ufA[0], ufA[1], ufA[2], ufA[3], ufA[4], ufA[5], ffA[0], ffA[1], ffA[2], ffA[3], ffA[4], ffA[5], 1.e-8 * dl * dl);
			f[E - 1] = xf_split_wenocu6(avB, ufB[0], ufB[1], ufB[2], ufB[3], ufB[4], ufB[5], ffB[0], ffB[1], ffB[2], ffB[3], ffB[4], ffB[5], 1.e-8 * dl * dl);
		}
		else if constexpr (WENO == 7)
		{
			f[0] = xf_split_weno7(avA, ufA[0], ufA[1], ufA[2], ufA[3], ufA[4], ufA[5], ufA[6], ufA[7], ffA[0], ffA[1], ffA[2], ffA[3], ffA[4], ffA[5], ffA[6], ffA[7]);
			f[E - 1] = xf_split_weno7(avB, ufB[0], ufB[1], ufB[2], ufB[3], ufB[4], ufB[5], ufB[6], ufB[7], ffB[0], ffB[1], ffB[2], ffB[3], ffB[4], ffB[5], ffB[6], ffB[7]);
		}
		else if constexpr (WENO == 5)
		{
			f[0] = xf_split_weno5(avA, ufA[0], ufA[1], ufA[2], ufA[3], ufA[4], ufA[5], ffA[0], ffA[1], ffA[2], ffA[3], ffA[4], ffA[5]);
			f[E - 1] = xf_split_weno5(avB, ufB[0], ufB[1], ufB[2], ufB[3], ufB[4], ufB[5], ffB[0], ffB[1], ffB[2], ffB[3], ffB[4], ffB[5]);
		}
		asm volatile("" ::: "memory");
	}
	// component 0 of the stencil enters every remaining row: read it from shared memory once (WENO5 only: with the longer CU6 / WENO7
	// bodies the 12-16 extra live doubles spill)
	constexpr bool HOIST = (WENO == 5);
	double U0s[NST], F0s[NST];
	if constexpr (HOIST)
	{
#pragma unroll
		for (int s = 0; s < NST; s++)
			U0s[s] = st.U(s, 0), F0s[s] = st.F(s, 0);
	}
#define XF_U0(s) (HOIST ? U0s[s] : st.U(s, 0))
#define XF_F0(s) (HOIST ? F0s[s] : st.F(s, 0))
#pragma unroll
	for (int n = 0; n < E; n++)
	{
		if (PAIR && (n == 0 || n == E - 1))
			continue;
		// ---- artificial viscosity for this field (Eigen_callback.h:134-145, 188-193) ----
		const int t = (n == 0) ? 0 : ((n == E - 1) ? 2 : 1);
		const double ev = (n == 0) ? fabs(un - R.c) : ((n == E - 1) ? fabs(un + R.c) : fabs(un));
		// WENO7: eigen_local == 0 for SCHEME_ORDER 7 (Eigen_value.hpp:29-32), the sign test never fires -> LLF == ROE
		const double av = (alpha == 1 || (alpha == 2 && WENO == 7)) ? ev : ((alpha == 2) ? lmax[t] : glf[t]);

		// ---- project the stencil on row n of L ----
		double uf[NST], ff[NST];
		if (n == 0 || n == E - 1 || n == ENT)
		{
			double l[E];
			if (n == ENT)
			{
				l[0] = (1.0 - b2 - b3) / b1;
				l[1] = R.u, l[2] = R.v, l[3] = R.w, l[4] = -1.0;
#pragma unroll
				for (int m = 0; m < NC; m++)
					l[5 + m] = R.z[m];
			}
			else if (n == 0)
			{
				l[0] = 0.5 * (b2 + un_c + b3);
				l[1] = DIR == 0 ? -0.5 * (b1 * R.u + c1) : -0.5 * (b1 * R.u);
				l[2] = DIR == 1 ? -0.5 * (b1 * R.v + c1) : -0.5 * (b1 * R.v);
				l[3] = DIR == 2 ? -0.5 * (b1 * R.w + c1) : -0.5 * (b1 * R.w);
				l[4] = 0.5 * b1;
#pragma unroll
				for (int m = 0; m < NC; m++)
					l[5 + m] = -0.5 * b1 * R.z[m];
			}
			else
			{
				l[0] = 0.5 * (b2 - un_c + b3);
				l[1] = DIR == 0 ? 0.5 * (-b1 * R.u + c1) : 0.5 * (-b1 * R.u);
				l[2] = DIR == 1 ? 0.5 * (-b1 * R.v + c1) : 0.5 * (-b1 * R.v);
				l[3] = DIR == 2 ? 0.5 * (-b1 * R.w + c1) : 0.5 * (-b1 * R.w);
				l[4] = 0.5 * b1;
#pragma unroll
				for (int m = 0; m < NC; m++)
					l[5 + m] = -0.5 * b1 * R.z[m];
			}
#pragma unroll
			for (int s = 0; s < NST; s++)
			{
				uf[s] = XF_U0(s) * l[0];
				ff[s] = XF_F0(s) * l[0];
			}
#pragma unroll
			for (int k = 1; k < E; k++)
			{
				if (n == ENT && k == 4)
				{ // * (-1.0)
#pragma unroll
					for (int s = 0; s < NST; s++)
					{
						uf[s] = uf[s] - st.U(s, 4);
						ff[s] = ff[s] - st.F(s, 4);
					}
				}
				else
				{
#pragma unroll
					for (int s = 0; s < NST; s++)
					{
						uf[s] = uf[s] + st.U(s, k) * l[k];
						ff[s] = ff[s] + st.F(s, k) * l[k];
					}
				}
			}
		}
		else if (n >= 1 && n <= 3)
		{ // shear rows: two non-zero entries, one of them +-1 (Eigen_matrix.hpp:38-59,177-209,327-348)
			// x: n=2 (v,-e2) n=3 (-w,+e3);  y: n=1 (-u,+e1) n=3 (w,-e3);  z: n=1 (u,-e1) n=2 (-v,+e2)
			const double vel = (n == 1) ? R.u : ((n == 2) ? R.v : R.w);
			const bool plus_unit = (DIR == 0) ? (n == 3) : ((DIR == 1) ? (n == 1) : (n == 2));
#pragma unroll
			for (int s = 0; s < NST; s++)
			{
				if (plus_unit)
				{ // U0*(-vel) + Un
					uf[s] = XF_U0(s) * (-vel) + st.U(s, n);
					ff[s] = XF_F0(s) * (-vel) + st.F(s, n);
				}
				else
				{ // U0*vel - Un
					uf[s] = XF_U0(s) * vel - st.U(s, n);
					ff[s] = XF_F0(s) * vel - st.F(s, n);
				}
			}
		}
		else
		{ // species rows: (-Y_s, 0,0,0,0, e_s)
			const double ys = R.y[(n - 4) < NC ? (n - 4) : 0];
#pragma unroll
			for (int s = 0; s < NST; s++)
			{
				uf[s] = XF_U0(s) * (-ys) + st.U(s, n + 1);
				ff[s] = XF_F0(s) * (-ys) + st.F(s, n + 1);
			}
		}

		// ---- Lax-Friedrichs splitting + WENO (one out-of-line copy: keeps the kernel inside the instruction cache) ----
		if constexpr (WENO == 7)
			f[n] = xf_split_weno7(av, uf[0], uf[1], uf[2], uf[3], uf[4], uf[5], uf[6], uf[7], ff[0], ff[1], ff[2], ff[3], ff[4], ff[5], ff[6], ff[7]);
		else if constexpr (WENO == 6)
			f[n] = xf_split_wenocu6(av, uf[0], uf[1], uf[2], uf[3], uf[4], uf[5], ff[0], ff[1], ff[2], ff[3], ff[4], ff[5], 1.e-8 * dl * dl);
		else
			f[n] = xf_split_weno5(av, uf[0], uf[1], uf[2], uf[3], uf[4], uf[5], ff[0], ff[1], ff[2], ff[3], ff[4], ff[5]);
		// compiler fence: forbid keeping stencil values loaded for this field alive into the next one
		// (without it the unrolled field loop is CSE'd into ~2*E*NST live doubles and spills)
		asm volatile("" ::: "memory");
	}

#undef XF_U0
#undef XF_F0
	// ---- back-projection Fw[k] = sum_n f[n] * R[n][k], n ascending, from 0.0 (Eigen_callback.h:222-230) ----
#pragma unroll
	for (int k = 0; k < E; k++)
		Fw[k] = 0.0;
#pragma unroll
	for (int n = 0; n < E; n++)
	{
		const double fn = f[n];
		if (n == 0 || n == E - 1)
		{
			const double sg = (n == 0) ? -1.0 : 1.0; // u -+ c
			Fw[0] = Fw[0] + fn;
			Fw[1] = Fw[1] + fn * (DIR == 0 ? (n == 0 ? R.u - R.c : R.u + R.c) : R.u);
			Fw[2] = Fw[2] + fn * (DIR == 1 ? (n == 0 ? R.v - R.c : R.v + R.c) : R.v);
			Fw[3] = Fw[3] + fn * (DIR == 2 ? (n == 0 ? R.w - R.c : R.w + R.c) : R.w);
			Fw[4] = Fw[4] + fn * (n == 0 ? R.H - un * R.c : R.H + un * R.c);
			(void)sg;
#pragma unroll
			for (int m = 0; m < NC; m++)
				Fw[5 + m] = Fw[5 + m] + fn * R.y[m];
		}
		else if (n == ENT)
		{
			Fw[0] = Fw[0] + fn * b1;
			Fw[1] = Fw[1] + fn * (R.u * b1);
			Fw[2] = Fw[2] + fn * (R.v * b1);
			Fw[3] = Fw[3] + fn * (R.w * b1);
			Fw[4] = Fw[4] + fn * (R.H * b1 - 1.0);
#pragma unroll
			for (int m = 0; m < NC; m++)
				Fw[5 + m] = Fw[5 + m] + fn * (b1 * R.y[m]);
		}
		else if (n >= 1 && n <= 3)
		{ // x: n=2 (0,0,-1,0,-v) n=3 (0,0,0,1,w); y: n=1 (0,1,0,0,u) n=3 (0,0,0,-1,-w); z: n=1 (0,-1,0,0,-u) n=2 (0,0,1,0,v)
			const double vel = (n == 1) ? R.u : ((n == 2) ? R.v : R.w);
			const bool plus_unit = (DIR == 0) ? (n == 3) : ((DIR == 1) ? (n == 1) : (n == 2));
			if (plus_unit)
			{
				Fw[n] = Fw[n] + fn;
				Fw[4] = Fw[4] + fn * vel;
			}
			else
			{
				Fw[n] = Fw[n] - fn;
				Fw[4] = Fw[4] + fn * (-vel);
			}
		}
		else
		{ // species column: (0,0,0,0,z_s, e_s)
			const int s = (n - 4) < NC ? (n - 4) : 0;
			Fw[4] = Fw[4] + fn * R.z[s];
			Fw[n + 1] = Fw[n + 1] + fn;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// Positivity-preserving flux limiter (PositivityPreservingKernel, PositivityPreserving_kernels.hpp:5-76) applied to the
// wall flux this thread has just formed, from the staged conserved variables / physical fluxes of the face's two cells
// (stencil slots P and P+1).  lambda_0 = uvw_c_max[dir] of the last GetDt, lambda = CFL / lambda_0.  epsilon = {1e-13 (rho),
// 1e-13, 0 (every Y_i)} (ConVenction_block.hpp:334-338).  The `FF[n] = ...` statement inside the reference's species loop
// (line 71) only rewrites entries that are never read again; it is dead and omitted.
// ------------------------------------------------------------------------------------------------
template <class C, int WENO, class ST>
XF_DEV void xf_positivity(const ST &st, double lambda_0, double CFL, double *Fw /*E, in/out*/)
{
	constexpr int E = C::E, NC = C::NC, P = XfStencil<WENO>::P;
	const double lambda = CFL / lambda_0;
	const double tl = 2.0 * lambda;
	double F_LF[E];
#pragma unroll
	for (int n = 0; n < E; n++)
		F_LF[n] = 0.5 * (2.0 * (st.F(P, n) + st.F(P + 1, n)) + lambda_0 * (st.U(P, n) - st.U(P + 1, n))); // st.F is 0.5 F: 2 (hFl + hFr) == Fl + Fr exactly
	const double UU0 = st.U(P, 0), UP0 = st.U(P + 1, 0);
	const double FF_LF0 = tl * F_LF[0];
	double FF0 = tl * Fw[0];
	double theta_u = 1.0, theta_p = 1.0;
	double rho_min = xf_min(UU0, 1.0e-13);
	if (UU0 - FF0 < rho_min)
		theta_u = (UU0 - FF_LF0 - rho_min + 1.0e-40) / (FF0 - FF_LF0 + 1.0e-40);
	rho_min = xf_min(UP0, 1.0e-13);
	if (UP0 + FF0 < rho_min)
		theta_p = (UP0 + FF_LF0 - rho_min + 1.0e-40) / (FF_LF0 - FF0 + 1.0e-40);
	double theta = xf_min(xf_max(xf_min(theta_u, theta_p), 0.0), 1.0);
	// FF of the species rows, limited with the density theta (the only FF entries read below)
	double FFs[NC > 0 ? NC : 1], FF_LFs[NC > 0 ? NC : 1];
#pragma unroll
	for (int n = 0; n < NC; n++)
	{
		FF_LFs[n] = tl * F_LF[5 + n];
		FFs[n] = (1.0 - theta) * FF_LFs[n] + theta * (tl * Fw[5 + n]);
	}
	FF0 = (1.0 - theta) * FF_LF0 + theta * FF0;
#pragma unroll
	for (int n = 0; n < E; n++)
		Fw[n] = (1.0 - theta) * F_LF[n] + theta * Fw[n];
	if constexpr (NC > 0)
	{
		const double _rhoq = 1.0 / (UU0 - FF0), _rhou = 1.0 / (UU0 - FF_LF0);
		const double _rhoqp = 1.0 / (UP0 + FF0), _rhoup = 1.0 / (UP0 + FF_LF0);
#pragma unroll
		for (int n = 0; n < NC; n++)
		{
			const double Us = st.U(P, 5 + n), Ups = st.U(P + 1, 5 + n);
			const double yi_q = (Us - FFs[n]) * _rhoq, yi_u = (Us - FF_LFs[n]) * _rhou;
			const double yi_qp = (Ups + FFs[n]) * _rhoqp, yi_up = (Ups + FF_LFs[n]) * _rhoup;
			theta_u = 1.0, theta_p = 1.0;
			if (yi_q < 0.0)
			{
				const double yi_min = xf_min(yi_u, 0.0);
				theta_u = (yi_u - yi_min + 1.0e-40) / (yi_u - yi_q + 1.0e-40);
			}
			if (yi_qp < 0.0)
			{
				const double yi_min = xf_min(yi_up, 0.0);
				theta_p = (yi_up - yi_min + 1.0e-40) / (yi_up - yi_qp + 1.0e-40);
			}
			theta = xf_min(xf_max(xf_min(theta_u, theta_p), 0.0), 1.0);
#pragma unroll
			for (int nn = 0; nn < E; nn++)
				Fw[nn] = (1.0 - theta) * F_LF[nn] + theta * Fw[nn];
		}
	}
}
