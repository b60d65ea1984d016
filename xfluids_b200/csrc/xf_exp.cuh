// xf_exp.cuh -- double-precision exp and pow, bit-identical to the exp() / pow() of the libm the reference CPU path links against
// (glibc >= 2.28: sysdeps/ieee754/dbl-64/e_exp.c and e_pow.c, the FMA code paths the x86-64 ifunc selects on every FMA-capable CPU).
//
// The viscous block evaluates its transport fits as exp(cubic in ln T) and the Wilke factor PHI with pow(mu_k / mu_i, 0.5) (reference
// src/solver_Reconstruction/viscosity/Visc_device.h:10-41).  With log (xf_log.cuh), exp and pow pinned, the strict build performs the same
// sequence of correctly rounded operations as the reference there too.
//
// The operation sequences are the ones the shipped binary executes (read off the disassembly of __exp_fma / __pow_fma in glibc 2.39; which
// products are fused is the compiler's choice, so the C source alone does not determine the bits):
//   exp:  kd0 = fma(x, InvLn2N, Shift); ki = bits(kd0); kd = kd0 - Shift
//         r = fma(kd, NegLn2loN, fma(kd, NegLn2hiN, x));  {tail, sbits} = T[ki % 128], sbits += ki << 45
//         tmp = fma(r2 * r2, fma(r, C5, C4), fma(fma(r, C3, C2), r2, tail + r)), r2 = r * r;   y = fma(scale, tmp, scale)
//   pow:  log_inline: r = fma(z, invc, -1); t1 = fma(kd, Ln2hi, logc); t2 = t1 + r; lo1 = fma(kd, Ln2lo, logctail); lo2 = (t1 - t2) + r
//         ar = A0 r; ar2 = r ar; ar3 = r ar2; hi = t2 + ar2; lo3 = fma(ar, r, -ar2); lo4 = (t2 - hi) + ar2
//         p = fma(ar2, fma(fma(r, A6, A5), ar2, fma(r, A4, A3)), fma(r, A2, A1));  lo = fma(ar3, p, ((lo1 + lo2) + lo3) + lo4)
//         lhi = hi + lo; ltail = (hi - lhi) + lo;   ehi = y lhi; elo = fma(y, ltail, fma(lhi, y, -ehi))
//         exp_inline(ehi, elo): as exp with r += elo after the reduction
// Arguments outside the main paths (exp: |x| < 2^-54 or >= 512; pow: x not a positive normal, |y| outside [2^-65, 2^63), result near
// over/underflow) take the platform's function; the path's call sites are far inside.  tools/check_xf_exp.cpp and tests/test_xf_log.py:
// > 1e8 arguments each, 0 mismatches on the CPU and on the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "xf_exp_data.h"

#ifdef __CUDACC__
#define XF_EXP_HD __host__ __device__ __forceinline__
static __device__ const ulonglong2 xf_exp_tab_dev[128] = {XF_EXP_TABLE};
struct XfPowLogEntry
{
	double invc, logc, logctail;
};
static __device__ const XfPowLogEntry xf_powlog_tab_dev[128] = {XF_POWLOG_TABLE};
#else
#define XF_EXP_HD static inline
#endif
static const uint64_t xf_exp_tab_host[128][2] = {XF_EXP_TABLE};
static const double xf_powlog_tab_host[128][3] = {XF_POWLOG_TABLE};

#ifdef __CUDA_ARCH__
#define XF_BITS(x) ((uint64_t)__double_as_longlong(x))
#define XF_DBL(b) __longlong_as_double((long long)(b))
#define XF_ADD(a, b) __dadd_rn(a, b)
#define XF_SUB(a, b) __dsub_rn(a, b)
#define XF_MUL(a, b) __dmul_rn(a, b)
#define XF_FMA(a, b, c) fma(a, b, c)
#else
static inline uint64_t xf_bits_host(double x)
{
	uint64_t b;
	std::memcpy(&b, &x, 8);
	return b;
}
static inline double xf_dbl_host(uint64_t b)
{
	double x;
	std::memcpy(&x, &b, 8);
	return x;
}
// host restatement (tests only): volatile temporaries keep the compiler from contracting the unfused operations
static inline double xf_add_host(double a, double b) { volatile double r = a + b; return r; }
static inline double xf_sub_host(double a, double b) { volatile double r = a - b; return r; }
static inline double xf_mul_host(double a, double b) { volatile double r = a * b; return r; }
#define XF_BITS(x) xf_bits_host(x)
#define XF_DBL(b) xf_dbl_host(b)
#define XF_ADD(a, b) xf_add_host(a, b)
#define XF_SUB(a, b) xf_sub_host(a, b)
#define XF_MUL(a, b) xf_mul_host(a, b)
#define XF_FMA(a, b, c) std::fma(a, b, c)
#endif

// the common tail of exp and pow: exp(x + xtail) * 1, x already range-checked
XF_EXP_HD double xf_exp_core(double x, double xtail, bool with_tail)
{
	const double kd0 = XF_FMA(x, XF_EXP_INVLN2N, XF_EXP_SHIFT);
	const uint64_t ki = XF_BITS(kd0);
	const double kd = XF_SUB(kd0, XF_EXP_SHIFT);
	double r = XF_FMA(kd, XF_EXP_NEGLN2LON, XF_FMA(kd, XF_EXP_NEGLN2HIN, x));
	if (with_tail)
		r = XF_ADD(xtail, r);
	const int idx = (int)(ki & 127);
	uint64_t tbits, sbits;
#ifdef __CUDA_ARCH__
	const ulonglong2 t = __ldg(&xf_exp_tab_dev[idx]);
	tbits = t.x, sbits = t.y;
#else
	tbits = xf_exp_tab_host[idx][0], sbits = xf_exp_tab_host[idx][1];
#endif
	sbits += ki << 45;
	const double tail = XF_DBL(tbits), scale = XF_DBL(sbits);
	const double p23 = XF_FMA(r, XF_EXP_C3, XF_EXP_C2);
	const double t1 = XF_ADD(r, tail);
	const double r2 = XF_MUL(r, r);
	const double p45 = XF_FMA(r, XF_EXP_C5, XF_EXP_C4);
	const double q = XF_FMA(p23, r2, t1);
	const double r4 = XF_MUL(r2, r2);
	const double tmp = XF_FMA(r4, p45, q);
	return XF_FMA(scale, tmp, scale);
}

XF_EXP_HD double xf_exp(double x)
{
	const uint32_t abstop = (uint32_t)(XF_BITS(x) >> 52) & 0x7ff;
	if (abstop - 0x3c9u > 0x3eu) // |x| < 2^-54 or |x| >= 512 or non-finite
		return exp(x);
	return xf_exp_core(x, 0.0, false);
}

XF_EXP_HD double xf_pow(double x, double y)
{
	const uint64_t ix = XF_BITS(x), iy = XF_BITS(y);
	const uint32_t topx = (uint32_t)(ix >> 52), topy = (uint32_t)(iy >> 52);
	if (topx - 1u > 0x7fdu || (topy & 0x7ff) - 0x3beu > 0x7fu) // x not a positive normal number, or |y| outside [2^-65, 2^63)
		return pow(x, y);
	// log_inline
	const uint64_t tmp = ix - 0x3fe6955500000000ull;
	const int i = (int)((tmp >> 45) & 127);
	const int k = (int)((int64_t)tmp >> 52);
	const double z = XF_DBL(ix - (tmp & 0xfff0000000000000ull));
	const double kd = (double)k;
	double invc, logc, logctail;
#ifdef __CUDA_ARCH__
	const XfPowLogEntry &e = xf_powlog_tab_dev[i];
	invc = __ldg(&e.invc), logc = __ldg(&e.logc), logctail = __ldg(&e.logctail);
#else
	invc = xf_powlog_tab_host[i][0], logc = xf_powlog_tab_host[i][1], logctail = xf_powlog_tab_host[i][2];
#endif
	const double r = XF_FMA(z, invc, -1.0);
	const double t1 = XF_FMA(kd, XF_POWLOG_LN2HI, logc);
	const double t2 = XF_ADD(t1, r);
	const double lo1 = XF_FMA(kd, XF_POWLOG_LN2LO, logctail);
	const double lo2 = XF_ADD(XF_SUB(t1, t2), r);
	const double ar = XF_MUL(r, XF_POWLOG_A0);
	const double ar2 = XF_MUL(r, ar);
	const double ar3 = XF_MUL(r, ar2);
	const double hi = XF_ADD(t2, ar2);
	const double lo3 = XF_FMA(ar, r, -ar2);
	const double lo4 = XF_ADD(XF_SUB(t2, hi), ar2);
	const double p = XF_FMA(ar2, XF_FMA(XF_FMA(r, XF_POWLOG_A6, XF_POWLOG_A5), ar2, XF_FMA(r, XF_POWLOG_A4, XF_POWLOG_A3)), XF_FMA(r, XF_POWLOG_A2, XF_POWLOG_A1));
	const double lo = XF_FMA(ar3, p, XF_ADD(XF_ADD(XF_ADD(lo1, lo2), lo3), lo4));
	const double lhi = XF_ADD(hi, lo);
	const double ltail = XF_ADD(XF_SUB(hi, lhi), lo);
	const double ehi = XF_MUL(y, lhi);
	const double elo = XF_FMA(y, ltail, XF_FMA(lhi, y, -ehi));
	const uint32_t abstop = (uint32_t)(XF_BITS(ehi) >> 52) & 0x7ff;
	if (abstop - 0x3c9u > 0x3eu) // result within rounding of 1, or near over / underflow
		return pow(x, y);
	return xf_exp_core(ehi, elo, true);
}
