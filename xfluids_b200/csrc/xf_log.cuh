// xf_log.cuh -- double-precision natural logarithm, bit-identical to the log() of the libm the reference CPU path links against
// (glibc >= 2.28: sysdeps/ieee754/dbl-64/e_log.c, the __FP_FAST_FMA code path that the x86-64 ifunc selects on every FMA-capable CPU).
//
// log is the only operation of the inviscid path that is not an IEEE-754 basic operation (get_Enthalpy_NASA, reference
// src/solver_Ini/Thermo_device.h:62-80); with it pinned, the strict build performs the same sequence of correctly-rounded operations
// as the reference and multi-species results agree bit for bit (round 1 used libdevice's log: <= 1 ulp apart on some arguments).
//
// The operation sequence below is the one the shipped binary executes (checked against the disassembly of __log_fma in glibc 2.39:
// which products are fused is the compiler's choice under -ffp-contract=fast, so the C source alone does not determine the bits):
//   tmp = ix - OFF; i = (tmp >> 45) & 127; k = tmp >> 52; z = x * 2^-k; invc, logc = T[i]
//   r  = fma(z, invc, -1)                 kd = (double) k
//   w  = fma(kd, Ln2hi, logc)             hi = w + r
//   lo = fma(kd, Ln2lo, (w - hi) + r)     r2 = r * r
//   p  = fma(fma(r, A4, A3), r2, fma(r, A2, A1))
//   y  = fma(r * r2, p, fma(r2, A0, lo)) + hi
// Arguments outside the main path (x within [1 - 2^-4, 1 + 0x1.09p-4), subnormal, <= 0, inf, nan) take the platform's log(); the
// path's only call site passes max(T, 200).  tests/test_xf_log.py: > 1e8 arguments of [200, 6000] (and wider), 0 mismatches on
// the CPU and on the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include "xf_log_data.h"

#ifdef __CUDACC__
#define XF_LOG_HD __host__ __device__ __forceinline__
static __device__ const double2 xf_log_tab_dev[128] = {XF_LOG_TABLE};
#else
#define XF_LOG_HD static inline
#endif
static const double xf_log_tab_host[128][2] = {XF_LOG_TABLE};

XF_LOG_HD double xf_log(double x)
{
	uint64_t ix;
#ifdef __CUDA_ARCH__
	ix = (uint64_t)__double_as_longlong(x);
#else
	std::memcpy(&ix, &x, 8);
#endif
	const uint64_t LO = 0x3fee000000000000ull, HI = 0x3ff1090000000000ull; // 1 - 2^-4, 1 + 0x1.09p-4
	const uint32_t top = (uint32_t)(ix >> 48);
	if (ix - LO < HI - LO || top - 0x0010u >= 0x7ff0u - 0x0010u)
		return log(x);
	const uint64_t tmp = ix - 0x3fe6000000000000ull;
	const int i = (int)((tmp >> 45) & 127);
	const int k = (int)((int64_t)tmp >> 52);
	const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
	double z, invc, logc;
#ifdef __CUDA_ARCH__
	z = __longlong_as_double((long long)iz);
	const double2 t = __ldg(&xf_log_tab_dev[i]);
	invc = t.x, logc = t.y;
	const double kd = (double)k;
	const double r = fma(z, invc, -1.0);
	const double w = fma(kd, XF_LOG_LN2HI, logc);
	const double hi = __dadd_rn(w, r);
	const double lo = fma(kd, XF_LOG_LN2LO, __dadd_rn(__dsub_rn(w, hi), r));
	const double r2 = __dmul_rn(r, r);
	const double p = fma(fma(r, XF_LOG_A4, XF_LOG_A3), r2, fma(r, XF_LOG_A2, XF_LOG_A1));
	return __dadd_rn(fma(__dmul_rn(r, r2), p, fma(r2, XF_LOG_A0, lo)), hi);
#else
	std::memcpy(&z, &iz, 8);
	invc = xf_log_tab_host[i][0], logc = xf_log_tab_host[i][1];
	// host restatement (tests only): volatile stores keep the compiler from contracting the unfused operations
	const double kd = (double)k;
	const double r = std::fma(z, invc, -1.0);
	const double w = std::fma(kd, XF_LOG_LN2HI, logc);
	volatile double hi = w + r;
	volatile double d0 = w - hi;
	volatile double d1 = d0 + r;
	const double lo = std::fma(kd, XF_LOG_LN2LO, d1);
	volatile double r2 = r * r;
	volatile double r3 = r * r2;
	const double p = std::fma(std::fma(r, XF_LOG_A4, XF_LOG_A3), r2, std::fma(r, XF_LOG_A2, XF_LOG_A1));
	volatile double y = std::fma(r3, p, std::fma(r2, XF_LOG_A0, lo));
	return y + hi;
#endif
}
