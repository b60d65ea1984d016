// check_xf_exp.cpp -- host harness: xf_exp / xf_pow (xfluids_b200/csrc/xf_exp.cuh, host restatement of the device sequences) against the
// exp() / pow() of this machine's libm, bit for bit.  usage: check_xf_exp [n] ; prints "mismatches=<m> of <n>".
//   exp: n equidistant arguments of [-40, 10] (the fits: ln mu ~ -12, ln lambda ~ -3, ln(p D) ~ 0..3) and n / 4 random ones of [-500, 500] (the table-driven path: |x| < 512)
//   pow: n arguments x of [0.02, 50] log-spaced with y = 0.5; n / 4 with y = 0.25 and -0.5; n / 4 random (x in [1e-3, 1e3], y in [-3, 3])
// build: g++ -O2 -std=c++17 -fopenmp -mfma -ffp-contract=off tools/check_xf_exp.cpp -o /tmp/check_xf_exp -lm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include "../xfluids_b200/csrc/xf_exp.cuh"

static inline bool same(double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a != a && b != b); }
static inline double u01(uint64_t i)
{
	uint64_t s = 0x9E3779B97F4A7C15ull * (i + 1);
	s ^= s >> 31, s *= 0xBF58476D1CE4E5B9ull, s ^= s >> 29;
	return (double)(s >> 11) / 9007199254740992.0;
}
int main(int argc, char **argv)
{
	const long long n = argc > 1 ? atoll(argv[1]) : 100000000LL;
	long long bad = 0, total = 0;
	double (*volatile libm_exp)(double) = std::exp; // no constant folding / builtin expansion
	double (*volatile libm_pow)(double, double) = std::pow;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n; i++)
	{
		const double x = -40.0 + 50.0 * ((double)i + 0.5) / (double)n;
		bad += !same(xf_exp(x), libm_exp(x));
	}
	total += n;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n / 4; i++)
	{
		const double x = -500.0 + 1000.0 * u01((uint64_t)i);
		bad += !same(xf_exp(x), libm_exp(x));
	}
	total += n / 4;
	const long long bad_exp = bad;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n; i++)
	{
		const double x = 0.02 * libm_exp(7.824046010856292 * ((double)i + 0.5) / (double)n); // 0.02 .. 50
		bad += !same(xf_pow(x, 0.5), libm_pow(x, 0.5));
	}
	total += n;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n / 4; i++)
	{
		const double x = 0.02 + 49.98 * u01((uint64_t)i);
		bad += !same(xf_pow(x, 0.25), libm_pow(x, 0.25));
		bad += !same(xf_pow(x, -0.5), libm_pow(x, -0.5));
	}
	total += n / 2;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n / 4; i++)
	{
		const double x = libm_exp(-6.9 + 13.8 * u01((uint64_t)(2 * i))), y = -3.0 + 6.0 * u01((uint64_t)(2 * i + 1));
		bad += !same(xf_pow(x, y), libm_pow(x, y));
	}
	total += n / 4;
	printf("mismatches=%lld of %lld (exp: %lld)\n", bad, total, bad_exp);
	return bad ? 1 : 0;
}
