// check_xf_log.cpp -- host harness: xf_log (xfluids_b200/csrc/xf_log.cuh, host restatement of the device sequence) against the
// log() of this machine's libm, bit for bit.  usage: check_xf_log [n_uniform] ; prints "mismatches=<m> of <n>".
//   set 1: n_uniform equidistant arguments of [200, 6000]  (the path's range: T = max(T, 200), NASA-9 fits end at 6000 K)
//   set 2: n_uniform / 4 random bit patterns of positive normal doubles (exponents -1000 .. 1000)
//   set 3: every double within 2^16 ulps of 200, 1000, 6000 and of the powers of two in between
// build: g++ -O2 -std=c++17 -fopenmp -mfma -ffp-contract=off tools/check_xf_log.cpp -o /tmp/check_xf_log -lm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cstdint>
#include "../xfluids_b200/csrc/xf_log.cuh"

static inline bool same(double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a != a && b != b); }
int main(int argc, char **argv)
{
	const long long n = argc > 1 ? atoll(argv[1]) : 120000000LL;
	long long bad = 0, total = 0;
	double (*volatile libm_log)(double) = std::log; // no constant folding / builtin expansion
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n; i++)
	{
		const double x = 200.0 + 5800.0 * ((double)i + 0.5) / (double)n;
		bad += !same(xf_log(x), libm_log(x));
	}
	total += n;
#pragma omp parallel for reduction(+ : bad)
	for (long long i = 0; i < n / 4; i++)
	{
		uint64_t s = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
		s ^= s >> 31, s *= 0xBF58476D1CE4E5B9ull, s ^= s >> 29;
		const uint64_t e = 23 + (s >> 53) % 2000; // biased exponent 23 .. 2022
		const uint64_t bits = (e << 52) | (s & 0xFFFFFFFFFFFFFull);
		double x;
		std::memcpy(&x, &bits, 8);
		bad += !same(xf_log(x), libm_log(x));
	}
	total += n / 4;
	const double pts[] = {200.0, 256.0, 512.0, 1000.0, 1024.0, 2048.0, 4096.0, 6000.0, 1.0, 0.9375, 1.064697265625, 2.0, 0.5};
	for (double c : pts)
	{
		uint64_t b;
		std::memcpy(&b, &c, 8);
		for (long long d = -65536; d <= 65536; d++)
		{
			const uint64_t bb = b + (uint64_t)d;
			double x;
			std::memcpy(&x, &bb, 8);
			bad += !same(xf_log(x), libm_log(x));
			total++;
		}
	}
	printf("mismatches=%lld of %lld\n", bad, total);
	return bad ? 1 : 0;
}
