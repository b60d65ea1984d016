// oracle/shim/sycl/sycl.hpp -- TEST INFRASTRUCTURE, not product code.
//
// A serial / OpenMP *host* stand-in for the small SYCL 2020 subset the XFluids reference
// sources use (queue/handler/parallel_for(nd_range<1|2|3>)/reduction/USM malloc/math).
// It exists so that the UNMODIFIED reference .cpp files under /root/reference can be compiled
// with plain g++ into oracle/_ref/ and executed as the parity oracle and CPU baseline
// (SURVEY.md 8c, Appendix E).  Every parallel_for runs its body once per global index:
// k (dim 2) outermost ... i (dim 0) innermost, which is the reference's memory order.
// With -fopenmp the two outer loops are work-shared, which is the execution model of
// AdaptiveCpp's omp backend (one work-group per thread); results do not depend on it because
// the reference kernels have no cross-item accumulation except exact max/min reductions.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <functional>
#include <iostream>
#include <type_traits>
#include <utility>
#include <tuple>
#include <memory>

#define SYCL_EXTERNAL

namespace sycl
{
	template <int N>
	struct range
	{
		size_t v[N];
		range() { for (int d = 0; d < N; d++) v[d] = 1; }
		template <typename... A, typename = std::enable_if_t<sizeof...(A) == N>>
		range(A... a) : v{size_t(a)...} {}
		size_t &operator[](int d) { return v[d]; }
		const size_t &operator[](int d) const { return v[d]; }
		bool operator==(const range &o) const
		{
			for (int d = 0; d < N; d++)
				if (v[d] != o.v[d])
					return false;
			return true;
		}
		size_t size() const
		{
			size_t s = 1;
			for (int d = 0; d < N; d++)
				s *= v[d];
			return s;
		}
	};
	template <typename... A>
	range(A...) -> range<sizeof...(A)>;
	template <int N>
	using id = range<N>;

	template <int N>
	struct nd_range
	{
		range<N> g, l;
		nd_range(range<N> G, range<N> L) : g(G), l(L) {}
		range<N> get_global_range() const { return g; }
		range<N> get_local_range() const { return l; }
	};

	template <int N>
	struct nd_item
	{
		size_t gid[N];
		size_t get_global_id(int d) const { return gid[d]; }
		size_t get_global_id() const { return gid[0]; }
		size_t get_global_linear_id() const { return gid[0]; }
	};

	struct event
	{
		void wait() {}
	};

	template <typename T = void>
	struct plus
	{
		template <typename U>
		U operator()(U a, U b) const { return a + b; }
	};
	template <typename T = void>
	struct maximum
	{
		template <typename U>
		U operator()(U a, U b) const { return a < b ? b : a; }
	};
	template <typename T = void>
	struct minimum
	{
		template <typename U>
		U operator()(U a, U b) const { return b < a ? b : a; }
	};

	// reduction descriptor: combines INTO the current *p (SYCL default without
	// initialize_to_identity) -- the reference's GLF running max relies on this.
	template <typename T, typename Op>
	struct reduction_desc
	{
		T *p;
		Op op;
	};
	template <typename T, typename Op>
	struct reducer
	{
		T val;
		bool touched;
		Op op;
		void combine(T x)
		{
			val = touched ? op(val, x) : x;
			touched = true;
		}
		reducer &operator+=(T x)
		{
			combine(x);
			return *this;
		}
	};
	template <typename T, typename Op>
	reduction_desc<T, Op> reduction(T *p, Op op) { return {p, op}; }

	namespace detail
	{
		template <typename K>
		inline void run(const nd_range<1> &r, K &k)
		{
			const long n0 = long(r.g[0]);
#pragma omp parallel for schedule(static)
			for (long i = 0; i < n0; i++)
			{
				nd_item<1> it{{size_t(i)}};
				k(it);
			}
		}
		template <typename K>
		inline void run(const nd_range<2> &r, K &k)
		{
			const long n0 = long(r.g[0]), n1 = long(r.g[1]);
#pragma omp parallel for schedule(static)
			for (long j = 0; j < n1; j++)
				for (long i = 0; i < n0; i++)
				{
					nd_item<2> it{{size_t(i), size_t(j)}};
					k(it);
				}
		}
		template <typename K>
		inline void run(const nd_range<3> &r, K &k)
		{
			const long n0 = long(r.g[0]), n1 = long(r.g[1]), n2 = long(r.g[2]);
#pragma omp parallel for collapse(2) schedule(static)
			for (long kk = 0; kk < n2; kk++)
				for (long j = 0; j < n1; j++)
					for (long i = 0; i < n0; i++)
					{
						nd_item<3> it{{size_t(i), size_t(j), size_t(kk)}};
						k(it);
					}
		}

		// one reduction object: per-thread partials, combined into *p at the end
		template <int N, typename T, typename Op, typename K>
		inline void run_red(const nd_range<N> &r, reduction_desc<T, Op> d, K &k)
		{
			const long n0 = long(r.g[0]), n1 = N > 1 ? long(r.g[N > 1 ? 1 : 0]) : 1, n2 = N > 2 ? long(r.g[N > 2 ? 2 : 0]) : 1;
#pragma omp parallel
			{
				reducer<T, Op> red{T(), false, d.op};
#pragma omp for collapse(2) schedule(static) nowait
				for (long kk = 0; kk < n2; kk++)
					for (long j = 0; j < n1; j++)
						for (long i = 0; i < n0; i++)
						{
							nd_item<N> it;
							it.gid[0] = size_t(i);
							if constexpr (N > 1)
								it.gid[1] = size_t(j);
							if constexpr (N > 2)
								it.gid[2] = size_t(kk);
							k(it, red);
						}
#pragma omp critical
				{
					if (red.touched)
						*d.p = d.op(*d.p, red.val);
				}
			}
		}
	} // namespace detail

	struct handler
	{
		void depends_on(event) {}
		template <typename E>
		void depends_on(const std::vector<E> &) {}
		// trailing argument = kernel, preceding arguments = reduction descriptors
		template <int N, typename... A>
		void parallel_for(nd_range<N> r, A &&...a)
		{
			auto tup = std::forward_as_tuple(a...);
			constexpr size_t M = sizeof...(A) - 1;
			auto &k = std::get<M>(tup);
			if constexpr (M == 0)
				detail::run(r, k);
			else if constexpr (M == 1)
				detail::run_red(r, std::get<0>(tup), k);
			else
				multi_red(r, tup, k, std::make_index_sequence<M>{});
		}

	private:
		template <typename T, typename Op>
		static reducer<T, Op> mk(reduction_desc<T, Op> &d) { return reducer<T, Op>{T(), false, d.op}; }
		template <typename T, typename Op>
		static void fin(reduction_desc<T, Op> &d, reducer<T, Op> &r)
		{
			if (r.touched)
				*d.p = d.op(*d.p, r.val);
		}
		// several reduction objects (the reference's diagnostics / viscous limiters only): run serially
		template <int N, typename Tup, typename K, size_t... I>
		void multi_red(nd_range<N> r, Tup &tup, K &k, std::index_sequence<I...>)
		{
			auto reds = std::make_tuple(mk(std::get<I>(tup))...);
			const size_t n0 = r.g[0], n1 = N > 1 ? r.g[N > 1 ? 1 : 0] : 1, n2 = N > 2 ? r.g[N > 2 ? 2 : 0] : 1;
			for (size_t kk = 0; kk < n2; kk++)
				for (size_t j = 0; j < n1; j++)
					for (size_t i = 0; i < n0; i++)
					{
						nd_item<N> it;
						it.gid[0] = i;
						if constexpr (N > 1)
							it.gid[1] = j;
						if constexpr (N > 2)
							it.gid[2] = kk;
						k(it, std::get<I>(reds)...);
					}
			(fin(std::get<I>(tup), std::get<I>(reds)), ...);
		}

	public:
	};

	namespace info
	{
		namespace device
		{
			struct name
			{
			};
			struct version
			{
			};
		}
	}
	struct device
	{
		template <typename I>
		std::string get_info() const
		{
			if constexpr (std::is_same_v<I, info::device::name>)
				return "host-shim (g++/OpenMP)";
			else
				return "oracle";
		}
	};
	struct platform
	{
		static std::vector<platform> get_platforms() { return std::vector<platform>(8); }
		std::vector<device> get_devices() const { return std::vector<device>(64); }
	};

	struct queue
	{
		queue() {}
		queue(const device &) {}
		template <typename F>
		event submit(F f)
		{
			handler h;
			f(h);
			return {};
		}
		event memcpy(void *d, const void *s, size_t n)
		{
			std::memcpy(d, s, n);
			return {};
		}
		event memset(void *d, int v, size_t n)
		{
			std::memset(d, v, n);
			return {};
		}
		void wait() {}
		void wait_and_throw() {}
		device get_device() const { return {}; }
	};

	template <typename T>
	inline T *malloc_device(size_t count, const queue &) { return static_cast<T *>(std::calloc(count ? count : 1, sizeof(T))); }
	template <typename T>
	inline T *malloc_host(size_t count, const queue &) { return static_cast<T *>(std::calloc(count ? count : 1, sizeof(T))); }
	template <typename T>
	inline T *malloc_shared(size_t count, const queue &) { return static_cast<T *>(std::calloc(count ? count : 1, sizeof(T))); }
	inline void *malloc_device(size_t bytes, const queue &) { return std::calloc(bytes ? bytes : 1, 1); }
	inline void *malloc_host(size_t bytes, const queue &) { return std::calloc(bytes ? bytes : 1, 1); }
	inline void *malloc_shared(size_t bytes, const queue &) { return std::calloc(bytes ? bytes : 1, 1); }
	inline void free(void *p, const queue &) { std::free(p); }

	class stream
	{
	public:
		stream(size_t, size_t, handler &) {}
		template <typename T>
		const stream &operator<<(const T &) const { return *this; }
	};

	// math -- plain libm in double, same as a CPU SYCL backend
	template <typename T>
	inline T sqrt(T x) { return std::sqrt(x); }
	template <typename T>
	inline T fabs(T x) { return std::fabs(x); }
	template <typename T>
	inline T abs(T x) { return x < 0 ? -x : x; }
	template <typename T>
	inline T log(T x) { return std::log(x); }
	template <typename T>
	inline T log10(T x) { return std::log10(x); }
	template <typename T>
	inline T exp(T x) { return std::exp(x); }
	template <typename T>
	inline T tanh(T x) { return std::tanh(x); }
	template <typename T>
	inline T ceil(T x) { return std::ceil(x); }
	template <typename T>
	inline T floor(T x) { return std::floor(x); }
	template <typename T, typename U>
	inline T pow(T x, U y) { return std::pow(x, T(y)); }
	template <typename T>
	inline T pown(T x, int n) { return std::pow(x, n); }
	template <typename T, typename U>
	inline auto max(T a, U b) -> std::common_type_t<T, U>
	{
		using C = std::common_type_t<T, U>;
		return C(a) < C(b) ? C(b) : C(a);
	}
	template <typename T, typename U>
	inline auto min(T a, U b) -> std::common_type_t<T, U>
	{
		using C = std::common_type_t<T, U>;
		return C(b) < C(a) ? C(b) : C(a);
	}
	// step(edge, x): 0 if x < edge else 1
	template <typename T>
	inline T step(T edge, T x) { return x < edge ? T(0) : T(1); }
	template <typename T>
	inline bool isnan(T x) { return std::isnan(x); }
	template <typename T>
	inline bool isinf(T x) { return std::isinf(x); }
} // namespace sycl
