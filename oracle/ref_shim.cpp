/*
 * ref_shim.cpp -- C-ABI shim around the UNMODIFIED reference headers.
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no sorting code: it instantiates
 * radix_sort<> / radix_sort_rank<> from the reference's own radix_sort.hpp /
 * radix_sort_rank.hpp (found on the include path, i.e. -I/root/reference; the sources are
 * compiled where they lie and are never copied into this repository) and exposes them to
 * ctypes.  Built by oracle/Makefile into oracle/_ref/libradix_ref.so (git-ignored).
 *
 * Used to (1) pin oracle/rsx_oracle.c against the real reference, (2) generate
 * tests/golden/ref_outputs.npz, (3) serve as the `--impl reference` / cpu_baseline arm of
 * bench.py (cpu_baseline.kind = "reference").
 */
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <type_traits>

#include "radix_sort.hpp"      // reference: radix_sort.hpp:98-115
#include "radix_sort_rank.hpp" // reference: radix_sort_rank.hpp:97-112

namespace {

// The record shapes the reference's tests and tutorial listings sort.
struct rec16_u8 { // radix_tests.cpp:15-18  struct sortrec { uint8_t key; const char *name; }
	uint8_t key;
	const char *name;
};
struct rec8_u32 { // BASELINE config C4b: { u32 key, u32 payload }
	uint32_t key;
	uint32_t payload;
};
struct rec16_u64 { // radix_sort_u64_multipass.c:8-11 shape: 64-bit key + 8-byte payload
	uint64_t key;
	uint64_t payload;
};

enum TypeCode {
	T_U8 = 0, T_U16, T_U32, T_U64, T_I8, T_I16, T_I32, T_I64, T_F32, T_F64,
	T_REC16_U8, T_REC8_U32, T_REC16_U64
};

template <typename T> struct is_record : std::false_type {};
template <> struct is_record<rec16_u8> : std::true_type {};
template <> struct is_record<rec8_u32> : std::true_type {};
template <> struct is_record<rec16_u64> : std::true_type {};

// Ascending key: the reference's default KDF for scalars, the record's key member otherwise
// (radix_tests.cpp:41-43).  Descending: bitwise complement (README.md:564-574,
// radix_tests.cpp:111-113,175-177).
template <typename T, bool Desc> struct keyfn {
	auto operator()(const T &v) const {
		if constexpr (is_record<T>::value) {
			using K = decltype(v.key);
			return Desc ? static_cast<K>(~v.key) : v.key;
		} else {
			auto k = basic_kdfs::kdf(v);
			using K = decltype(k);
			return Desc ? static_cast<K>(~k) : k;
		}
	}
};

template <typename T> int sort_one(void *src, void *aux, size_t n, int desc) {
	T *s = static_cast<T *>(src), *a = static_cast<T *>(aux), *r;
	if constexpr (!is_record<T>::value) {
		if (!desc)
			r = radix_sort(s, a, n); // default-KDF call, exactly as radix_experiment.cpp:205
		else
			r = radix_sort(s, a, n, keyfn<T, true>{});
	} else {
		r = desc ? radix_sort(s, a, n, keyfn<T, true>{}) : radix_sort(s, a, n, keyfn<T, false>{});
	}
	return r == a ? 1 : 0;
}

template <typename T, typename I> int rank_one(const void *src, void *ib, size_t n, int desc) {
	const T *s = static_cast<const T *>(src);
	I *b = static_cast<I *>(ib), *r;
	r = desc ? radix_sort_rank(s, b, n, keyfn<T, true>{}) : radix_sort_rank(s, b, n, keyfn<T, false>{});
	return r == b ? 0 : 1;
}

template <typename T> int rank_idx(const void *src, void *ib, size_t n, int idx_bytes, int desc) {
	switch (idx_bytes) {
	case 1: return rank_one<T, uint8_t>(src, ib, n, desc);
	case 2: return rank_one<T, uint16_t>(src, ib, n, desc);
	case 4: return rank_one<T, uint32_t>(src, ib, n, desc);
	case 8: return rank_one<T, uint64_t>(src, ib, n, desc);
	}
	return -1;
}

} // namespace

#define DISPATCH(code, CALL)                                  \
	switch (code) {                                           \
	case T_U8: return CALL(uint8_t);                          \
	case T_U16: return CALL(uint16_t);                        \
	case T_U32: return CALL(uint32_t);                        \
	case T_U64: return CALL(uint64_t);                        \
	case T_I8: return CALL(int8_t);                           \
	case T_I16: return CALL(int16_t);                         \
	case T_I32: return CALL(int32_t);                         \
	case T_I64: return CALL(int64_t);                         \
	case T_F32: return CALL(float);                           \
	case T_F64: return CALL(double);                          \
	case T_REC16_U8: return CALL(rec16_u8);                   \
	case T_REC8_U32: return CALL(rec8_u32);                   \
	case T_REC16_U64: return CALL(rec16_u64);                 \
	}                                                         \
	return -1

extern "C" {

// Returns 1 if the reference returned `aux`, 0 if `src`, -1 on an unknown type code.
int ref_radix_sort(int type_code, void *src, void *aux, size_t n, int descending) {
#define CALL_SORT(T) sort_one<T>(src, aux, n, descending)
	DISPATCH(type_code, CALL_SORT);
}

// Returns 1 if the reference returned `index_buffer + n`, 0 if `index_buffer`, -1 on error.
// NOTE: this is the header AS SHIPPED, including radix_sort_rank.hpp:82 (SURVEY.md finding 2).
int ref_radix_sort_rank(int type_code, const void *src, void *index_buffer, size_t n,
                        int idx_bytes, int descending) {
#define CALL_RANK(T) rank_idx<T>(src, index_buffer, n, idx_bytes, descending)
	DISPATCH(type_code, CALL_RANK);
}

// The n-sweep of the reference's benchmark harness (radix_bench.cpp:86-138) for uint32_t keys, timed
// in here so that tiny n is not dominated by FFI overhead: kind 0 = radix_sort, 1 = radix_sort_rank
// (u32 indices), 2 = std::sort, 3 = qsort.  The input is restored from `pristine` before every
// iteration OUTSIDE the timed region (the reference's own loop does not, radix_bench.cpp:91-93);
// returns the best wall-clock seconds of `iters` runs (CLOCK_MONOTONIC_RAW, radix_experiment.cpp:203).
static int cmp_u32(const void *a, const void *b) {
	const uint32_t x = *static_cast<const uint32_t *>(a), y = *static_cast<const uint32_t *>(b);
	return x < y ? -1 : (x > y ? 1 : 0);
}
double ref_time_u32(int kind, const uint32_t *pristine, uint32_t *work, uint32_t *aux /* 2n for kind 1 */, size_t n,
                    int iters) {
	double best = 1e300;
	for (int it = 0; it < iters; ++it) {
		memcpy(work, pristine, n * sizeof(uint32_t));
		timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC_RAW, &t0);
		switch (kind) {
		case 0: { volatile auto *r = radix_sort(work, aux, n); (void)r; break; }
		case 1: { volatile auto *r = radix_sort_rank(work, aux, n); (void)r; break; }
		case 2: std::sort(work, work + n); break;
		case 3: qsort(work, n, sizeof(uint32_t), cmp_u32); break;
		default: return -1.0;
		}
		clock_gettime(CLOCK_MONOTONIC_RAW, &t1);
		const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
		if (dt < best)
			best = dt;
	}
	return best;
}

size_t ref_record_bytes(int type_code) {
#define CALL_SIZE(T) sizeof(T)
	DISPATCH(type_code, CALL_SIZE);
}

} // extern "C"
