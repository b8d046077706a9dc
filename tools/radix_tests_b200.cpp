/*
	radix_tests_b200.cpp -- the reference's test scenarios (radix_tests.cpp) recompiled against
	the drop-in headers in include/, i.e. running on the GPU through librsx.so.

	Host C++ mirror of the reference interface: the calls below have the reference's syntax;
	the only edits a reference user makes are the KDF lambdas -> basic_kdfs descriptor functors
	(radix_tests.cpp:41-43 -> by_member<&sortrec::key>, :175-177 -> descending).
	Unlike the reference's tests this also asserts STABILITY (the reference only prints it).

	Scenarios: test_sortrec (:45-69), by-value version of test_sortrec_ptr (:121-146),
	test_float (:156-173), test_int + reverse re-sort (:179-207), test_rank_sortrec with uint8_t
	indices (:71-105), plus a two-column rank sort that the shipped reference header gets wrong.
	Exit code 0 = all passed.
*/
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "radix_sort.hpp"
#include "radix_sort_rank.hpp"

struct sortrec {
	uint8_t key;
	const char *name;
};

static const sortrec source_arr[] = {
	{255, "1st 255"}, {45, "1st 45"}, {3, "3"}, {45, "2nd 45"}, {2, "2"}, {45, "3rd 45"}, {1, "1"}, {255, "2nd 255"},
};
static const size_t N8 = sizeof(source_arr) / sizeof(source_arr[0]);

static bool report(const char *what, bool ok) {
	printf("%-58s %s\n", what, ok ? "OK" : "FAILED");
	return ok;
}

static bool test_sortrec() {
	std::vector<sortrec> src(source_arr, source_arr + N8), aux(N8);
	sortrec *res = radix_sort(src.data(), aux.data(), N8, basic_kdfs::by_member<&sortrec::key>{});
	if (!res)
		return report("Sorting struct sortrec", false);
	const char *want[] = {"1", "2", "3", "1st 45", "2nd 45", "3rd 45", "1st 255", "2nd 255"};
	bool ok = res == aux.data(); // one live column -> the reference returns aux
	for (size_t i = 0; i < N8; ++i)
		ok = ok && strcmp(res[i].name, want[i]) == 0;
	return report("Sorting struct sortrec (stable, result in aux)", ok);
}

static bool test_sortrec_reverse() {
	std::vector<sortrec> src(source_arr, source_arr + N8), aux(N8);
	sortrec *res = radix_sort(src.data(), aux.data(), N8, basic_kdfs::by_member<&sortrec::key, basic_kdfs::desc>{});
	if (!res)
		return report("Sorting struct sortrec (reverse)", false);
	const char *want[] = {"1st 255", "2nd 255", "1st 45", "2nd 45", "3rd 45", "3", "2", "1"};
	bool ok = true;
	for (size_t i = 0; i < N8; ++i)
		ok = ok && strcmp(res[i].name, want[i]) == 0;
	return report("Sorting struct sortrec (reverse via ~key, stable)", ok);
}

static bool test_float() {
	float src[] = {128.0f, 646464.0f, 0.0f, -0.0f, -0.5f, 0.5f, -128.0f, -INFINITY, NAN, INFINITY};
	const size_t N = sizeof(src) / sizeof(src[0]);
	float aux[N];
	float *res = radix_sort(src, aux, N); // default KDF, exactly the reference's call
	if (!res)
		return report("Sorting float[]", false);
	const uint32_t want[] = {0xff800000, 0xc3000000, 0xbf000000, 0x80000000, 0x00000000,
	                         0x3f000000, 0x43000000, 0x491dd400, 0x7f800000, 0x7fc00000}; // README.md:612-623
	bool ok = res == src;
	for (size_t i = 0; i < N; ++i) {
		uint32_t b;
		memcpy(&b, res + i, 4);
		ok = ok && b == want[i];
	}
	return report("Sorting float[] incl. -0, +-inf, NaN (bit patterns)", ok);
}

static bool test_int() {
	std::default_random_engine generator;
	std::normal_distribution<double> distribution(0.0, 1.0e9);
	const size_t N = 50000;
	std::vector<int> buf(2 * N), ref(N);
	int *src = buf.data(), *aux = src + N;
	for (size_t i = 0; i < N; ++i) {
		double a = std::max(-2147483648.0, std::min(2147483647.0, distribution(generator)));
		src[i] = ref[i] = int(a);
	}
	std::stable_sort(ref.begin(), ref.end());
	int *res = radix_sort(src, aux, N);
	bool ok = res && std::equal(ref.begin(), ref.end(), res);
	if (ok) {
		int *other = res == src ? aux : src;
		res = radix_sort(res, other, N, basic_kdfs::descending{}); // radix_tests.cpp:198 re-sorts the result
		std::reverse(ref.begin(), ref.end());
		ok = res && std::equal(ref.begin(), ref.end(), res);
	}
	return report("Sorting int[50000], then re-sorting descending", ok);
}

static bool test_rank_sortrec() {
	uint8_t ib[2 * 8];
	uint8_t *ranks = radix_sort_rank(source_arr, ib, N8, basic_kdfs::by_member<&sortrec::key>{});
	const uint8_t want[] = {6, 4, 2, 1, 3, 5, 0, 7};
	bool ok = ranks == ib + N8 && std::equal(want, want + N8, ranks);
	return report("Rank sorting struct sortrec (uint8_t indices)", ok);
}

static bool test_rank_two_columns() {
	// radix_sort_u32_ranks.c:8-19; the shipped radix_sort_rank.hpp returns {7,5,1,2,0,3,4,6,8,9} here
	const uint32_t keys[] = {4255, 45, 45, 45, 0, 0x800201, 255, 256, 0xFFFFFFFF, 4255};
	uint32_t ib[20];
	uint32_t *ranks = radix_sort_rank(keys, ib, 10);
	const uint32_t want[] = {4, 1, 2, 3, 6, 7, 0, 9, 5, 8};
	return report("Rank sorting u32 with 4 live columns (listing 6 output)", ranks && std::equal(want, want + 10, ranks));
}

int main() {
	bool passed = test_sortrec() & test_sortrec_reverse() & test_float() & test_int() & test_rank_sortrec() &
	              test_rank_two_columns();
	if (!passed) {
		fprintf(stderr, "Tests failed (last status %d: %s; %s).\n", radix_sort_last_status(),
		        rsx_strerror(radix_sort_last_status()), rsx_last_cuda_error());
		return EXIT_FAILURE;
	}
	printf("All tests OK.\n");
	return EXIT_SUCCESS;
}
