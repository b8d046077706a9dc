/*
	radix_b200.cpp -- GPU counterpart of the reference's `radix` experiment harness
	(radix_experiment.cpp:176-285): same positional arguments, same key-file format, now timing
	the device path with a CPU yardstick beside it: std::stable_sort by the same derived key on one
	host core, whose output is byte-identical to the reference's radix_sort (SURVEY.md finding 1) and
	therefore doubles as the verification (memcmp, not just "is it ordered", radix_experiment.cpp:208-212).

	    ./radix_b200 <count> [<use_mmap> <use_huge> <type> <hex-mask>]

	<count> 0 = whole file.  <use_mmap>/<use_huge> are accepted for command-line compatibility
	and ignored (buffers are cudaMallocHost / cudaMalloc).  The key file is 40M_32bit_keys.dat
	(raw little-endian bytes, reference Makefile:79-82) in the current directory; if it is
	missing, 160 000 000 seeded bytes are generated instead (the reference file is /dev/urandom
	output and cannot be reproduced).  Prints one time for the host-buffer call (H2D + sort +
	D2H, what a drop-in user sees), one for device-resident buffers, the CPU yardstick's time and a
	64-bit FNV-1a digest of the sorted bytes (tests/test_headers.py compares it with the oracle's).
*/
#include <algorithm>
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "radix_sort.hpp"

static uint64_t mix64(uint64_t z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
	return z;
}

static std::vector<unsigned char> load_keys(const char *fn) {
	std::vector<unsigned char> buf;
	if (FILE *f = fopen(fn, "rb")) {
		fseek(f, 0, SEEK_END);
		long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		buf.resize(sz);
		if (fread(buf.data(), 1, sz, f) != (size_t)sz)
			buf.clear();
		fclose(f);
		printf("Read %zu bytes from '%s'.\n", buf.size(), fn);
	}
	if (buf.empty()) {
		buf.resize(160000000);
		uint64_t *w = reinterpret_cast<uint64_t *>(buf.data());
		for (size_t i = 0; i < buf.size() / 8; ++i)
			w[i] = mix64(1 * 0x9E3779B97F4A7C15ULL + i);
		printf("'%s' not found: generated %zu seeded bytes.\n", fn, buf.size());
	}
	return buf;
}

static uint64_t fnv1a(const void *p, size_t bytes) {
	const unsigned char *b = static_cast<const unsigned char *>(p);
	uint64_t h = 0xCBF29CE484222325ULL;
	for (size_t i = 0; i < bytes; ++i)
		h = (h ^ b[i]) * 0x100000001B3ULL;
	return h;
}

template <typename T> static bool verify(const T *keys, size_t n) {
	for (size_t i = 1; i < n; ++i)
		if (basic_kdfs::kdf(keys[i - 1]) > basic_kdfs::kdf(keys[i])) {
			printf("Sort failed at %zu.\n", i);
			return false;
		}
	return true;
}

template <typename T> static int run(const std::vector<unsigned char> &file, size_t entries, uint64_t mask) {
	size_t n = file.size() / sizeof(T);
	if (entries && entries < n)
		n = entries;
	T *src, *aux;
	if (cudaMallocHost((void **)&src, n * sizeof(T)) != cudaSuccess || cudaMallocHost((void **)&aux, n * sizeof(T)) != cudaSuccess) {
		fprintf(stderr, "cudaMallocHost failed: no CUDA device? (there is no CPU path)\n");
		return 2;
	}
	memcpy(src, file.data(), n * sizeof(T));
	if (mask != (uint64_t)-1) { // radix_experiment.cpp:188-198: force column skipping
		printf("Applying value mask to input.\n");
		for (size_t i = 0; i < n; ++i) {
			uint64_t b = 0;
			memcpy(&b, src + i, sizeof(T));
			b &= mask;
			memcpy(src + i, &b, sizeof(T));
		}
	}
	std::vector<T> pristine(src, src + n);
	printf("Sorting %zu entries...\n", n);
	radix_sort(src, aux, std::min<size_t>(n, 1024)); // warm-up: CUDA context, workspace
	memcpy(src, pristine.data(), n * sizeof(T));
	auto t0 = std::chrono::steady_clock::now();
	T *sorted = radix_sort(src, aux, n);
	auto t1 = std::chrono::steady_clock::now();
	if (!sorted) {
		fprintf(stderr, "radix_sort failed: %s (%s)\n", rsx_strerror(radix_sort_last_status()), rsx_last_cuda_error());
		return 2;
	}
	if (!verify(sorted, n))
		return 1;
	const double host_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();

	// CPU yardstick + verification: a stable sort by the derived key (one core)
	std::vector<T> cpu(pristine);
	auto c0 = std::chrono::steady_clock::now();
	std::stable_sort(cpu.begin(), cpu.end(), [](const T &a, const T &b) { return basic_kdfs::kdf(a) < basic_kdfs::kdf(b); });
	auto c1 = std::chrono::steady_clock::now();
	const double cpu_ms = std::chrono::duration<double, std::milli>(c1 - c0).count();
	if (memcmp(cpu.data(), sorted, n * sizeof(T)) != 0) {
		printf("Sort differs from std::stable_sort by derived key (host buffers).\n");
		return 1;
	}

	T *dsrc, *daux;
	cudaMalloc((void **)&dsrc, n * sizeof(T));
	cudaMalloc((void **)&daux, n * sizeof(T));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int r = 0; r < 4; ++r) {
		cudaMemcpy(dsrc, pristine.data(), n * sizeof(T), cudaMemcpyHostToDevice);
		cudaEventRecord(e0);
		T *dres = radix_sort(dsrc, daux, n);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (r && ms < best)
			best = ms;
		if (!dres)
			return 2;
		if (r == 3) { // the device-resident result, byte for byte
			std::vector<T> back(n);
			cudaMemcpy(back.data(), dres, n * sizeof(T), cudaMemcpyDeviceToHost);
			if (memcmp(cpu.data(), back.data(), n * sizeof(T)) != 0) {
				printf("Sort differs from std::stable_sort by derived key (device buffers).\n");
				return 1;
			}
		}
	}
	const size_t np = std::min<size_t>(n, 10);
	for (size_t i = 0; i < np; ++i) {
		uint64_t b = 0;
		memcpy(&b, sorted + i, sizeof(T));
		printf("%08zx: %0*" PRIx64 "\n", i, (int)(2 * sizeof(T)), b);
	}
	printf("Sorted %zu entries in %.4f ms (host buffers: H2D + sort + D2H), %.1f Mkeys/s\n", n, host_ms, n / host_ms / 1e3);
	printf("Sorted %zu entries in %.4f ms (device-resident buffers), %.1f Mkeys/s\n", n, best, n / best / 1e3);
	printf("Sorted %zu entries in %.4f ms (CPU yardstick: std::stable_sort by derived key, 1 core), %.1f Mkeys/s; "
	       "device outputs memcmp-equal\n", n, cpu_ms, n / cpu_ms / 1e3);
	printf("digest fnv1a64=%016" PRIx64 "\n", fnv1a(sorted, n * sizeof(T)));
	cudaFreeHost(src);
	cudaFreeHost(aux);
	cudaFree(dsrc);
	cudaFree(daux);
	return 0;
}

int main(int argc, char *argv[]) {
	if (argc == 1) {
		printf("Usage: %s <count> [<use_mmap> <use_huge> <uint8_t|uint16_t|uint32_t|uint64_t|int32_t|int64_t|float|double> <hex-mask>]\n", argv[0]);
		return 0;
	}
	const size_t entries = strtoull(argv[1], nullptr, 10);
	const char *ktype = argc > 4 ? argv[4] : "uint32_t";
	const uint64_t mask = argc > 5 ? strtoull(argv[5], nullptr, 16) : (uint64_t)-1;
	const char *fn = "40M_32bit_keys.dat";
	printf("src='%s', entries=%zu, type='%s', mask=0x%08" PRIx64 "\n", fn, entries, ktype, mask);
	const std::vector<unsigned char> file = load_keys(fn);
	if (!strcmp(ktype, "uint8_t")) return run<uint8_t>(file, entries, mask);
	if (!strcmp(ktype, "uint16_t")) return run<uint16_t>(file, entries, mask);
	if (!strcmp(ktype, "uint32_t")) return run<uint32_t>(file, entries, mask);
	if (!strcmp(ktype, "uint64_t")) return run<uint64_t>(file, entries, mask);
	if (!strcmp(ktype, "int32_t")) return run<int32_t>(file, entries, mask);
	if (!strcmp(ktype, "int64_t")) return run<int64_t>(file, entries, mask);
	if (!strcmp(ktype, "float")) return run<float>(file, entries, mask);
	if (!strcmp(ktype, "double")) return run<double>(file, entries, mask);
	printf("Error: unknown key type, '%s'.\n", ktype);
	return 100;
}
