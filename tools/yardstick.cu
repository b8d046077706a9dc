// tools/yardstick.cu -- bench-only context, NOT product code and never linked into librsx.so.
//
// Puts three on-box yardsticks next to our kernels so that DESIGN.md can state what the
// hardware allows:
//   1. device-to-device copy bandwidth at the workload's size (the roofline denominator's twin)
//   2. cub::DeviceRadixSort (the library onesweep that ships with CUDA 12.9) on the same keys
//   3. micro-rates that bound the per-key instruction budget of the scatter pass:
//      __match_any_sync on random 8-bit digits, shared-memory atomics (random vs lane-private)
//
//   ./yardstick [n_keys=1000000000]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#define CK(x)                                                                          \
	do {                                                                               \
		cudaError_t e = (x);                                                           \
		if (e != cudaSuccess) {                                                        \
			fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));  \
			exit(1);                                                                   \
		}                                                                              \
	} while (0)

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
	return z;
}
template <typename K> __global__ void fill(K *d, size_t n, unsigned long long seed) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		d[i] = (K)mix64(seed * 0x9E3779B97F4A7C15ULL + i);
}

// ---- micro: match_any throughput -------------------------------------------------------------------
__global__ void k_match(uint32_t *out, int iters, uint32_t mask) {
	uint32_t x = (uint32_t)mix64(blockIdx.x * 1024 + threadIdx.x), acc = 0;
	for (int i = 0; i < iters; ++i) {
		acc += __match_any_sync(0xFFFFFFFFu, x & mask);
		x = x * 1664525u + 1013904223u + acc;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// ballot-based 8-bit match (8 votes)
__global__ void k_ballot8(uint32_t *out, int iters) {
	uint32_t x = (uint32_t)mix64(blockIdx.x * 1024 + threadIdx.x), acc = 0;
	for (int i = 0; i < iters; ++i) {
		uint32_t d = x >> 24, peers = 0xFFFFFFFFu;
#pragma unroll
		for (int b = 0; b < 8; ++b) {
			const bool p = (d >> b) & 1;
			const uint32_t v = __ballot_sync(0xFFFFFFFFu, p);
			peers &= p ? v : ~v;
		}
		acc += peers;
		x = x * 1664525u + 1013904223u + acc;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// shared atomics: mode 0 = one 256-bin histogram per warp, random digits; 1 = lane-private (bank = lane)
template <int MODE> __global__ void k_atoms(uint32_t *out, int iters) {
	extern __shared__ uint32_t sh[];
	for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	uint32_t x = (uint32_t)mix64(blockIdx.x * 1024 + threadIdx.x);
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	for (int i = 0; i < iters; ++i) {
		const uint32_t d = x >> 24;
		if (MODE == 0) atomicAdd(&sh[(warp & 31) * 256 + d], 1u);
		else atomicAdd(&sh[d * 32 + lane], 1u);
		x = x * 1664525u + 1013904223u;
	}
	__syncthreads();
	out[blockIdx.x * blockDim.x + threadIdx.x] = sh[threadIdx.x] + x;
}

template <typename F> float time_ms(F f, int reps = 5) {
	cudaEvent_t a, b;
	CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	f();
	CK(cudaDeviceSynchronize());
	float best = 1e30f;
	for (int r = 0; r < reps; ++r) {
		CK(cudaEventRecord(a));
		f();
		CK(cudaEventRecord(b));
		CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		if (ms < best) best = ms;
	}
	return best;
}

template <typename K> void cub_sort(size_t n, const char *name) {
	K *in, *out, *pristine; void *tmp = nullptr; size_t tmp_bytes = 0;
	CK(cudaMalloc(&in, n * sizeof(K))); CK(cudaMalloc(&out, n * sizeof(K))); CK(cudaMalloc(&pristine, n * sizeof(K)));
	fill<K><<<148 * 8, 256>>>(pristine, n, 2);
	CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, in, out, n));
	CK(cudaMalloc(&tmp, tmp_bytes));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	float best = 1e30f, sum = 0; int reps = 5;
	for (int r = 0; r < reps + 1; ++r) {
		CK(cudaMemcpy(in, pristine, n * sizeof(K), cudaMemcpyDeviceToDevice));
		CK(cudaEventRecord(a));
		CK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, in, out, n));
		CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		if (r) { sum += ms; if (ms < best) best = ms; }
	}
	const double passes = sizeof(K), bytes = (double)n * sizeof(K) * (1 + 2 * passes);
	printf("{\"yardstick\":\"cub::DeviceRadixSort::SortKeys\",\"type\":\"%s\",\"n\":%zu,\"ms_best\":%.3f,\"ms_mean\":%.3f,"
	       "\"gkeys_s\":%.2f,\"alg_GBps\":%.1f,\"temp_bytes\":%zu}\n",
	       name, n, best, sum / reps, n / (best * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9, tmp_bytes);
	float cp = time_ms([&] { CK(cudaMemcpyAsync(out, in, n * sizeof(K), cudaMemcpyDeviceToDevice)); });
	printf("{\"yardstick\":\"cudaMemcpy D2D\",\"bytes\":%zu,\"ms\":%.3f,\"GBps_rw\":%.1f}\n", n * sizeof(K), cp,
	       2.0 * n * sizeof(K) / (cp * 1e-3) / 1e9);
	cudaFree(in); cudaFree(out); cudaFree(pristine); cudaFree(tmp);
}

int main(int argc, char **argv) {
	size_t n = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1000000000ULL;
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
	printf("{\"device\":\"%s\",\"sms\":%d,\"smem_optin\":%zu,\"l2\":%d,\"clock_khz\":%d}\n", p.name, p.multiProcessorCount,
	       p.sharedMemPerBlockOptin, p.l2CacheSize, clk);
	uint32_t *out; CK(cudaMalloc(&out, 148 * 8 * 1024 * 4));
	const int iters = 4096, grid = 148 * 2, thr = 1024;
	const double ops = (double)grid * thr * iters / 32.0; // warp-instructions
	for (uint32_t mask : {0xFFu << 24, 0xFu << 24, 0x1u << 24, 0u}) {
		float ms = time_ms([&] { k_match<<<grid, thr>>>(out, iters, mask); });
		printf("{\"micro\":\"match_any\",\"digit_mask\":\"%08x\",\"warp_instr_per_clk_per_sm\":%.4f,\"ms\":%.3f}\n", mask,
		       ops / (ms * 1e-3) / 148.0 / (clk * 1e3), ms);
	}
	{
		float ms = time_ms([&] { k_ballot8<<<grid, thr>>>(out, iters); });
		printf("{\"micro\":\"ballot8_match\",\"warp_match_per_clk_per_sm\":%.4f,\"ms\":%.3f}\n", ops / (ms * 1e-3) / 148.0 / (clk * 1e3), ms);
	}
	{
		CK(cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
		CK(cudaFuncSetAttribute(k_atoms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
		float m0 = time_ms([&] { k_atoms<0><<<148, thr, 32768>>>(out, iters); });
		float m1 = time_ms([&] { k_atoms<1><<<148, thr, 32768>>>(out, iters); });
		const double lane_ops = 148.0 * thr * iters;
		printf("{\"micro\":\"smem_atomic_add\",\"per_warp_hist_lanes_per_clk_per_sm\":%.3f,\"lane_private_lanes_per_clk_per_sm\":%.3f}\n",
		       lane_ops / (m0 * 1e-3) / 148.0 / (clk * 1e3), lane_ops / (m1 * 1e-3) / 148.0 / (clk * 1e3));
	}
	cub_sort<uint32_t>(n, "u32");
	cub_sort<unsigned long long>(n, "u64");
	return 0;
}
