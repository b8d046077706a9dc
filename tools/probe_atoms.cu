// tools/probe_atoms.cu -- hardware probe, bench-only context (not linked into librsx.so).
//
// Question: when the 32 lanes of ONE warp instruction do atomicAdd(+1) with return on
// shared-memory counters, and several lanes hit the SAME counter, are the returned values
// handed out in ascending lane order?  And are back-to-back atomics of one warp (item i, then
// item i+1, no barrier in between) applied in program order?
//
// If both hold, `rank = atomicAdd(&warp_counter[digit], 1)` is a complete STABLE ranking of a
// warp-striped tile in one shared-memory instruction per key, instead of a match_any (measured
// here at ~60 clk/SM per 32 random 8-bit digits) or an 8-vote ballot match (~24 clk/SM).
//
// The probe computes the provably stable rank with ballots and compares it with the ticket
// rank over many warps, CTAs, iterations and digit distributions, with all warps of the SM
// hammering shared memory concurrently.  Prints mismatch counts.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL; z ^= z >> 27; z *= 0x94D049BB133111EBULL; z ^= z >> 31;
	return z;
}

template <int ITEMS>
__global__ void __launch_bounds__(512) probe(unsigned long long *mismatch, unsigned long long *groups, int iters, uint32_t digit_mask, unsigned long long seed) {
	__shared__ uint32_t wh[16][256];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt = (1u << lane) - 1u;
	unsigned long long bad = 0, grp = 0;
	for (int it = 0; it < iters; ++it) {
		for (int b = lane; b < 256; b += 32) wh[warp][b] = 0;
		__syncwarp();
		uint32_t d[ITEMS], ticket[ITEMS], want[ITEMS];
#pragma unroll
		for (int i = 0; i < ITEMS; ++i)
			d[i] = (uint32_t)(mix64(seed + ((unsigned long long)blockIdx.x << 40) + ((unsigned long long)it << 20) + threadIdx.x * 64 + i) >> 24) & digit_mask;
		// ticket ranks: ITEMS back-to-back atomics, nothing in between
#pragma unroll
		for (int i = 0; i < ITEMS; ++i)
			ticket[i] = atomicAdd(&wh[warp][d[i]], 1u);
		__syncwarp();
		// reference: stable rank from ballots + a private running count per digit kept in smem (reset first)
		for (int b = lane; b < 256; b += 32) wh[warp][b] = 0;
		__syncwarp();
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
			for (int b = 0; b < 8; ++b) {
				const bool p = (d[i] >> b) & 1;
				const uint32_t v = __ballot_sync(0xFFFFFFFFu, p);
				peers &= p ? v : ~v;
			}
			const uint32_t leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (lane == leader) { old = wh[warp][d[i]]; wh[warp][d[i]] = old + __popc(peers); }
			__syncwarp();
			old = __shfl_sync(0xFFFFFFFFu, old, leader);
			want[i] = old + __popc(peers & lt);
			grp += (__popc(peers) > 1);
		}
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) bad += ticket[i] != want[i];
		__syncwarp();
	}
	atomicAdd(mismatch, bad);
	atomicAdd(groups, grp);
}

int main() {
	unsigned long long *d; CK(cudaMalloc(&d, 16));
	for (uint32_t mask : {0xFFu, 0x3Fu, 0xFu, 0x3u, 0x1u, 0x0u}) {
		CK(cudaMemset(d, 0, 16));
		probe<16><<<148 * 2, 512>>>(d, d + 1, 2000, mask, 12345 + mask);
		CK(cudaDeviceSynchronize());
		unsigned long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
		printf("{\"probe\":\"smem atomicAdd lane order\",\"digit_mask\":\"%02x\",\"keys\":%llu,\"keys_in_collision_groups\":%llu,\"mismatches\":%llu}\n",
		       mask, 148ULL * 2 * 512 * 2000 * 16, h[1], h[0]);
	}
	return 0;
}
