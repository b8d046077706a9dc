// probe_bulk.cu -- can TMA bulk stores (shared -> global) carry a radix pass's write-out?
//
// A scatter pass writes, per 10 240-key tile, one short run per digit bucket (u32: ~160 bytes).
// This probe replays exactly that store pattern -- 256 buckets, every tile appends one run to each --
// from a static shared buffer, (a) with per-thread 4-byte STG like the kernel's write-out and (b) with
// one cp.async.bulk.global.shared::cta per run issued by 256 threads, and reports GB/s written.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_bulk probe_bulk.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kThreads = 512, kBins = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE> // 0 = STG.32, 1 = bulk per run, 2 = STG.128
__global__ void __launch_bounds__(kThreads, 2) store_kernel(unsigned char *out, size_t bucket_bytes, uint32_t tiles,
                                                             uint32_t run_bytes, unsigned int *ticket) {
	extern __shared__ __align__(128) unsigned char smem[];
	const uint32_t tile_bytes = run_bytes * kBins;
	for (uint32_t i = threadIdx.x; i < tile_bytes / 4; i += kThreads)
		reinterpret_cast<uint32_t *>(smem)[i] = i * 2654435761u;
	__shared__ uint32_t s_tile;
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncthreads();
	for (;;) {
		if (threadIdx.x == 0)
			s_tile = atomicAdd(ticket, 1u);
		__syncthreads();
		const uint32_t tile = s_tile;
		if (tile >= tiles)
			break;
		if (MODE == 1) {
			if (threadIdx.x < kBins) {
				const uint32_t d = threadIdx.x;
				unsigned char *g = out + (size_t)d * bucket_bytes + (size_t)tile * run_bytes;
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
				             "r"(smem_u32(smem + d * run_bytes)), "r"(run_bytes)
				             : "memory");
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
			}
		} else if (MODE == 0) {
			const uint32_t per = run_bytes / 4;
			for (uint32_t s = threadIdx.x; s < tile_bytes / 4; s += kThreads) {
				const uint32_t d = s / per, o = s % per;
				reinterpret_cast<uint32_t *>(out + (size_t)d * bucket_bytes + (size_t)tile * run_bytes)[o] =
				    reinterpret_cast<const uint32_t *>(smem)[s];
			}
		} else {
			const uint32_t per = run_bytes / 16;
			for (uint32_t s = threadIdx.x; s < tile_bytes / 16; s += kThreads) {
				const uint32_t d = s / per, o = s % per;
				reinterpret_cast<uint4 *>(out + (size_t)d * bucket_bytes + (size_t)tile * run_bytes)[o] =
				    reinterpret_cast<const uint4 *>(smem)[s];
			}
		}
		__syncthreads();
	}
}

template <int MODE> float run(unsigned char *out, size_t total, uint32_t run_bytes, unsigned int *ticket, int sms) {
	const uint32_t tile_bytes = run_bytes * kBins;
	const uint32_t tiles = (uint32_t)(total / tile_bytes);
	const size_t bucket_bytes = (size_t)tiles * run_bytes;
	cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e30f;
	for (int it = 0; it < 4; ++it) {
		cudaMemset(ticket, 0, 4);
		cudaEventRecord(e0);
		store_kernel<MODE><<<sms * 2, kThreads, tile_bytes>>>(out, bucket_bytes, tiles, run_bytes, ticket);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (it && ms < best)
			best = ms;
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
	return best;
}

int main() {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const size_t total = 4ULL << 30;
	unsigned char *out;
	unsigned int *ticket;
	cudaMalloc(&out, total + (1 << 20));
	cudaMalloc(&ticket, 4);
	const uint32_t runs[] = {64, 96, 160, 320, 640};
	for (uint32_t rb : runs) {
		const float a = run<0>(out, total, rb, ticket, sms);
		const float b = run<1>(out, total, rb, ticket, sms);
		const float c = run<2>(out, total, rb, ticket, sms);
		printf("{\"run_bytes\": %u, \"stg32_ms\": %.3f, \"stg32_gbs\": %.0f, \"bulk_ms\": %.3f, \"bulk_gbs\": %.0f, \"stg128_ms\": %.3f, \"stg128_gbs\": %.0f}\n",
		       rb, a, total / a / 1e6, b, total / b / 1e6, c, total / c / 1e6);
	}
	return 0;
}
