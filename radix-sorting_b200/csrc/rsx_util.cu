// rsx_util.cu -- small kernels around the hot path: rank-sort index fix-ups, the seeded key
// generator (device twin of keygen.py) and the on-device verifier used at sizes no CPU oracle
// can hold.
#include "rsx_device.cuh"

namespace rsx {

namespace {

// radix_sort_rank.hpp:52-57: on early exit the reference returns the identity permutation in
// the first half of index_buffer.  The host cannot know about the early exit before the
// passes are enqueued, so this kernel checks the device pass table itself.
template <typename I>
__global__ void iota_if_early_kernel(I *ib, size_t n, const Ctl *ctl) {
	if (!ctl->early_exit)
		return;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		ib[i] = (I)i;
}

// 1- and 2-byte IdxType (radix_tests.cpp:75 sorts with uint8_t indices): the passes carry
// 32-bit indices in workspace, this narrows them into the half the parity rule designates.
template <typename I>
__global__ void narrow_index_kernel(const uint32_t *w0, const uint32_t *w1, I *ib, size_t n, const Ctl *ctl) {
	if (ctl->early_exit)
		return; // identity already written by iota_if_early_kernel
	const uint32_t ncols = ctl->ncols;
	if (ncols == 0)
		return;
	// pass j writes pl_buf[j & 1]; the last live pass is j = ncols - 1
	const uint32_t *from = ((ncols - 1) & 1u) ? w1 : w0;
	I *to = (ncols & 1u) ? ib + n : ib;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		to[i] = (I)from[i];
}

__device__ __forceinline__ unsigned long long stream_word(unsigned long long seed, unsigned long long i, unsigned long long stream) {
	const unsigned long long base = (seed + 0x632BE59BD9B4E019ULL * stream) * 0x9E3779B97F4A7C15ULL;
	return mix64(base + i);
}

template <typename K>
__global__ void fill_kernel(K *dst, size_t count, unsigned long long seed, unsigned long long start,
                            int dist, unsigned long long mask, unsigned long long orv) {
	for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long i = start + t;
		unsigned long long x;
		switch (dist) {
		case 0: x = stream_word(seed, i, 0); break;
		case 1: case 2: case 3:
			x = stream_word(seed, i, 0);
			for (int s = 1; s <= dist; ++s)
				x &= stream_word(seed, i, s);
			break;
		case 4: {
			const unsigned long long hi = stream_word(seed, i, 0) >> 32;
			const unsigned long long e = hi >> 27, f = hi & ((1ULL << 27) - 1);
			x = (((1ULL << 27) + f) << e) >> 27;
			break;
		}
		case 5: x = i; break;
		case 6: x = ~0ULL - i; break;
		default: x = stream_word(seed, 0, 0); break;
		}
		dst[t] = (K)((x & mask) | orv);
	}
}

template <int ES>
__global__ void verify_kernel(const typename Rec<ES>::type *__restrict__ data, size_t n, KeyDesc kd,
                              unsigned long long *out3) {
	unsigned long long desc = 0, sum = 0, x = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long k = derive_key(key_word<ES>(data[i], kd.word_sel), kd);
		if (i + 1 < n)
			desc += k > derive_key(key_word<ES>(data[i + 1], kd.word_sel), kd);
		const unsigned long long h = mix64(k + 0x9E3779B97F4A7C15ULL);
		sum += h;
		x ^= h;
	}
	for (int o = 16; o; o >>= 1) {
		desc += __shfl_xor_sync(0xFFFFFFFFu, desc, o);
		sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
		x ^= __shfl_xor_sync(0xFFFFFFFFu, x, o);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicAdd(&out3[0], desc);
		atomicAdd(&out3[1], sum);
		atomicXor(&out3[2], x);
	}
}

// Key-range routing: counts[j] = number of records whose derived key falls in range j, where the
// range index is the number of splitters <= key (splitters ascending).  One read of the records.
// The splitters arrive as a kernel argument (constant bank: compares against uniform operands, no
// shared-memory loads per key); unused entries are ~0 so that no key reaches them.
struct SplitTable {
	unsigned long long s[16];
};
template <int ES>
__global__ void __launch_bounds__(512) split_counts_kernel(const typename Rec<ES>::type *__restrict__ data, size_t n, KeyDesc kd,
                                                           SplitTable split, uint32_t nsplit, unsigned long long *counts) {
	__shared__ unsigned int s_cnt[16];
	if (threadIdx.x < 16)
		s_cnt[threadIdx.x] = 0;
	__syncthreads();
	uint32_t local[16];
#pragma unroll
	for (int j = 0; j < 16; ++j)
		local[j] = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long k = derive_key(key_word<ES>(data[i], kd.word_sel), kd);
		uint32_t d = 0;
#pragma unroll
		for (int j = 0; j < 15; ++j)
			if (j < (int)nsplit) // uniform
				d += k >= split.s[j];
#pragma unroll
		for (int j = 0; j < 16; ++j)
			if (j <= (int)nsplit) // uniform
				local[j] += (d == (uint32_t)j);
	}
#pragma unroll
	for (int j = 0; j < 16; ++j) {
		const uint32_t v = __reduce_add_sync(0xFFFFFFFFu, local[j]);
		if ((threadIdx.x & 31) == 0 && v)
			atomicAdd(&s_cnt[j], v);
	}
	__syncthreads();
	if (threadIdx.x < 16 && s_cnt[threadIdx.x])
		atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// Hardware probe behind RANK_TICKET (rsx_scatter.cuh): are same-address shared-memory atomicAdd
// tickets of one warp instruction handed out in ascending lane order, and a warp's back-to-back
// atomics applied in program order?  Compared against the ballot-derived stable rank.
// The atomic is issued exactly the way the production kernels issue it (ptxas turns every one of
// these into a -- possibly predicated -- ATOMS.POPC.INC, check with cuobjdump -sass):
//   MODE 0  all 32 lanes, convergent
//   MODE 1  scatter_kernel: lanes holding the tile's "hot" digit are ranked by a vote and skip
//           the atomic (`if (is_hot) ... else atomicAdd`), a different lane subset per item
//   MODE 2  small_sort_kernel: `valid ? atomicAdd(..) : 0` with a ragged tail of invalid lanes
template <int MODE>
__global__ void __launch_bounds__(512) ticket_probe_kernel(unsigned long long *mismatch, int iters, uint32_t digit_mask) {
	constexpr int ITEMS = 16;
	__shared__ uint32_t wh[16][kBins];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t lt = lanemask_lt();
	unsigned long long bad = 0;
	for (int it = 0; it < iters; ++it) {
		for (int b = lane; b < kBins; b += 32)
			wh[warp][b] = 0;
		__syncwarp();
		uint32_t d[ITEMS], ticket[ITEMS];
#pragma unroll
		for (int i = 0; i < ITEMS; ++i)
			d[i] = (uint32_t)(mix64(((unsigned long long)blockIdx.x << 40) + ((unsigned long long)it << 20) + threadIdx.x * 64 + i) >> 24) & digit_mask;
		// which lanes take part in item i's atomic
		const uint32_t hot = __shfl_sync(0xFFFFFFFFu, d[0], it & 31); // some digit that occurs
		const uint32_t nvalid = 1u + (uint32_t)(mix64((unsigned long long)it * 977u + blockIdx.x) % (ITEMS * 32u));
		uint32_t hotcnt = 0;
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			if constexpr (MODE == 0) {
				ticket[i] = atomicAdd(&wh[warp][d[i]], 1u);
			} else if constexpr (MODE == 1) {
				const bool is_hot = d[i] == hot;
				const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_hot);
				if (is_hot)
					ticket[i] = hotcnt + __popc(m & lt);
				else
					ticket[i] = atomicAdd(&wh[warp][d[i]], 1u);
				hotcnt += __popc(m);
			} else {
				const bool valid = (uint32_t)i * 32u + lane < nvalid;
				ticket[i] = valid ? atomicAdd(&wh[warp][d[i]], 1u) : 0u;
			}
		}
		__syncwarp();
		for (int b = lane; b < kBins; b += 32)
			wh[warp][b] = 0;
		__syncwarp();
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			const bool valid = MODE != 2 || (uint32_t)i * 32u + lane < nvalid;
			uint32_t peers = __ballot_sync(0xFFFFFFFFu, valid);
#pragma unroll
			for (int b = 0; b < 8; ++b) {
				const bool bit = (d[i] >> b) & 1u;
				const uint32_t v = __ballot_sync(0xFFFFFFFFu, bit);
				peers &= bit ? v : ~v;
			}
			const uint32_t leader = __ffs(peers) - 1;
			uint32_t old = 0;
			if (valid && lane == leader) {
				old = wh[warp][d[i]];
				wh[warp][d[i]] = old + __popc(peers);
			}
			__syncwarp();
			old = __shfl_sync(0xFFFFFFFFu, old, valid ? leader : lane);
			bad += valid && ticket[i] != old + __popc(peers & lt);
		}
		__syncwarp();
	}
	if (bad)
		atomicAdd(mismatch, bad);
}

// ---- records of any size (12-byte {u32 key, 64-bit payload} on ILP32, 24-byte rows, ...) --------
// The tile kernels move records of 1/2/4/8/16 bytes.  Every other size -- and keys that straddle
// an 8-byte word -- is sorted as: extract the keys, rank-sort them (keys + indices travel through
// the passes, the records stay put), then ONE gather moves each record to its final place.
template <typename K>
__global__ void extract_keys_kernel(const unsigned char *__restrict__ recs, size_t n, uint32_t rb, uint32_t ko, K *__restrict__ out) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned char *p = recs + i * rb + ko;
		K k = 0;
#pragma unroll
		for (int b = 0; b < (int)sizeof(K); ++b)
			k |= (K)p[b] << (8 * b);
		out[i] = k;
	}
}

// dst[i] = src[rank[i]], W-byte words (W = 4 when record size and both buffers allow it, else 1)
template <typename I, typename W>
__global__ void gather_records_kernel(const W *__restrict__ src, const I *__restrict__ rank, W *__restrict__ dst, size_t n,
                                      uint32_t words_per_rec) {
	const size_t total = n * words_per_rec;
	for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (size_t)gridDim.x * blockDim.x) {
		const size_t i = w / words_per_rec;
		const uint32_t j = (uint32_t)(w - i * words_per_rec);
		dst[w] = src[(size_t)rank[i] * words_per_rec + j];
	}
}

// out[i] = derived key of record i * stride (key-range routing: the splitters are quantiles of a sample)
template <int ES>
__global__ void sample_keys_kernel(const typename Rec<ES>::type *__restrict__ data, size_t count, size_t stride, KeyDesc kd,
                                   unsigned long long *out) {
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < count)
		out[i] = derive_key(key_word<ES>(data[i * stride], kd.word_sel), kd);
}

// Digit histogram of column `col` over every stride-th record (append-mode routing estimate).
template <int ES>
__global__ void sample_column_hist_kernel(const typename Rec<ES>::type *__restrict__ data, size_t n, KeyDesc kd, uint32_t col,
                                          size_t stride, unsigned long long *hist) {
	__shared__ unsigned int s_h[kBins];
	for (int b = threadIdx.x; b < kBins; b += blockDim.x)
		s_h[b] = 0;
	__syncthreads();
	const size_t m = (n + stride - 1) / stride;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned long long k = derive_key(key_word<ES>(data[i * stride], kd.word_sel), kd);
		atomicAdd(&s_h[(uint32_t)(k >> (8u * col)) & 0xFFu], 1u);
	}
	__syncthreads();
	for (int b = threadIdx.x; b < kBins; b += blockDim.x)
		if (s_h[b])
			atomicAdd(&hist[b], (unsigned long long)s_h[b]);
}

inline int grid_for(size_t n, int threads, int cap) {
	size_t g = (n + threads - 1) / threads;
	if (g < 1) g = 1;
	return (int)(g > (size_t)cap ? (size_t)cap : g);
}

} // namespace

cudaError_t launch_split_counts(const void *data, size_t n, uint32_t record_bytes, const KeyDesc &kd,
                                const unsigned long long *h_split, uint32_t nsplit, unsigned long long *d_counts,
                                int num_sms, cudaStream_t st) {
	const int g = grid_for(n, 512, num_sms * 4);
	SplitTable t;
	for (uint32_t j = 0; j < 16; ++j)
		t.s[j] = j < nsplit ? h_split[j] : ~0ULL;
	switch (record_bytes) {
	case 1: split_counts_kernel<1><<<g, 512, 0, st>>>(static_cast<const uint8_t *>(data), n, kd, t, nsplit, d_counts); break;
	case 2: split_counts_kernel<2><<<g, 512, 0, st>>>(static_cast<const uint16_t *>(data), n, kd, t, nsplit, d_counts); break;
	case 4: split_counts_kernel<4><<<g, 512, 0, st>>>(static_cast<const uint32_t *>(data), n, kd, t, nsplit, d_counts); break;
	case 8: split_counts_kernel<8><<<g, 512, 0, st>>>(static_cast<const unsigned long long *>(data), n, kd, t, nsplit, d_counts); break;
	case 16: split_counts_kernel<16><<<g, 512, 0, st>>>(static_cast<const ulonglong2 *>(data), n, kd, t, nsplit, d_counts); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_ticket_probe(unsigned long long *d_mismatch, int num_sms, cudaStream_t st) {
	// ~5 M tickets per launch, on every SM (0.3 ms in total, once per device); the offline probe
	// (tools/probe_atoms.cu) runs billions
	for (uint32_t mask : {0xFFu, 0x0Fu, 0x01u}) {
		ticket_probe_kernel<0><<<num_sms, 512, 0, st>>>(d_mismatch, 4, mask);
		ticket_probe_kernel<1><<<num_sms, 512, 0, st>>>(d_mismatch, 4, mask);
		ticket_probe_kernel<2><<<num_sms, 512, 0, st>>>(d_mismatch, 4, mask);
	}
	count_launch(9);
	return cudaGetLastError();
}

cudaError_t launch_iota_if_early(void *ib, int idx_bytes, size_t n, const Ctl *ctl, cudaStream_t st) {
	const int g = grid_for(n, 256, 148 * 8);
	switch (idx_bytes) {
	case 1: iota_if_early_kernel<<<g, 256, 0, st>>>(static_cast<uint8_t *>(ib), n, ctl); break;
	case 2: iota_if_early_kernel<<<g, 256, 0, st>>>(static_cast<uint16_t *>(ib), n, ctl); break;
	case 4: iota_if_early_kernel<<<g, 256, 0, st>>>(static_cast<uint32_t *>(ib), n, ctl); break;
	case 8: iota_if_early_kernel<<<g, 256, 0, st>>>(static_cast<unsigned long long *>(ib), n, ctl); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_narrow_index(const uint32_t *w0, const uint32_t *w1, void *ib, int idx_bytes, size_t n,
                                const Ctl *ctl, cudaStream_t st) {
	const int g = grid_for(n, 256, 148 * 8);
	switch (idx_bytes) {
	case 1: narrow_index_kernel<<<g, 256, 0, st>>>(w0, w1, static_cast<uint8_t *>(ib), n, ctl); break;
	case 2: narrow_index_kernel<<<g, 256, 0, st>>>(w0, w1, static_cast<uint16_t *>(ib), n, ctl); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_fill(void *dst, size_t count, int key_bytes, uint64_t seed, uint64_t start, int dist,
                        uint64_t mask, uint64_t orv, cudaStream_t st) {
	const int g = grid_for(count, 256, 148 * 16);
	switch (key_bytes) {
	case 1: fill_kernel<<<g, 256, 0, st>>>(static_cast<uint8_t *>(dst), count, seed, start, dist, mask, orv); break;
	case 2: fill_kernel<<<g, 256, 0, st>>>(static_cast<uint16_t *>(dst), count, seed, start, dist, mask, orv); break;
	case 4: fill_kernel<<<g, 256, 0, st>>>(static_cast<uint32_t *>(dst), count, seed, start, dist, mask, orv); break;
	case 8: fill_kernel<<<g, 256, 0, st>>>(static_cast<unsigned long long *>(dst), count, seed, start, dist, mask, orv); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_verify(const void *data, size_t n, uint32_t record_bytes, const KeyDesc &kd,
                          unsigned long long *out3, int num_sms, cudaStream_t st) {
	const int g = grid_for(n, 256, num_sms * 8);
	switch (record_bytes) {
	case 1: verify_kernel<1><<<g, 256, 0, st>>>(static_cast<const uint8_t *>(data), n, kd, out3); break;
	case 2: verify_kernel<2><<<g, 256, 0, st>>>(static_cast<const uint16_t *>(data), n, kd, out3); break;
	case 4: verify_kernel<4><<<g, 256, 0, st>>>(static_cast<const uint32_t *>(data), n, kd, out3); break;
	case 8: verify_kernel<8><<<g, 256, 0, st>>>(static_cast<const unsigned long long *>(data), n, kd, out3); break;
	case 16: verify_kernel<16><<<g, 256, 0, st>>>(static_cast<const ulonglong2 *>(data), n, kd, out3); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

} // namespace rsx

namespace rsx {
cudaError_t launch_extract_keys(const void *recs, size_t n, uint32_t record_bytes, uint32_t key_offset, uint32_t key_bytes,
                                void *keys_out, int num_sms, cudaStream_t st) {
	const int g = grid_for(n, 256, num_sms * 16);
	const unsigned char *r = static_cast<const unsigned char *>(recs);
	switch (key_bytes) {
	case 1: extract_keys_kernel<<<g, 256, 0, st>>>(r, n, record_bytes, key_offset, static_cast<uint8_t *>(keys_out)); break;
	case 2: extract_keys_kernel<<<g, 256, 0, st>>>(r, n, record_bytes, key_offset, static_cast<uint16_t *>(keys_out)); break;
	case 4: extract_keys_kernel<<<g, 256, 0, st>>>(r, n, record_bytes, key_offset, static_cast<uint32_t *>(keys_out)); break;
	case 8: extract_keys_kernel<<<g, 256, 0, st>>>(r, n, record_bytes, key_offset, static_cast<unsigned long long *>(keys_out)); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_sample_keys(const void *data, size_t count, size_t stride, uint32_t record_bytes, const KeyDesc &kd,
                               unsigned long long *d_out, cudaStream_t st) {
	const int g = (int)((count + 255) / 256);
	switch (record_bytes) {
	case 1: sample_keys_kernel<1><<<g, 256, 0, st>>>(static_cast<const uint8_t *>(data), count, stride, kd, d_out); break;
	case 2: sample_keys_kernel<2><<<g, 256, 0, st>>>(static_cast<const uint16_t *>(data), count, stride, kd, d_out); break;
	case 4: sample_keys_kernel<4><<<g, 256, 0, st>>>(static_cast<const uint32_t *>(data), count, stride, kd, d_out); break;
	case 8: sample_keys_kernel<8><<<g, 256, 0, st>>>(static_cast<const unsigned long long *>(data), count, stride, kd, d_out); break;
	case 16: sample_keys_kernel<16><<<g, 256, 0, st>>>(static_cast<const ulonglong2 *>(data), count, stride, kd, d_out); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_sample_column_hist(const void *data, size_t n, uint32_t record_bytes, const KeyDesc &kd, int col, size_t stride,
                                      unsigned long long *d_hist, int num_sms, cudaStream_t st) {
	const int g = grid_for((n + stride - 1) / stride, 256, num_sms * 8);
	switch (record_bytes) {
	case 1: sample_column_hist_kernel<1><<<g, 256, 0, st>>>(static_cast<const uint8_t *>(data), n, kd, (uint32_t)col, stride, d_hist); break;
	case 2: sample_column_hist_kernel<2><<<g, 256, 0, st>>>(static_cast<const uint16_t *>(data), n, kd, (uint32_t)col, stride, d_hist); break;
	case 4: sample_column_hist_kernel<4><<<g, 256, 0, st>>>(static_cast<const uint32_t *>(data), n, kd, (uint32_t)col, stride, d_hist); break;
	case 8: sample_column_hist_kernel<8><<<g, 256, 0, st>>>(static_cast<const unsigned long long *>(data), n, kd, (uint32_t)col, stride, d_hist); break;
	case 16: sample_column_hist_kernel<16><<<g, 256, 0, st>>>(static_cast<const ulonglong2 *>(data), n, kd, (uint32_t)col, stride, d_hist); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_gather_records(const void *src, const void *rank, int idx_bytes, void *dst, size_t n, uint32_t record_bytes,
                                  int num_sms, cudaStream_t st) {
	const bool words = record_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
	const uint32_t wpr = words ? record_bytes / 4 : record_bytes;
	const int g = grid_for(n * wpr, 256, num_sms * 16);
	if (words) {
		if (idx_bytes == 4)
			gather_records_kernel<<<g, 256, 0, st>>>(static_cast<const uint32_t *>(src), static_cast<const uint32_t *>(rank), static_cast<uint32_t *>(dst), n, wpr);
		else
			gather_records_kernel<<<g, 256, 0, st>>>(static_cast<const uint32_t *>(src), static_cast<const unsigned long long *>(rank), static_cast<uint32_t *>(dst), n, wpr);
	} else {
		if (idx_bytes == 4)
			gather_records_kernel<<<g, 256, 0, st>>>(static_cast<const uint8_t *>(src), static_cast<const uint32_t *>(rank), static_cast<uint8_t *>(dst), n, wpr);
		else
			gather_records_kernel<<<g, 256, 0, st>>>(static_cast<const uint8_t *>(src), static_cast<const unsigned long long *>(rank), static_cast<uint8_t *>(dst), n, wpr);
	}
	count_launch();
	return cudaGetLastError();
}
} // namespace rsx
