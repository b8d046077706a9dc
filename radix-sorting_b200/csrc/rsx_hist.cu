// rsx_hist.cu -- K1 fused upfront histogram + K2 setup.
//
// K1 replaces the reference's histogram loop (radix_sort.hpp:48-58, radix_sort_rank.hpp:42-53):
// ONE read of the records produces every column's 256-bin digit histogram and the pre-sorted
// verdict.  K2 replaces radix_sort.hpp:60-80: early-exit test, trivial-column probe with the
// first key, exclusive scans -- all on the device, so the host never waits to learn `cols[]`.
//
// K1 design (HBM-bound: algorithmic bytes = n * record_bytes, read once):
//   * 128-bit coalesced loads, 4 in flight per thread, 1024 threads per CTA, one CTA per SM.
//   * Shared-memory histograms are LANE-PRIVATE: counter (column c, digit d) has 32 copies, one
//     per lane id, laid out so that copy l lives in bank l.  A warp's 32 increments therefore
//     never collide in a bank, whatever the digit distribution (a constant column -- the
//     worst case for a shared counter -- is as fast as a uniform one), and no per-key
//     warp-match is needed: the warp-level aggregation happens once per flush, when the 32
//     copies of a bin are summed with a warp reduction and added to the global histogram with
//     one 64-bit atomic per bin.
//   * <= 4 columns: 32-bit copies (4 x 256 x 32 x 4 B = 128 KiB).  8 columns: 16-bit copies
//     packed two per word (128 KiB), flushed before a copy can reach 65535.
//   * descents (kdf(a[i]) > kdf(a[i+1])) are counted on the fly: inside a thread's vector,
//     across lanes with a shuffle, across warps with one extra scalar load by lane 31.
#include <algorithm>
#include <type_traits>

#include "rsx_device.cuh"

namespace rsx {

namespace {

constexpr int kHistThreads = 1024;
constexpr int kHistUnroll = 4;

template <int KB> struct HistSmem {
	static constexpr bool kPacked = KB > 4;
	static constexpr int kWords = KB * kBins * (kPacked ? 16 : 32);
	static constexpr size_t kBytes = (size_t)kWords * 4;
};

// Key derivation with everything data-independent folded into three uniform constants:
//   derived = raw ^ xor_const ^ (sign(raw) ? float_mask : 0)
// unsigned: 0 / 0, signed: top / 0, float: top / (mask ^ top); descending adds mask to xor_const.
template <typename KT> struct KeyXform {
	uint32_t word_sel, shift;
	KT mask, xor_const, float_mask;
	uint32_t sign_shift;
};
template <typename KT> __host__ KeyXform<KT> make_xform(const KeyDesc &kd) {
	KeyXform<KT> x;
	x.word_sel = kd.word_sel;
	x.shift = kd.key_shift;
	const unsigned long long m = kd.key_bytes >= 8 ? ~0ULL : ((1ULL << (8 * kd.key_bytes)) - 1ULL);
	const unsigned long long top = 1ULL << (8 * kd.key_bytes - 1);
	x.mask = (KT)m;
	x.xor_const = (KT)((kd.kdf_kind != RSX_KDF_UNSIGNED ? top : 0ULL) ^ (kd.invert ? m : 0ULL));
	x.float_mask = (KT)(kd.kdf_kind == RSX_KDF_FLOAT ? (m ^ top) : 0ULL);
	x.sign_shift = 8 * kd.key_bytes - 1;
	return x;
}
template <int ES, typename KT>
__device__ __forceinline__ KT derive_fast(const typename Rec<ES>::type &r, const KeyXform<KT> &x) {
	KT raw;
	if constexpr (ES <= 4)
		raw = (KT)((uint32_t)r >> x.shift);
	else
		raw = (KT)(key_word<ES>(r, x.word_sel) >> x.shift);
	raw &= x.mask;
	KT k = raw ^ x.xor_const;
	if (x.float_mask != 0) // uniform
		k ^= ((KT)0 - (KT)((raw >> x.sign_shift) & 1u)) & x.float_mask;
	return k;
}

template <typename KT, int NR> __device__ __forceinline__ KT compact_key(KT k, const Compaction &c) {
	KT out = 0;
#pragma unroll // compile-time indices: the run table stays in the constant bank (a runtime index spills it to local memory)
	for (int i = 0; i < NR; ++i)
		out |= ((k >> c.src_shift[i]) & (KT)c.wmask[i]) << c.dst_shift[i]; // unused runs have wmask 0
	return out;
}

// One increment per column.  `lane_base` already contains the lane's bank offset, so each
// digit costs a shift, a mask, an add and the shared atomic.
template <int KB, typename KT>
__device__ __forceinline__ void hist_one(unsigned char *lane_base, KT key, uint32_t inc, int cols = KB) {
	constexpr bool kPacked = HistSmem<KB>::kPacked;
	constexpr uint32_t kColBytes = kPacked ? kBins * 64u : kBins * 128u; // bytes per column
	constexpr uint32_t kSh = kPacked ? 6u : 7u;                          // bytes per bin = 1 << kSh
	constexpr uint32_t kMaskOff = 0xFFu << kSh;
#pragma unroll
	for (int c = 0; c < KB; ++c) {
		uint32_t half = (uint32_t)key;
		if constexpr (sizeof(KT) == 8) {
			if (c >= 4)
				half = (uint32_t)(key >> 32);
		}
		const int cc = c & 3;
		const uint32_t off = cc == 0 ? (half << kSh) & kMaskOff : (half >> (8 * cc - kSh)) & kMaskOff;
		if (c < cols) // compacted keys: the columns above the compacted width are all zero, counted wholesale
			atomicAdd(reinterpret_cast<uint32_t *>(lane_base + c * kColBytes + off), inc);
	}
}

// Sum the 32 lane-private copies of every bin, add to the global histogram, zero the copies.
template <int KB>
__device__ __forceinline__ void hist_flush(uint32_t *sh, unsigned long long *ghist) {
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
	for (uint32_t bin = warp; bin < KB * kBins; bin += nwarps) {
		uint32_t v;
		if constexpr (HistSmem<KB>::kPacked) {
			uint32_t w = 0;
			if (lane < 16) {
				w = sh[bin * 16u + lane];
				sh[bin * 16u + lane] = 0;
			}
			v = (w & 0xFFFFu) + (w >> 16);
		} else {
			v = sh[bin * 32u + lane];
			sh[bin * 32u + lane] = 0;
		}
		v = __reduce_add_sync(0xFFFFFFFFu, v);
		if (lane == 0 && v)
			atomicAdd(&ghist[bin], (unsigned long long)v);
	}
}

// COMPACT: 0 = off, else the number of runs the software PEXT is unrolled for (2, 4 or 8)
template <int ES, int KB, int COMPACT>
__global__ void __launch_bounds__(kHistThreads, 1)
histogram_kernel(const typename Rec<ES>::type *__restrict__ src, size_t n, size_t head, size_t n_vec,
                 KeyXform<std::conditional_t<(KB > 4), unsigned long long, uint32_t>> xf,
                 unsigned long long *__restrict__ ghist, unsigned long long *__restrict__ gdescents,
                 unsigned long long *__restrict__ gor, const Compaction cmp, typename Rec<ES>::type *__restrict__ cout) {
	using R = typename Rec<ES>::type;
	using KT = std::conditional_t<(KB > 4), unsigned long long, uint32_t>;
	constexpr int VEC = 16 / ES;
	extern __shared__ __align__(16) uint32_t sh[];
	__shared__ uint32_t s_desc;

	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	for (int i = tid; i < HistSmem<KB>::kWords; i += kHistThreads)
		sh[i] = 0;
	if (tid == 0)
		s_desc = 0;
	__syncthreads();
	// lane-private copy: 32-bit copies -> word `lane` of each bin; 16-bit copies -> half (lane / 16)
	// of word (lane % 16)
	unsigned char *lane_base = reinterpret_cast<unsigned char *>(sh) +
	                           (HistSmem<KB>::kPacked ? (lane & 15u) * 4u : lane * 4u);
	const uint32_t inc = HistSmem<KB>::kPacked ? (1u << (lane & 16u)) : 1u;

	uint32_t descents = 0;
	const int hcols = COMPACT ? (int)(cmp.bits + 7) / 8 : KB;
	const bool vec_out = COMPACT != 0 && ES <= 8 && (reinterpret_cast<uintptr_t>(cout + head) & 15) == 0;
	KT acc_or = 0, acc_nand = 0; // which bits of the derived key vary over the input (key compaction)
	auto derive = [&](const R &r) -> KT {
		KT k = derive_fast<ES, KT>(r, xf);
		if constexpr (COMPACT)
			k = compact_key<KT, COMPACT>(k, cmp);
		return k;
	};
	const uint4 *vsrc = reinterpret_cast<const uint4 *>(src + head);
	const size_t stride = (size_t)gridDim.x * kHistThreads;
	// CTA-uniform trip count so that the flush barrier is reached by every thread.
	const size_t per_iter = stride * kHistUnroll;
	const size_t iters = (n_vec + per_iter - 1) / per_iter;
	// packed 16-bit copies: a lane copy gains at most (warps * VEC * unroll) per iteration
	constexpr uint32_t kFlushEvery =
		HistSmem<KB>::kPacked ? 65535u / ((kHistThreads / 32) * VEC * kHistUnroll) : 0xFFFFFFFFu;
	uint32_t since_flush = 0;

	for (size_t it = 0; it < iters; ++it) {
		const size_t v0 = it * per_iter + (size_t)blockIdx.x * kHistThreads + tid;
		uint4 q[kHistUnroll];
#pragma unroll
		for (int u = 0; u < kHistUnroll; ++u) {
			const size_t v = v0 + (size_t)u * stride;
			q[u] = v < n_vec ? __ldg(vsrc + v) : make_uint4(0, 0, 0, 0);
		}
#pragma unroll
		for (int u = 0; u < kHistUnroll; ++u) {
			const size_t v = v0 + (size_t)u * stride;
			const bool live = v < n_vec;
			KT k[VEC];
			const R *e = reinterpret_cast<const R *>(&q[u]);
#pragma unroll
			for (int j = 0; j < VEC; ++j)
				k[j] = derive(e[j]);
			// successor of this vector's last record: next lane's first key, or a scalar load
			KT nxt = __shfl_down_sync(0xFFFFFFFFu, k[0], 1);
			const size_t next_idx = head + (v + 1) * VEC;
			const bool have_next = live && next_idx < n;
			if ((lane == 31 || v + 1 >= n_vec) && have_next)
				nxt = derive(src[next_idx]);
			if (live) {
#pragma unroll
				for (int j = 0; j < VEC; ++j) {
					hist_one<KB, KT>(lane_base, k[j], inc, hcols);
					if constexpr (!COMPACT) {
						acc_or |= k[j];
						acc_nand |= ~k[j];
					}
				}
				if constexpr (COMPACT != 0 && ES <= 8) {
					if (vec_out) { // one 16-byte store per input vector
						R o[VEC];
#pragma unroll
						for (int j = 0; j < VEC; ++j)
							o[j] = (R)k[j];
						reinterpret_cast<uint4 *>(cout + head)[v] = *reinterpret_cast<const uint4 *>(o);
					} else {
#pragma unroll
						for (int j = 0; j < VEC; ++j)
							cout[head + v * VEC + j] = (R)k[j];
					}
				}
#pragma unroll
				for (int j = 0; j + 1 < VEC; ++j)
					descents += k[j] > k[j + 1];
				if (have_next)
					descents += k[VEC - 1] > nxt;
			}
		}
		if constexpr (HistSmem<KB>::kPacked) {
			if (++since_flush == kFlushEvery) {
				__syncthreads();
				hist_flush<KB>(sh, ghist);
				__syncthreads();
				since_flush = 0;
			}
		}
	}

	// unaligned head and the tail that does not fill a vector: CTA 0, one record per thread
	if (blockIdx.x == 0) {
		const size_t tail0 = head + n_vec * VEC;
		const size_t extra = head + (n - tail0);
		for (size_t t = tid; t < extra; t += kHistThreads) {
			const size_t i = t < head ? t : tail0 + (t - head);
			const KT k = derive(src[i]);
			hist_one<KB, KT>(lane_base, k, inc, hcols);
			acc_or |= k;
			acc_nand |= ~k;
			if constexpr (COMPACT != 0 && ES <= 8)
				cout[i] = (R)k;
			if (i + 1 < n)
				descents += k > derive(src[i + 1]);
		}
	}

	descents = __reduce_add_sync(0xFFFFFFFFu, descents);
	if (lane == 0 && descents)
		atomicAdd(&s_desc, descents);
	if constexpr (!COMPACT) {
		unsigned long long o = (unsigned long long)acc_or, na = (unsigned long long)(KT)acc_nand;
		for (int s_ = 16; s_; s_ >>= 1) {
			o |= __shfl_xor_sync(0xFFFFFFFFu, o, s_);
			na |= __shfl_xor_sync(0xFFFFFFFFu, na, s_);
		}
		if (lane == 0) {
			atomicOr(&gor[0], o);
			atomicOr(&gor[1], na);
		}
	}
	__syncthreads();
	hist_flush<KB>(sh, ghist);
	if (tid == 0 && s_desc)
		atomicAdd(gdescents, (unsigned long long)s_desc);
	if constexpr (COMPACT) {
		if (blockIdx.x == 0 && tid < KB && (int)tid >= hcols)
			atomicAdd(&ghist[tid * kBins], (unsigned long long)n); // every key has digit 0 in this column
	}
}

// K2: one CTA, 256 threads; thread d owns bin d of every column.
template <int ES>
__global__ void __launch_bounds__(kBins, 1)
setup_kernel(const typename Rec<ES>::type *__restrict__ src, size_t n, KeyDesc kd, WsHead *ws, Ctl *host_ctl) {
	__shared__ unsigned long long s_warp[8];
	__shared__ uint32_t s_live[kMaxCols];
	const uint32_t d = threadIdx.x, lane = d & 31u, warp = d >> 5;
	// radix_sort.hpp:65: sample the first key
	const unsigned long long key0 = derive_key(key_word<ES>(src[0], kd.word_sel), kd);
	if (d < kd.key_bytes)
		s_live[d] = ws->hist[d * kBins + ((key0 >> (8 * d)) & 0xFFu)] != n; // radix_sort.hpp:66-69
	__syncthreads();
	for (uint32_t c = 0; c < kd.key_bytes; ++c) {
		// exclusive scan of column c (radix_sort.hpp:73-80); computed for every column, only the
		// live ones are consumed.
		const unsigned long long v = ws->hist[c * kBins + d];
		unsigned long long x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o);
			if (lane >= o)
				x += y;
		}
		if (lane == 31)
			s_warp[warp] = x;
		__syncthreads();
		unsigned long long base = 0;
		for (uint32_t w = 0; w < warp; ++w)
			base += s_warp[w];
		ws->offs[c * kBins + d] = base + x - v;
		__syncthreads();
	}
	if (d == 0) {
		Ctl ctl;
		ctl.early_exit = ws->descents == 0; // <=> n_unsorted < 2, radix_sort.hpp:60-62
		ctl.ncols = 0;
		ctl.live_mask = 0;
		for (uint32_t c = 0; c < kMaxCols; ++c) {
			ctl.ordinal[c] = ctl.ncols;
			if (c < kd.key_bytes && s_live[c]) {
				ctl.live_mask |= 1u << c;
				++ctl.ncols;
			}
		}
		ctl.pad = 0;
		ctl.n = n;
		ctl.key_or = ws->key_or;
		ctl.key_nand = ws->key_nand & width_mask(kd.key_bytes);
		ws->ctl = ctl;
		if (host_ctl != nullptr)
			*host_ctl = ctl; // mapped pinned memory: the host reads it after the stream has drained
	}
}

template <int ES, int KB, int COMPACT>
cudaError_t launch_hist_t(const void *src, size_t n, const KeyDesc &kd, WsHead *ws, int num_sms,
                          cudaStream_t st, const Compaction *cmp, void *cout) {
	using R = typename Rec<ES>::type;
	constexpr int VEC = 16 / ES;
	auto kern = histogram_kernel<ES, KB, COMPACT>;
	static bool configured[64] = {}; // per device; benign race: setting the attribute is idempotent
	int dev = 0;
	cudaGetDevice(&dev);
	if (!configured[dev & 63]) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
		                                     (int)HistSmem<KB>::kBytes);
		if (e != cudaSuccess)
			return e;
		configured[dev & 63] = true;
	}
	const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
	size_t head = ((16 - (addr & 15)) & 15) / ES;
	if (head > n)
		head = n;
	const size_t n_vec = (n - head) / VEC;
	size_t want = (n_vec + (size_t)kHistThreads * kHistUnroll - 1) / ((size_t)kHistThreads * kHistUnroll);
	int grid = (int)(want < (size_t)num_sms ? (want ? want : 1) : (size_t)num_sms);
	using KT = std::conditional_t<(KB > 4), unsigned long long, uint32_t>;
	kern<<<grid, kHistThreads, HistSmem<KB>::kBytes, st>>>(static_cast<const R *>(src), n, head, n_vec,
	                                                      make_xform<KT>(kd), ws->hist, &ws->descents, &ws->key_or,
	                                                      cmp ? *cmp : Compaction{}, static_cast<R *>(cout));
	count_launch();
	return cudaGetLastError();
}

template <int ES>
cudaError_t launch_hist_es(const void *src, size_t n, const KeyDesc &kd, WsHead *ws, int num_sms,
                           cudaStream_t st, const Compaction *cmp, void *cout) {
	if (cmp != nullptr) { // compaction: keys-only records of 4 or 8 bytes
		if constexpr (ES == 4 || ES == 8) {
			if (kd.key_bytes == ES) {
				if (cmp->nruns <= 2)
					return launch_hist_t<ES, ES, 2>(src, n, kd, ws, num_sms, st, cmp, cout);
				if (cmp->nruns <= 4)
					return launch_hist_t<ES, ES, 4>(src, n, kd, ws, num_sms, st, cmp, cout);
				return launch_hist_t<ES, ES, 8>(src, n, kd, ws, num_sms, st, cmp, cout);
			}
		}
		return cudaErrorInvalidValue;
	}
	switch (kd.key_bytes) {
	case 1: return launch_hist_t<ES, 1, 0>(src, n, kd, ws, num_sms, st, nullptr, nullptr);
	case 2: if constexpr (ES >= 2) return launch_hist_t<ES, 2, 0>(src, n, kd, ws, num_sms, st, nullptr, nullptr); break;
	case 4: if constexpr (ES >= 4) return launch_hist_t<ES, 4, 0>(src, n, kd, ws, num_sms, st, nullptr, nullptr); break;
	case 8: if constexpr (ES >= 8) return launch_hist_t<ES, 8, 0>(src, n, kd, ws, num_sms, st, nullptr, nullptr); break;
	}
	return cudaErrorInvalidValue;
}

// compacted keys -> original keys: PDEP over the runs, constant bits, inverse key derivation.
// 16-byte accesses, two vectors in flight per thread (a 4-byte-per-thread version ran at half the
// HBM rate: too few bytes in flight).
template <typename K> __device__ __forceinline__ K expand_one(K cv, const KeyDesc &kd, const Compaction &c, unsigned long long m,
                                                              unsigned long long top) {
	const unsigned long long v = (unsigned long long)cv;
	unsigned long long k = c.const_bits;
#pragma unroll
	for (int r = 0; r < kMaxRuns; ++r)
		k |= ((v >> c.dst_shift[r]) & c.wmask[r]) << c.src_shift[r]; // unused runs have wmask 0
	// inverse of derive_key (rsx_device.cuh): complement, then undo the sign / float flip
	if (kd.invert)
		k = ~k & m;
	if (kd.kdf_kind == RSX_KDF_SIGNED)
		k ^= top;
	else if (kd.kdf_kind == RSX_KDF_FLOAT)
		k ^= (k & top) ? top : m; // derived top bit set <=> the float was non-negative
	return (K)k;
}

template <typename K>
__global__ void expand_keys_kernel(const K *__restrict__ in, K *__restrict__ out, size_t n, KeyDesc kd, Compaction c) {
	constexpr int VEC = 16 / (int)sizeof(K);
	const unsigned long long m = width_mask(kd.key_bytes), top = 1ULL << (8u * kd.key_bytes - 1u);
	const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
	const size_t nv = aligned ? n / VEC : 0;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	const uint4 *vin = reinterpret_cast<const uint4 *>(in);
	uint4 *vout = reinterpret_cast<uint4 *>(out);
	for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += 2 * stride) {
		const size_t v2 = v + stride;
		uint4 a = vin[v], b = v2 < nv ? vin[v2] : make_uint4(0, 0, 0, 0);
		K *ka = reinterpret_cast<K *>(&a), *kb = reinterpret_cast<K *>(&b);
#pragma unroll
		for (int j = 0; j < VEC; ++j) {
			ka[j] = expand_one<K>(ka[j], kd, c, m, top);
			kb[j] = expand_one<K>(kb[j], kd, c, m, top);
		}
		vout[v] = a;
		if (v2 < nv)
			vout[v2] = b;
	}
	for (size_t i = nv * VEC + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
		out[i] = expand_one<K>(in[i], kd, c, m, top);
}

} // namespace

cudaError_t launch_histogram(const void *src, size_t n, uint32_t record_bytes, const KeyDesc &kd,
                             WsHead *ws, int num_sms, cudaStream_t st, const Compaction *cmp, void *compact_out) {
	switch (record_bytes) {
	case 1: return launch_hist_es<1>(src, n, kd, ws, num_sms, st, cmp, compact_out);
	case 2: return launch_hist_es<2>(src, n, kd, ws, num_sms, st, cmp, compact_out);
	case 4: return launch_hist_es<4>(src, n, kd, ws, num_sms, st, cmp, compact_out);
	case 8: return launch_hist_es<8>(src, n, kd, ws, num_sms, st, cmp, compact_out);
	case 16: return launch_hist_es<16>(src, n, kd, ws, num_sms, st, cmp, compact_out);
	}
	return cudaErrorInvalidValue;
}

cudaError_t launch_expand_keys(const void *in, void *out, size_t n, const KeyDesc &kd, const Compaction &cmp,
                               int num_sms, cudaStream_t st) {
	const int g = (int)std::min<size_t>((n / 8 + 255) / 256 + 1, (size_t)num_sms * 8);
	if (kd.key_bytes == 4)
		expand_keys_kernel<<<g, 256, 0, st>>>(static_cast<const uint32_t *>(in), static_cast<uint32_t *>(out), n, kd, cmp);
	else if (kd.key_bytes == 8)
		expand_keys_kernel<<<g, 256, 0, st>>>(static_cast<const unsigned long long *>(in), static_cast<unsigned long long *>(out), n, kd, cmp);
	else
		return cudaErrorInvalidValue;
	count_launch();
	return cudaGetLastError();
}

cudaError_t launch_setup(const void *src, size_t n, uint32_t record_bytes, const KeyDesc &kd,
                         WsHead *ws, Ctl *host_ctl, cudaStream_t st) {
	switch (record_bytes) {
	case 1: setup_kernel<1><<<1, kBins, 0, st>>>(static_cast<const uint8_t *>(src), n, kd, ws, host_ctl); break;
	case 2: setup_kernel<2><<<1, kBins, 0, st>>>(static_cast<const uint16_t *>(src), n, kd, ws, host_ctl); break;
	case 4: setup_kernel<4><<<1, kBins, 0, st>>>(static_cast<const uint32_t *>(src), n, kd, ws, host_ctl); break;
	case 8: setup_kernel<8><<<1, kBins, 0, st>>>(static_cast<const unsigned long long *>(src), n, kd, ws, host_ctl); break;
	case 16: setup_kernel<16><<<1, kBins, 0, st>>>(static_cast<const ulonglong2 *>(src), n, kd, ws, host_ctl); break;
	default: return cudaErrorInvalidValue;
	}
	count_launch();
	return cudaGetLastError();
}

} // namespace rsx
