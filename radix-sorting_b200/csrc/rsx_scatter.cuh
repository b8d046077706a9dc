// rsx_scatter.cuh -- K3: one stable 8-bit-digit scatter pass ("onesweep" style).
//
// Replaces the reference's sort loop for one live column (radix_sort.hpp:83-88):
//     for j in 0..n: k = src[j]; dst = offsets[digit(kf(k))]++; aux[dst] = k
// The serial `offsets[..]++` is what makes the reference stable; here the same destination
// index is computed in parallel as
//     dst = column_offset[d]                 (exclusive scan of the global histogram, K2)
//         + #records with digit d in earlier tiles      (decoupled look-back, one chain per digit)
//         + #records with digit d earlier in this tile  (warp match ranking + cross-warp prefix)
// which is exactly the value the reference's counter would have had, so the output is
// bit-identical, including the order of equal keys (stability) and of payloads.
//
// Per pass the algorithmic HBM traffic is n * (record + payload) read + the same written;
// the look-back state adds 256 words written + ~256 read per tile (L2 resident).
//
// Tile pipeline (persistent CTAs, tiles handed out by an atomic ticket so that a tile's
// predecessors are always owned by running CTAs -> the look-back cannot deadlock):
//   1. coalesced warp-striped load of the tile into registers; KDF folded into digit_of()
//   2. per warp, per item: stable rank of the record among the warp's records with the same
//      digit.  Two implementations (template RANK):
//        RANK_TICKET  rank = atomicAdd(&warp_counter[digit], 1): ONE shared-memory instruction
//                     per key.  Stable because this hardware hands out same-address tickets of
//                     one warp instruction in ascending lane order and applies a warp's
//                     back-to-back atomics in program order -- verified per device at first use
//                     by ticket_probe_kernel (and offline by tools/probe_atoms.cu); if the probe
//                     ever disagreed the library would use RANK_BALLOT.
//        RANK_BALLOT  8 votes -> peer mask; the group leader bumps the warp counter and
//                     broadcasts the old value.  Provably stable, ~2.5x cheaper on B200 than
//                     __match_any_sync on 8 random bits (tools/yardstick.cu: 24 vs 60 clk/SM per
//                     32 keys), but still above the ~13 clk/SM budget of a 70 %-of-peak pass.
//   3. threads 0..255 (one per digit): sum / prefix the warp counters, scan the 256 tile
//      counts, publish the tile aggregate, later walk back over predecessor tiles
//   4. records (and payloads) are written to shared memory at their tile-sorted position
//   5. thread t stores shared slot t, t+T, ...: consecutive threads hit consecutive addresses
//      inside each digit bucket, so global writes coalesce per bucket
#pragma once

#include "rsx_device.cuh"

namespace rsx {

struct ScatterParams {
	PassBuffers pb;
	size_t n;
	uint32_t num_tiles;
	uint32_t col;
	DigitDesc dd;
	const unsigned long long *offs; // this column's exclusive scan (256 entries)
	const Ctl *ctl;                 // nullptr: forced pass
	void *status;                   // OffT[num_tiles][256]
	unsigned int *ticket;
	ulonglong2 pad_rec;             // record whose derived key is all ones (tail padding)
};

// RANK_TICKET or RANK_BALLOT for the current device (probe result or rsx_set_option override).
int rank_mode();

template <typename OffT> struct StatusBits;
template <> struct StatusBits<uint32_t> {
	static constexpr uint32_t kAgg = 1u << 30, kPfx = 2u << 30, kMask = (1u << 30) - 1u;
};
template <> struct StatusBits<unsigned long long> {
	static constexpr unsigned long long kAgg = 1ULL << 62, kPfx = 2ULL << 62, kMask = (1ULL << 62) - 1ULL;
};

template <typename T> __device__ __forceinline__ T ld_status(const T *p) {
	return *reinterpret_cast<const volatile T *>(p);
}
template <typename T> __device__ __forceinline__ void st_status(T *p, T v) {
	*reinterpret_cast<volatile T *>(p) = v;
}

template <int ES> __device__ __forceinline__ typename Rec<ES>::type make_pad(const ulonglong2 &p) {
	if constexpr (ES == 16)
		return p;
	else
		return (typename Rec<ES>::type)p.x;
}

template <int ES, int PL, int THREADS, int ITEMS> struct ScatterSmem {
	static constexpr int kTile = THREADS * ITEMS;
	static constexpr int kWarps = THREADS / 32;
	static constexpr size_t kRecBytes = (size_t)kTile * ES;
	static constexpr size_t kPlBytes = (size_t)kTile * PL;
	static constexpr size_t kWhBytes = (size_t)kWarps * kBins * 4;
	static constexpr size_t kAdjBytes = (size_t)kBins * 8;
	static constexpr size_t kBytes = kRecBytes + kPlBytes + kWhBytes + kAdjBytes + 64;
};

enum { RANK_TICKET = 0, RANK_BALLOT = 1 };

template <int ES, int PL, bool FLOAT, typename OffT, int RANK, int THREADS, int ITEMS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) scatter_kernel(const ScatterParams p) {
	using R = typename Rec<ES>::type;
	using P = typename Payload<PL>::type;
	using SM = ScatterSmem<ES, PL, THREADS, ITEMS>;
	using SB = StatusBits<OffT>;
	constexpr int TILE = SM::kTile;
	constexpr int WARPS = SM::kWarps;
	constexpr uint32_t FULL = 0xFFFFFFFFu;

	extern __shared__ __align__(16) unsigned char smem[];
	R *s_rec = reinterpret_cast<R *>(smem);
	P *s_pl = reinterpret_cast<P *>(smem + SM::kRecBytes);
	uint32_t *s_wh = reinterpret_cast<uint32_t *>(smem + SM::kRecBytes + SM::kPlBytes);
	OffT *s_gadj = reinterpret_cast<OffT *>(smem + SM::kRecBytes + SM::kPlBytes + SM::kWhBytes);
	uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem + SM::kRecBytes + SM::kPlBytes + SM::kWhBytes + SM::kAdjBytes);
	// s_misc[0] = ticket broadcast, s_misc[1..8] = warp totals of the digit scan

	// ---- pass table (device-side column skipping, radix_sort.hpp:60-70) ----
	uint32_t ord = 0;
	bool last = true;
	if (p.ctl != nullptr) {
		const uint32_t early = p.ctl->early_exit, live = p.ctl->live_mask;
		if (early || !((live >> p.col) & 1u))
			return;
		ord = p.ctl->ordinal[p.col];
		last = ord + 1 == p.ctl->ncols;
	}
	const R *__restrict__ in = static_cast<const R *>(ord == 0 ? p.pb.rec_first : p.pb.rec_buf[(ord - 1) & 1]);
	R *__restrict__ out = static_cast<R *>(p.pb.rec_buf[ord & 1]);
	const P *__restrict__ pin = static_cast<const P *>(ord == 0 ? p.pb.pl_first : p.pb.pl_buf[(ord - 1) & 1]);
	P *__restrict__ pout = static_cast<P *>(p.pb.pl_buf[ord & 1]);
	const bool synth = PL != 0 && ord == 0 && p.pb.synth_index;
	const bool write_rec = !(last && p.pb.skip_last_rec);

	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t lt = lanemask_lt();
	const DigitDesc dd = p.dd;
	uint32_t *wh = s_wh + warp * kBins;
	OffT *status = static_cast<OffT *>(p.status);
	const R pad = make_pad<ES>(p.pad_rec);

	for (;;) {
		if (tid == 0)
			s_misc[0] = atomicAdd(p.ticket, 1u);
		__syncthreads(); // (E) also fences the previous tile's reads of s_rec / s_wh
		const uint32_t tile = s_misc[0];
		if (tile >= p.num_tiles)
			break;
		const size_t base = (size_t)tile * TILE;
		const size_t left = p.n - base;
		const uint32_t valid = left < (size_t)TILE ? (uint32_t)left : (uint32_t)TILE;
		const bool full = valid == (uint32_t)TILE;

		// ---- 1. load (warp-striped: item i of lane l is record warp*ITEMS*32 + i*32 + l) ----
		R rec[ITEMS];
		P pl[PL ? ITEMS : 1];
		const uint32_t t0 = warp * (ITEMS * 32) + lane;
		if (full) {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i)
				rec[i] = __ldg(in + base + t0 + i * 32);
		} else {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i)
				rec[i] = (t0 + i * 32 < valid) ? __ldg(in + base + t0 + i * 32) : pad;
		}
		if constexpr (PL != 0) {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t t = t0 + i * 32;
				if (synth)
					pl[i] = (P)(base + t);
				else
					pl[i] = (t < valid) ? __ldg(pin + base + t) : (P)0;
			}
		}
#pragma unroll
		for (int b = lane; b < kBins; b += 32)
			wh[b] = 0;
		__syncwarp();

		// ---- 2. rank inside the warp (stable: items ascending, lanes ascending) ----
		uint32_t rank[ITEMS];
		if constexpr (RANK == RANK_TICKET) {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i)
				rank[i] = atomicAdd(&wh[digit_of<ES, FLOAT>(rec[i], dd)], 1u);
		} else {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t d = digit_of<ES, FLOAT>(rec[i], dd);
				uint32_t peers = FULL;
#pragma unroll
				for (int b = 0; b < 8; ++b) {
					const bool bit = (d >> b) & 1u;
					const uint32_t v = __ballot_sync(FULL, bit);
					peers &= bit ? v : ~v;
				}
				const uint32_t leader = __ffs(peers) - 1;
				uint32_t old = 0;
				if (lane == leader)
					old = atomicAdd(&wh[d], (uint32_t)__popc(peers));
				old = __shfl_sync(FULL, old, leader);
				rank[i] = old + __popc(peers & lt);
			}
		}
		__syncthreads(); // (A)

		// ---- 3a. digit threads: warp prefixes, tile scan, publish aggregate ----
		uint32_t tcount = 0, tstart = 0;
		if (tid < kBins) {
			uint32_t c[WARPS];
#pragma unroll
			for (int w = 0; w < WARPS; ++w)
				c[w] = s_wh[w * kBins + tid];
#pragma unroll
			for (int w = 0; w < WARPS; ++w)
				tcount += c[w];
			// exclusive scan of tcount over the 256 digit threads (8 warps)
			uint32_t x = tcount;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t y = __shfl_up_sync(FULL, x, o);
				if (lane >= o)
					x += y;
			}
			if (lane == 31)
				s_misc[1 + warp] = x;
			asm volatile("bar.sync 1, 256;" ::: "memory");
			uint32_t wbase = 0;
#pragma unroll
			for (int w = 0; w < 8; ++w)
				wbase += (w < (int)warp) ? s_misc[1 + w] : 0u;
			tstart = wbase + x - tcount;
			uint32_t run = tstart;
#pragma unroll
			for (int w = 0; w < WARPS; ++w) {
				s_wh[w * kBins + tid] = run;
				run += c[w];
			}
			// tail padding sorts last (digit 255, after every real record): drop it from the count
			if (!full && tid == kBins - 1)
				tcount -= (uint32_t)TILE - valid;
			st_status(&status[(size_t)tile * kBins + tid],
			          (OffT)((tile == 0 ? SB::kPfx : SB::kAgg) | (OffT)tcount));
		}
		__syncthreads(); // (C)

		// ---- 4. records / payloads to their tile-sorted slot ----
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			const uint32_t d = digit_of<ES, FLOAT>(rec[i], dd);
			const uint32_t pos = wh[d] + rank[i];
			s_rec[pos] = rec[i];
			if constexpr (PL != 0)
				s_pl[pos] = pl[i];
		}

		// ---- 3b. decoupled look-back, one chain per digit ----
		if (tid < kBins) {
			OffT excl = 0;
			if (tile != 0) {
				size_t q = (size_t)(tile - 1) * kBins + tid;
				for (;;) {
					const OffT w = ld_status(&status[q]);
					if ((w & ~SB::kMask) == 0)
						continue; // predecessor has not published yet
					excl += w & SB::kMask;
					if (w & SB::kPfx)
						break;
					q -= kBins;
				}
				st_status(&status[(size_t)tile * kBins + tid], (OffT)(SB::kPfx | (excl + (OffT)tcount)));
			}
			s_gadj[tid] = (OffT)p.offs[tid] + excl - (OffT)tstart;
		}
		__syncthreads(); // (D)

		// ---- 5. coalesced per-bucket stores ----
#pragma unroll
		for (int i = 0; i < ITEMS; ++i) {
			const uint32_t s = tid + i * THREADS;
			if (full || s < valid) {
				const R r = s_rec[s];
				const uint32_t d = digit_of<ES, FLOAT>(r, dd);
				const OffT g = s_gadj[d] + (OffT)s;
				if (write_rec)
					out[g] = r;
				if constexpr (PL != 0)
					pout[g] = s_pl[s];
			}
		}
	}
}

// ---- geometry ---------------------------------------------------------------------------------
template <int ES, int PL> struct ScatterCfg {
	// records/thread shrink as the record + payload footprint grows (registers and smem)
	static constexpr int kThreads = 512;
	static constexpr int kItems = (ES + PL <= 4) ? 16 : (ES + PL <= 8) ? 12 : (ES + PL <= 16) ? 8 : 4;
	static constexpr int kMinBlocks = (ES + PL <= 8) ? 2 : 1;
};

template <int ES, int PL, bool FLOAT, typename OffT, int RANK>
cudaError_t launch_scatter_r(const ScatterParams &sp, int num_sms, cudaStream_t st) {
	using Cfg = ScatterCfg<ES, PL>;
	using SM = ScatterSmem<ES, PL, Cfg::kThreads, Cfg::kItems>;
	auto kern = scatter_kernel<ES, PL, FLOAT, OffT, RANK, Cfg::kThreads, Cfg::kItems, Cfg::kMinBlocks>;
	static int occ_cache[64] = {}; // per device
	int dev = 0;
	cudaGetDevice(&dev);
	int &ctas_per_sm = occ_cache[dev & 63];
	if (ctas_per_sm == 0) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::kBytes);
		if (e != cudaSuccess)
			return e;
		int occ = 0;
		e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::kThreads, SM::kBytes);
		if (e != cudaSuccess)
			return e;
		ctas_per_sm = occ > 0 ? occ : 1;
	}
	uint32_t grid = (uint32_t)num_sms * (uint32_t)ctas_per_sm;
	if (grid > sp.num_tiles)
		grid = sp.num_tiles;
	kern<<<grid, Cfg::kThreads, SM::kBytes, st>>>(sp);
	count_launch();
	return cudaGetLastError();
}

template <int ES, int PL, bool FLOAT, typename OffT>
cudaError_t launch_scatter_t(const ScatterParams &sp, int num_sms, cudaStream_t st) {
	return rank_mode() == RANK_TICKET ? launch_scatter_r<ES, PL, FLOAT, OffT, RANK_TICKET>(sp, num_sms, st)
	                                  : launch_scatter_r<ES, PL, FLOAT, OffT, RANK_BALLOT>(sp, num_sms, st);
}

template <int ES, int PL>
cudaError_t launch_scatter_pl(const ScatterParams &sp, bool is_float, bool wide, int num_sms, cudaStream_t st) {
	if constexpr (ES == 4 || ES == 8) {
		if (is_float)
			return wide ? launch_scatter_t<ES, PL, true, unsigned long long>(sp, num_sms, st)
			            : launch_scatter_t<ES, PL, true, uint32_t>(sp, num_sms, st);
	}
	if (is_float)
		return cudaErrorInvalidValue;
	return wide ? launch_scatter_t<ES, PL, false, unsigned long long>(sp, num_sms, st)
	            : launch_scatter_t<ES, PL, false, uint32_t>(sp, num_sms, st);
}

template <int ES>
cudaError_t launch_scatter_es(const ScatterParams &sp, int payload_bytes, bool is_float, bool wide,
                              int num_sms, cudaStream_t st) {
	switch (payload_bytes) {
	case 0: return launch_scatter_pl<ES, 0>(sp, is_float, wide, num_sms, st);
	case 4: return launch_scatter_pl<ES, 4>(sp, is_float, wide, num_sms, st);
	case 8: return launch_scatter_pl<ES, 8>(sp, is_float, wide, num_sms, st);
	}
	return cudaErrorInvalidValue;
}

// one translation unit per record size (parallel builds)
cudaError_t launch_scatter_1(const ScatterParams &, int, bool, bool, int, cudaStream_t);
cudaError_t launch_scatter_2(const ScatterParams &, int, bool, bool, int, cudaStream_t);
cudaError_t launch_scatter_4(const ScatterParams &, int, bool, bool, int, cudaStream_t);
cudaError_t launch_scatter_8(const ScatterParams &, int, bool, bool, int, cudaStream_t);
cudaError_t launch_scatter_16(const ScatterParams &, int, bool, bool, int, cudaStream_t);

} // namespace rsx
