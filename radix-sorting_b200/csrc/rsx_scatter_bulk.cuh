// rsx_scatter_bulk.cuh -- K3 with the write-out done by the TMA engine.
//
// Same pass as scatter_kernel (rsx_scatter.cuh: tile tickets, TMA-staged input, ticket/ballot
// ranking, decoupled look-back, tile-sorted shared buffer), different last step.  scatter_kernel
// stores the sorted tile with one 4..16-byte STG per record, which costs the shared-memory/LSU
// pipe three more wavefronts per 32 keys (sorted-slot load, bucket-base lookup, the store) on a
// kernel that is bound by exactly that pipe.  Here every digit bucket's run leaves as ONE
// cp.async.bulk.global.shared::cta copy issued by the bucket's digit thread: the copy engine reads
// shared memory and writes global memory asynchronously while the CTA is already ranking its next
// tile (tools/probe_bulk.cu: 256 runs of 160 bytes per tile sustain 6.5 TB/s as bulk copies vs
// 4.0 TB/s as 4-byte stores, with no SM instruction issue per record).
//
// Bulk copies need 16-byte aligned addresses on both sides and a multiple of 16 bytes.  A run's
// global address is only known after the look-back, so the order of the steps changes:
//   rank -> digit threads: counts, publish aggregate, LOOK-BACK, publish prefix
//        -> run d is placed in shared memory at a slot S[d] with S[d] == g[d] (mod A), where g[d] is
//           the run's global element index and A the number of elements per 16 bytes: slot sizes
//           roundup_A(g[d] mod A + count[d]) are scanned instead of the counts, which costs at most
//           2 (A-1) padding elements per bucket
//        -> records go to their slot, fence.proxy.async, barrier
//        -> digit thread d: <= A-1 head and <= A-1 tail elements with plain stores, the 16-byte
//           aligned body with one bulk copy; wait_group.read before the buffer is rewritten.
// Stability and the destination index are unchanged (dst = column offset + look-back prefix +
// position inside the tile's run), so the output is bit-identical to scatter_kernel's.
// Tiles that cannot use the copy engine (partial last tile, output buffers not 16-byte aligned)
// are placed densely and stored by the threads, as in scatter_kernel.
// (included from the middle of rsx_scatter.cuh, which provides everything used here)
#pragma once

namespace rsx {

template <int ES, int PL, class Cfg> struct BulkSmem {
	static constexpr int kTile = Cfg::kThreads * Cfg::kItems;
	static constexpr int kWarps = Cfg::kThreads / 32;
	static constexpr int kAlignRec = ES >= 16 ? 1 : 16 / ES;
	static constexpr int kAlignPl = PL == 0 ? 1 : (PL >= 16 ? 1 : 16 / PL);
	static constexpr int kAlign = kAlignRec > kAlignPl ? kAlignRec : kAlignPl; // elements per 16 bytes, coarsest lane
	static constexpr int kCap = (kTile + 2 * (kAlign - 1) * kBins + 15) / 16 * 16; // sorted-buffer slots
	static constexpr size_t kRecBytes = (size_t)kTile * ES;
	static constexpr size_t kPlBytes = (size_t)kTile * PL;
	static constexpr size_t kStageBytes = kRecBytes + kPlBytes;
	static constexpr size_t kSortedRecBytes = (size_t)kCap * ES;
	static constexpr size_t kSortedPlBytes = (size_t)kCap * PL;
	static constexpr size_t kWhBytes = (size_t)kWarps * kBins * 4;
	// layout: [stage rec | stage pl | sorted rec | sorted pl | warp counters | gadj | look-back partials | misc]
	static constexpr size_t kOffSorted = kStageBytes;
	static constexpr size_t kOffWh = kOffSorted + kSortedRecBytes + kSortedPlBytes;
	static constexpr size_t kOffAdj = kOffWh + kWhBytes;
	static constexpr size_t kOffLb = kOffAdj + (size_t)kBins * 8;
	static constexpr size_t kOffMisc = kOffLb + (size_t)kBins * 12;
	static constexpr size_t kBytes = kOffMisc + 96;
};

__device__ __forceinline__ void bulk_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
	             : "memory");
}

// One run of `count` elements: shared slots [S, S + count) -> global elements [g, g + count).
// S == g (mod 16 / sizeof(E)) by construction, both bases are 16-byte aligned.
template <typename E> __device__ __forceinline__ void emit_run(E *__restrict__ gout, const E *s, uint32_t S, unsigned long long g, uint32_t count) {
	constexpr uint32_t G = sizeof(E) >= 16 ? 1u : 16u / (uint32_t)sizeof(E);
	uint32_t head = (G - ((uint32_t)g & (G - 1u))) & (G - 1u);
	if (head > count)
		head = count;
	const uint32_t body = (count - head) / G * G;
	const uint32_t tail = count - head - body;
	for (uint32_t i = 0; i < head; ++i)
		gout[g + i] = s[S + i];
	if (body)
		bulk_s2g(gout + g + head, s + S + head, body * (uint32_t)sizeof(E));
	for (uint32_t i = 0; i < tail; ++i)
		gout[g + head + body + i] = s[S + head + body + i];
}

template <int ES, int PL, int DM, typename OffT, int RANK, class Cfg>
__global__ void __launch_bounds__(Cfg::kThreads, Cfg::kMinBlocks) scatter_bulk_kernel(const ScatterParams p) {
	using R = typename Rec<ES>::type;
	using P = typename Payload<PL>::type;
	using SM = BulkSmem<ES, PL, Cfg>;
	using SB = StatusBits<OffT>;
	constexpr int THREADS = Cfg::kThreads, ITEMS = Cfg::kItems;
	constexpr int TILE = SM::kTile;
	constexpr int WARPS = SM::kWarps;
	constexpr int LB = Cfg::kLookback;
	constexpr uint32_t FULL = 0xFFFFFFFFu;
	constexpr uint32_t A = (uint32_t)SM::kAlign;

	extern __shared__ __align__(128) unsigned char smem[];
	R *s_stage = reinterpret_cast<R *>(smem);
	P *s_stage_pl = reinterpret_cast<P *>(smem + SM::kRecBytes);
	R *s_rec = reinterpret_cast<R *>(smem + SM::kOffSorted);
	P *s_pl = reinterpret_cast<P *>(smem + SM::kOffSorted + SM::kSortedRecBytes);
	uint32_t *s_wh = reinterpret_cast<uint32_t *>(smem + SM::kOffWh);
	OffT *s_gadj = reinterpret_cast<OffT *>(smem + SM::kOffAdj);
	OffT *s_lbsum = reinterpret_cast<OffT *>(smem + SM::kOffLb);
	uint32_t *s_lbst = reinterpret_cast<uint32_t *>(smem + SM::kOffLb + (size_t)kBins * 8);
	uint32_t *s_misc = reinterpret_cast<uint32_t *>(smem + SM::kOffMisc);
	unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem + SM::kOffMisc + 80);
	// s_misc[0] = next tile ticket, [1..8] = warp totals of the digit scan, [9] = hot digit,
	// [10..17] = per-warp maxima of (count << 8 | digit)

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
	const bool stage_pl = PL != 0 && !synth;
	// the copy engine needs 16-byte aligned global addresses; tiles are multiples of 16 bytes
	const bool can_stage = (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
	                       (!stage_pl || (reinterpret_cast<uintptr_t>(pin) & 15) == 0);
	const bool can_bulk = (!write_rec || (reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
	                      (PL == 0 || (reinterpret_cast<uintptr_t>(pout) & 15) == 0);

	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const uint32_t lt = lanemask_lt();
	const DigitDesc dd = p.dd;
	uint32_t *wh = s_wh + warp * kBins;
	OffT *status = static_cast<OffT *>(p.status);
	const R pad = make_pad<ES>(p.pad_rec);
	const uint32_t full_tiles = (uint32_t)(p.n / TILE); // tiles [0, full_tiles) are complete
	const unsigned long long col_off = tid < kBins ? p.offs[tid] : 0ULL;

	auto prefetch = [&](uint32_t t) { // one thread
		const size_t base = (size_t)t * TILE;
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		mbar_expect_tx(s_bar, (uint32_t)(SM::kRecBytes + (stage_pl ? SM::kPlBytes : 0)));
		bulk_g2s(s_stage, in + base, (uint32_t)SM::kRecBytes, s_bar);
		if constexpr (PL != 0) {
			if (stage_pl)
				bulk_g2s(s_stage_pl, pin + base, (uint32_t)SM::kPlBytes, s_bar);
		}
	};

	// Tile tickets are claimed as late as possible (see scatter_kernel): every later tile's
	// look-back waits for this tile's aggregate.
	if (tid == 0) {
		mbar_init(s_bar, 1);
		const uint32_t t = atomicAdd(p.ticket, 1u);
		s_misc[0] = t;
		s_misc[9] = 0; // first tile: digit 0 as the hot-digit guess
		if (can_stage && t < full_tiles)
			prefetch(t);
	}
	__syncthreads();
	uint32_t tile = s_misc[0];
	uint32_t phase = 0;
#ifdef RSX_PHASE_TIMING
	unsigned long long dbg_acc[12] = {};
	long long dbg_last = clock64();
#endif

	while (tile < p.num_tiles) {
		RSX_T(9);
		const size_t base = (size_t)tile * TILE;
		const bool full = tile < full_tiles;
		const uint32_t valid = full ? (uint32_t)TILE : (uint32_t)(p.n - base);
		const bool staged = can_stage && full;
		const bool bulk = can_bulk && full;

		// ---- 1. tile -> registers (warp-striped ownership: item i of lane l is record
		//         warp*ITEMS*32 + i*32 + l), normally from the staging buffer the TMA prefetch filled.
		//         After this step nobody reads the staging buffer again, so the NEXT tile's prefetch
		//         can be issued before this tile is placed (step 3). ----
		const uint32_t t0 = warp * (ITEMS * 32) + lane;
		{
			uint4 *z = reinterpret_cast<uint4 *>(wh);
			z[lane] = make_uint4(0, 0, 0, 0);
			z[lane + 32] = make_uint4(0, 0, 0, 0);
		}
		R k[ITEMS];
		P pl[PL != 0 ? ITEMS : 1];
		if (staged) {
			mbar_wait(s_bar, phase);
			phase ^= 1u;
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				k[i] = s_stage[t0 + i * 32];
				if constexpr (PL != 0)
					pl[i] = synth ? (P)(base + t0 + i * 32) : s_stage_pl[t0 + i * 32];
			}
		} else {
			// unaligned input or the partial last tile: plain loads
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t t = t0 + i * 32;
				k[i] = (t < valid) ? __ldg(in + base + t) : pad;
				if constexpr (PL != 0)
					pl[i] = synth ? (P)(base + t) : ((t < valid) ? __ldg(pin + base + t) : (P)0);
			}
		}
		__syncwarp();
		RSX_T(0);

		// ---- 2. count: the ranking primitive of scatter_kernel, result discarded.  It runs a second
		//         time in step 4 on counters preset to the slot bases and then returns final slots:
		//         the same number of shared-memory operations as "rank + base lookup", but no rank
		//         registers, which is what lets the keys stay in registers instead. ----
		const uint32_t hot = s_misc[9];
		if constexpr (RANK == RANK_TICKET) {
			uint32_t hotcnt = 0;
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t d = tile_digit<ES, DM>(p, k[i], dd);
				const bool is_hot = d == hot;
				const uint32_t m = __ballot_sync(FULL, is_hot);
				if (!is_hot)
					atomicAdd(&wh[d], 1u);
				hotcnt += __popc(m);
			}
			if (lane == 0)
				wh[hot] = hotcnt; // no atomic touched this counter
		} else {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t d = tile_digit<ES, DM>(p, k[i], dd);
				uint32_t peers = FULL;
#pragma unroll
				for (int b = 0; b < 8; ++b) {
					const bool bit = (d >> b) & 1u;
					const uint32_t v = __ballot_sync(FULL, bit);
					peers &= bit ? v : ~v;
				}
				if (lane == (uint32_t)__ffs(peers) - 1)
					atomicAdd(&wh[d], (uint32_t)__popc(peers));
			}
		}
		RSX_T(1);
		__syncthreads(); // (A) all warp counters final; the staging buffer is free
		RSX_T(2);

		// ---- 3. digit threads: tile counts, publish aggregate, look-back, slots ----
		constexpr bool kPair = THREADS >= 2 * kBins;
		const uint32_t dgt = tid & (kBins - 1), half = tid / kBins;
		uint32_t run_slot = 0, run_count = 0;   // digit thread: this tile's run of its bucket
		unsigned long long run_g = 0;
		{
			uint32_t c[WARPS];
			uint32_t tcount = 0, agg = 0;
			if (half == 0) {
#pragma unroll
				for (int w = 0; w < WARPS; ++w)
					c[w] = s_wh[w * kBins + tid];
#pragma unroll
				for (int w = 0; w < WARPS; ++w)
					tcount += c[w];
				// tail padding sorts last: not part of the aggregate
				agg = (!full && tid == (uint32_t)kBins - 1) ? tcount - ((uint32_t)TILE - valid) : tcount;
				st_status(&status[(size_t)tile * kBins + tid], (OffT)((tile == 0 ? SB::kPfx : SB::kAgg) | (OffT)agg));
			}
			// decoupled look-back, one chain per digit; the first round is split over two threads per
			// digit so that 2*LB predecessors cost one L2 round trip
			OffT part = 0;
			uint32_t st = 0, used = 0; // st: 0 = only aggregates so far, 1 = reached a prefix, 2 = hit an unpublished word
			if (half < (kPair ? 2u : 1u)) {
				const int q = (int)tile - 1 - (int)half * LB;
				OffT w[LB];
#pragma unroll
				for (int j = 0; j < LB; ++j)
					w[j] = (q - j >= 0) ? ld_status(&status[(size_t)(q - j) * kBins + dgt]) : (OffT)SB::kPfx;
#pragma unroll
				for (int j = 0; j < LB; ++j) {
					if (st == 0) {
						if ((w[j] & ~SB::kMask) == 0) {
							st = 2;
						} else {
							part += w[j] & SB::kMask;
							++used;
							if (w[j] & SB::kPfx)
								st = 1;
						}
					}
				}
			}
			if constexpr (kPair) {
				if (half == 1) {
					s_lbsum[dgt] = part;
					s_lbst[dgt] = st;
					asm volatile("bar.arrive %0, 64;" ::"r"(2 + (warp & 7)) : "memory");
				}
			}
			if (half == 0) {
				OffT excl = part;
				bool done = st == 1;
				int q = (int)tile - 1 - (int)used;
				if constexpr (kPair) {
					asm volatile("bar.sync %0, 64;" ::"r"(2 + warp) : "memory");
					if (st == 0) { // own window was all aggregates: splice the partner's window
						const uint32_t pst = s_lbst[dgt];
						if (pst != 2) {
							excl += s_lbsum[dgt];
							q -= LB;
							done = pst == 1;
						}
					}
				}
				while (!done) { // rare: long chains and unpublished predecessors
#ifdef RSX_PHASE_TIMING
					if (tid == 0) dbg_acc[10] += 1;
#endif
					OffT w[LB];
#pragma unroll
					for (int j = 0; j < LB; ++j)
						w[j] = (q - j >= 0) ? ld_status(&status[(size_t)(q - j) * kBins + dgt]) : (OffT)SB::kPfx;
#pragma unroll
					for (int j = 0; j < LB; ++j) {
						if (!done) {
							OffT v = w[j];
							while ((v & ~SB::kMask) == 0) { // not published yet: poll this one word, politely
#ifdef RSX_PHASE_TIMING
								if (tid == 0) dbg_acc[11] += 1;
#endif
								__nanosleep(40);
								v = ld_status(&status[(size_t)(q - j) * kBins + dgt]);
							}
							excl += v & SB::kMask;
							done = (v & SB::kPfx) != 0;
						}
					}
					q -= LB;
				}
				if (tile != 0)
					st_status(&status[(size_t)tile * kBins + dgt], (OffT)(SB::kPfx | (excl + (OffT)agg)));
				RSX_T(3);

				// slot of this bucket's run: congruent to its global index modulo A when the copy
				// engine stores the tile, dense otherwise
				run_g = col_off + (unsigned long long)excl;
				const uint32_t a = bulk ? ((uint32_t)run_g & (A - 1u)) : 0u;
				const uint32_t slot = bulk ? (tcount ? (a + tcount + A - 1u) / A * A : 0u) : tcount;
				uint32_t x = slot;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t y = __shfl_up_sync(FULL, x, o);
					if (lane >= o)
						x += y;
				}
				const uint32_t wmax = __reduce_max_sync(FULL, (tcount << 8) | tid); // hot digit for the next tile
				if (lane == 31) {
					s_misc[1 + warp] = x;
					s_misc[10 + warp] = wmax;
				}
				asm volatile("bar.sync 1, 256;" ::: "memory");
				uint32_t wbase = 0, hmax = 0;
#pragma unroll
				for (int w = 0; w < 8; ++w) {
					wbase += (w < (int)warp) ? s_misc[1 + w] : 0u;
					hmax = max(hmax, s_misc[10 + w]);
				}
				if (tid == 0)
					s_misc[9] = hmax & 0xFFu;
				run_slot = wbase + x - slot + a;
				run_count = agg;
				uint32_t run = run_slot;
#pragma unroll
				for (int w = 0; w < WARPS; ++w) {
					s_wh[w * kBins + tid] = run;
					run += c[w];
				}
				if (!bulk)
					s_gadj[tid] = (OffT)(run_g - run_slot);
				RSX_T(6);
				// the previous tile's bulk stores must have finished reading the sorted buffer
				asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
				RSX_T(9);
				if (tid == 0) {
					// next ticket, claimed as late as the prefetch allows: the load then overlaps this
					// tile's placement and store issue
					const uint32_t nt = atomicAdd(p.ticket, 1u);
					s_misc[0] = nt;
					if (can_stage && nt < full_tiles)
						prefetch(nt);
				}
			}
		}
		__syncthreads(); // (C)
		const uint32_t next_tile = s_misc[0];
		RSX_T(4);

		// ---- 4. records / payloads to their slot: the same ticket sequence as step 2, now on
		//         counters that start at the warp's slot base for each digit ----
		if constexpr (RANK == RANK_TICKET) {
			const uint32_t hotbase = wh[hot];
			uint32_t hotrun = 0;
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t d = tile_digit<ES, DM>(p, k[i], dd);
				const bool is_hot = d == hot;
				const uint32_t m = __ballot_sync(FULL, is_hot);
				uint32_t pos;
				if (is_hot)
					pos = hotbase + hotrun + __popc(m & lt);
				else
					pos = atomicAdd(&wh[d], 1u);
				hotrun += __popc(m);
				s_rec[pos] = k[i];
				if constexpr (PL != 0)
					s_pl[pos] = pl[i];
			}
		} else {
#pragma unroll
			for (int i = 0; i < ITEMS; ++i) {
				const uint32_t d = tile_digit<ES, DM>(p, k[i], dd);
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
				const uint32_t pos = old + __popc(peers & lt);
				s_rec[pos] = k[i];
				if constexpr (PL != 0)
					s_pl[pos] = pl[i];
			}
		}
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // slots become visible to the copy engine
		RSX_T(5);
		__syncthreads(); // (D) sorted tile complete
		RSX_T(7);

		// ---- 5. write-out ----
		if (bulk) {
			if (half == 0) {
				if (write_rec)
					emit_run<R>(out, s_rec, run_slot, run_g, run_count);
				if constexpr (PL != 0)
					emit_run<P>(pout, s_pl, run_slot, run_g, run_count);
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
		} else {
			for (uint32_t s = tid; s < valid; s += THREADS) {
				const R r = s_rec[s];
				const OffT g = s_gadj[tile_digit<ES, DM>(p, r, dd)] + (OffT)s;
				if (write_rec)
					out[g] = r;
				if constexpr (PL != 0)
					pout[g] = s_pl[s];
			}
		}
		RSX_T(8);
		tile = next_tile;
	}
#ifdef RSX_PHASE_TIMING
	if (tid == 0 && p.dbg) {
		for (int k = 0; k < 12; ++k)
			atomicAdd(&p.dbg[k], dbg_acc[k]);
		atomicAdd(&p.dbg[12], 1ULL);
	}
#endif
	// shared memory must outlive the copy engine's reads
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <int ES, int PL, int DM, typename OffT, int RANK, class Cfg>
cudaError_t launch_scatter_bulk(const ScatterParams &sp, int num_sms, cudaStream_t st) {
	using SM = BulkSmem<ES, PL, Cfg>;
	auto kern = scatter_bulk_kernel<ES, PL, DM, OffT, RANK, Cfg>;
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
	ScatterParams q = sp;
	q.num_tiles = (uint32_t)((sp.n + SM::kTile - 1) / SM::kTile);
	uint32_t grid = (uint32_t)num_sms * (uint32_t)ctas_per_sm;
	if (grid > q.num_tiles)
		grid = q.num_tiles;
	kern<<<grid, Cfg::kThreads, SM::kBytes, st>>>(q);
	count_launch();
	return cudaGetLastError();
}

} // namespace rsx
