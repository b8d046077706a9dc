// rsx_multi.cu -- single-box multi-GPU partitioned sort: the host-side orchestration, in C++.
//
// The reference (eloj/radix-sorting) is single-threaded host code with no multi-device path; this
// is the MSD-then-LSD composition of its own primitives that BASELINE config 5 asks for
// (SURVEY.md §8e).  One "shard" per GPU; the concatenation of the shards in rank order is the
// logical input, the concatenation of the outputs is radix_sort() of it, bit for bit:
//   1. K1 on the shard -> (columns x 256) digit counts              rsx_histogram
//   2. all-gather of the counts (KiBs): every rank knows the global histogram, the globally live
//      columns and every (source, destination) transfer size
//   3. routing: the highest globally-live column's 256 buckets go to ranks as contiguous balanced
//      ranges; when one bucket is too heavy for that (zipf keys) the sort routes by KEY RANGE
//      instead: splitters at the quantiles of a pooled sample, exact counts from one counting pass
//   4. fused partition + exchange: the stable K3 pass on the routing digit stores every record
//      straight into its owner's receive buffer (peer memory over NVLink), chunks in source-rank
//      order (global stability)                                  rsx_scatter_pass_to / rsx_split_pass_to
//      -- or, without peer mappings, a local stable partition followed by the caller's all-to-all
//   5. local LSD sort of what was received                          rsx_sort
// Everything here is host logic on top of the C ABI's own entry points; the collectives are two
// callbacks (all-gather of small host buffers, barrier), so the same code serves one process with a
// thread per GPU (rsx_sort_multi below) and one process per GPU under torchrun (dist.py passes
// torch.distributed callbacks).  The local primitives are a table too (rsx_shard_ops): NULL selects
// the CUDA kernels; the CPU test-suite plugs oracle-backed ones in to run this orchestration over
// gloo without a GPU.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "rsx.h"

namespace {

using clk = std::chrono::steady_clock;
double since(clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); }

unsigned long long width_mask(uint32_t kb) { return kb >= 8 ? ~0ULL : ((1ULL << (8u * kb)) - 1ULL); }

// radix_sort_basic_kdf.hpp:19-46 on the host, for sampled records
unsigned long long derive_host(const unsigned char *rec, const rsx_layout &L) {
	unsigned long long k = 0;
	memcpy(&k, rec + L.key_offset, L.key_bytes);
	const unsigned long long m = width_mask(L.key_bytes), top = 1ULL << (8u * L.key_bytes - 1u);
	if (L.kdf_kind == RSX_KDF_SIGNED)
		k ^= top;
	else if (L.kdf_kind == RSX_KDF_FLOAT)
		k ^= (k & top) ? m : top;
	if (L.flags & RSX_FLAG_INVERT)
		k = ~k & m;
	return k;
}

// ---- default (CUDA) local primitives ----------------------------------------------------------
int cuda_hist(void *, const void *src, size_t n, const rsx_layout *L, uint64_t *hist, void *stream) {
	memset(hist, 0, sizeof(uint64_t) * 256 * L->key_bytes);
	if (n == 0)
		return RSX_OK;
	if (n == 1) { // K1 needs n >= 2 (like the reference, which returns before counting)
		unsigned char rec[16];
		if (cudaMemcpy(rec, src, L->record_bytes, cudaMemcpyDeviceToHost) != cudaSuccess)
			return RSX_ERR_CUDA;
		const unsigned long long k = derive_host(rec, *L);
		for (uint32_t c = 0; c < L->key_bytes; ++c)
			hist[c * 256 + ((k >> (8 * c)) & 0xFF)] = 1;
		return RSX_OK;
	}
	return rsx_histogram(src, n, L, hist, nullptr, nullptr, stream);
}

int cuda_sample(void *, const void *src, size_t n, const rsx_layout *L, size_t count, uint64_t *derived, void *stream) {
	if (count == 0)
		return RSX_OK;
	return rsx_sample_keys(src, n, L, count, derived, stream); // record i * (n / count), i < count
}

int cuda_split_counts(void *, const void *src, size_t n, const rsx_layout *L, const uint64_t *split, int nsplit,
                      uint64_t *counts, void *stream) {
	return rsx_split_counts(src, n, L, split, nsplit, counts, stream);
}

int cuda_partition_to(void *, const void *src, size_t n, const rsx_layout *L, int col, const uint8_t *owner,
                      const uint64_t *split, int nsplit, const uint64_t *dest_base, int ndest, void *stream) {
	if (n == 0)
		return RSX_OK;
	if (col >= 0)
		return rsx_scatter_pass_to(src, n, L, col, owner, dest_base, ndest, stream);
	return rsx_split_pass_to(src, n, L, split, nsplit, dest_base, stream);
}

int cuda_sort(void *, void *src, void *aux, size_t n, const rsx_layout *L, void **result, void *stream) {
	return rsx_sort(src, aux, n, L, result, nullptr, stream);
}

int cuda_hist_column(void *, const void *src, size_t n, const rsx_layout *L, int col, uint64_t *hist, void *stream) {
	memset(hist, 0, sizeof(uint64_t) * 256);
	if (n == 0)
		return RSX_OK;
	return rsx_histogram_column(src, n, L, col, hist, stream);
}

const rsx_shard_ops kCudaOps = {cuda_hist, cuda_sample, cuda_split_counts, cuda_partition_to, cuda_sort, nullptr, cuda_hist_column};

} // namespace

extern "C" {

// ---- routing: pure host arithmetic, identical on every rank -----------------------------------------
int rsx_multi_route(const uint64_t *hist_all, int world, int cols, int rank, double skew_threshold, rsx_route *out) {
	if (!hist_all || !out || world < 1 || world > RSX_MAX_RANKS || cols < 1 || cols > 8 || rank < 0 || rank >= world)
		return RSX_ERR_INVALID;
	memset(out, 0, sizeof(*out));
	std::vector<uint64_t> total((size_t)cols * 256, 0);
	for (int g = 0; g < world; ++g)
		for (int i = 0; i < cols * 256; ++i)
			total[i] += hist_all[(size_t)g * cols * 256 + i];
	uint64_t n_total = 0;
	for (int b = 0; b < 256; ++b)
		n_total += total[b];
	out->n_total = n_total;
	out->routing_column = -1;
	// routing digit = highest column that is not constant over ALL ranks (column skipping of
	// radix_sort.hpp:65-70 lifted to the global level)
	for (int c = 0; c < cols; ++c) {
		uint64_t mx = 0;
		for (int b = 0; b < 256; ++b)
			mx = std::max(mx, total[(size_t)c * 256 + b]);
		if (mx != n_total) {
			out->live_mask |= 1u << c;
			out->routing_column = c;
		}
	}
	for (int g = 0; g < world; ++g) { // per-rank input sizes (column 0 sums to the shard size)
		uint64_t s = 0;
		for (int b = 0; b < 256; ++b)
			s += hist_all[(size_t)g * cols * 256 + b];
		out->n_in[g] = s;
	}
	if (out->routing_column < 0 || world == 1) { // nothing to route: every rank sorts its own shard
		out->routing_column = -1;
		out->n_out = out->n_in[rank];
		out->max_n_out = *std::max_element(out->n_in, out->n_in + world);
		out->imbalance = 1.0;
		return RSX_OK;
	}
	const int top = out->routing_column;
	const uint64_t *gc = &total[(size_t)top * 256];
	// contiguous bucket ranges per rank: a bucket goes to the rank whose ideal share contains its midpoint
	double cum = 0;
	int prev = 0;
	for (int b = 0; b < 256; ++b) {
		const double mid = cum + (double)gc[b] / 2.0;
		int o = n_total ? (int)(mid * world / (double)n_total) : 0;
		o = std::min(o, world - 1);
		o = std::max(o, prev); // monotone
		out->owner[b] = (uint8_t)o;
		prev = o;
		cum += (double)gc[b];
	}
	uint64_t recv_tot[RSX_MAX_RANKS] = {};
	for (int g = 0; g < world; ++g) {
		const uint64_t *h = &hist_all[((size_t)g * cols + top) * 256];
		for (int b = 0; b < 256; ++b) {
			const int d = out->owner[b];
			if (g == rank)
				out->send_counts[d] += h[b];
			if (d == rank)
				out->recv_counts[g] += h[b];
			if (g < rank)
				out->dest_offset[d] += h[b]; // records of lower ranks precede ours in destination d
			recv_tot[d] += h[b];
		}
	}
	out->n_out = recv_tot[rank];
	out->max_n_out = *std::max_element(recv_tot, recv_tot + world);
	out->imbalance = (double)out->max_n_out / std::max(1.0, (double)n_total / world);
	if (out->imbalance > skew_threshold)
		out->key_range = 1; // bucket-granular ranges cannot balance: route by key range
	return RSX_OK;
}

int rsx_multi_splitters(const uint64_t *samples, size_t count, int world, uint64_t *splitters) {
	if (!samples || !splitters || world < 2 || world > RSX_MAX_RANKS)
		return RSX_ERR_INVALID;
	std::vector<uint64_t> s(samples, samples + count);
	std::sort(s.begin(), s.end());
	for (int i = 1; i < world; ++i)
		splitters[i - 1] = count ? s[std::min(count - 1, (size_t)i * count / world)] : ~0ULL;
	return RSX_OK;
}

// ---- one rank's view of the partitioned sort ---------------------------------------------------------
int rsx_sort_shard(const rsx_comm *comm, const rsx_shard_ops *ops_in, void *src, size_t n, void *recv,
                   void *const *recv_peers, size_t capacity, const rsx_layout *L, uint32_t flags, void **result,
                   size_t *n_out, rsx_multi_report *rep, void *stream) {
	if (!comm || !L || !result || !n_out || comm->world < 1 || comm->world > RSX_MAX_RANKS || !comm->allgather || !comm->barrier)
		return RSX_ERR_INVALID;
	if (!(L->key_bytes == 1 || L->key_bytes == 2 || L->key_bytes == 4 || L->key_bytes == 8) || L->record_bytes < L->key_bytes)
		return RSX_ERR_INVALID;
	const rsx_shard_ops &ops = ops_in ? *ops_in : kCudaOps;
	void *octx = ops.ctx;
	const int world = comm->world, rank = comm->rank, cols = (int)L->key_bytes;
	const size_t rb = L->record_bytes;
	rsx_multi_report local;
	if (!rep)
		rep = &local;
	memset(rep, 0, sizeof(*rep));
	rep->routing_column = -1;
	int r;
	auto t0 = clk::now();

	// Keys-only records over peer-mapped buffers: append-mode exchange.  Equal records are
	// indistinguishable, so neither the order in which the sources' runs land in a receive buffer
	// nor exact per-source offsets matter: bucket ranges (or key ranges) are balanced on a SAMPLE of
	// the routing digit, every (tile, destination) run reserves its place with one atomic on the
	// destination's append cursor, and the full routing histogram pass (and, for skewed keys, the
	// counting pass) disappears.  Anything else takes the exact path below.
	if (!ops_in && recv_peers && !(flags & (RSX_MULTI_NO_FUSED | RSX_MULTI_EXACT)) && L->record_bytes == L->key_bytes &&
	    cols > 1 && world > 1 && capacity * rb >= 4096) {
		const size_t stride = std::max<size_t>(1, n / 131072);
		std::vector<uint64_t> h(256 + 2), g((256 + 2) * (size_t)world);
		if (n && (r = rsx_histogram_column_sampled(src, n, L, cols - 1, stride, h.data(), stream)))
			return r;
		h[256] = n;
		h[257] = capacity;
		if ((r = comm->allgather(comm->ctx, h.data(), g.data(), h.size() * sizeof(uint64_t))))
			return r;
		const size_t hw = (size_t)cols * 256;
		std::vector<uint64_t> hall(hw * world, 0);
		uint64_t n_total = 0, n_max = 0, cap_min = ~0ULL, tot = 0, mx = 0, colsum[256] = {};
		for (int q = 0; q < world; ++q) {
			const uint64_t *hq = &g[(size_t)q * 258];
			uint64_t sq = 0;
			for (int b = 0; b < 256; ++b) {
				hall[(size_t)q * hw + (size_t)(cols - 1) * 256 + b] = hq[b];
				colsum[b] += hq[b];
				sq += hq[b];
			}
			for (int c = 0; c + 1 < cols; ++c) // unknown columns: "not constant"
				hall[(size_t)q * hw + (size_t)c * 256] = sq - sq / 2, hall[(size_t)q * hw + (size_t)c * 256 + 1] = sq / 2;
			n_total += hq[256];
			n_max = std::max(n_max, hq[256]);
			cap_min = std::min(cap_min, hq[257]);
		}
		for (int b = 0; b < 256; ++b) {
			tot += colsum[b];
			mx = std::max(mx, colsum[b]);
		}
		// the cursor of every rank sits at the SAME offset of its receive buffer (the last aligned
		// 16 bytes of the smallest one), so that every rank can address every cursor
		const uint64_t coff = ((cap_min * rb) & ~(uint64_t)15) - 16;
		const uint64_t cap_eff = coff / rb;
		cap_min = cap_eff;
		// Routing by the TOP digit is valid for any input; when the sample shows it constant or too
		// skewed to balance, key ranges (splitters over whole derived keys, equally valid for any
		// input) take over.  Only an input whose sampled keys are all equal goes to the exact path,
		// which knows how to leave constant shards where they are.
		bool go = tot != 0;
		rsx_route route;
		memset(&route, 0, sizeof(route));
		if (go && (r = rsx_multi_route(hall.data(), world, cols, rank, (flags & RSX_MULTI_NO_KEY_RANGE) ? 1e30 : 1.15, &route)))
			return r;
		const bool top_constant = mx == tot;
		uint64_t splitters[RSX_MAX_RANKS] = {};
		int nsplit = 0;
		if (go && (top_constant || route.key_range)) {
			if (world - 1 > 15 || (flags & RSX_MULTI_NO_KEY_RANGE)) {
				go = false;
			} else {
				constexpr size_t kSamples = 2048;
				std::vector<uint64_t> mine(1 + kSamples, ~0ULL), all((1 + kSamples) * world);
				const size_t cnt = std::min(kSamples, n);
				mine[0] = cnt;
				if (cnt && (r = ops.sample(octx, src, n, L, cnt, mine.data() + 1, stream)))
					return r;
				if ((r = comm->allgather(comm->ctx, mine.data(), all.data(), mine.size() * sizeof(uint64_t))))
					return r;
				std::vector<uint64_t> pooled;
				for (int q = 0; q < world; ++q) {
					const uint64_t *pq = &all[(size_t)q * (1 + kSamples)];
					pooled.insert(pooled.end(), pq + 1, pq + 1 + pq[0]);
				}
				if (pooled.empty() || *std::min_element(pooled.begin(), pooled.end()) == *std::max_element(pooled.begin(), pooled.end())) {
					go = false; // (nearly) constant keys: the exact path decides
				} else {
					nsplit = world - 1;
					rsx_multi_splitters(pooled.data(), pooled.size(), world, splitters);
					rep->key_range = 1;
				}
			}
		}
		// Every (tile, destination) run costs one atomic on the destination's cursor: world x tiles
		// of them per cursor.  Measured on 8 GPUs with 4-byte keys (780 K atomics per cursor in a
		// 5 ms pass) the cursor becomes the bottleneck (exchange 7.1 ms vs 5.3 ms exact), with
		// 8-byte keys (half the tiles per byte) and at 2 GPUs it does not: bucket-range routing
		// appends only while world <= record_bytes; key-range routing always does (it also saves
		// the counting pass, which outweighs the contention: config 5 zipf 87 -> 66 ms on 8 GPUs).
		if (go && !rep->key_range && world > (int)rb && !(flags & RSX_MULTI_APPEND))
			go = false;
		if (go) {
			rep->live_mask = 1u << (cols - 1);
			rep->n_total = n_total;
			if (!rep->key_range)
				rep->routing_column = route.routing_column;
			rep->seconds_histogram = since(t0);
			t0 = clk::now();
			// room for the estimated shard + sampling error; every rank reaches the same verdict
			const double est = rep->key_range ? (double)n_total / world : (double)route.max_n_out / (double)std::max<uint64_t>(tot, 1) * (double)n_total;
			rep->needed_capacity = (uint64_t)(std::max(est * 1.03, (double)n_max)) + 4096;
			if (rep->needed_capacity > cap_min)
				return RSX_ERR_WORKSPACE;
			rep->seconds_routing = since(t0);
			t0 = clk::now();
			uint64_t base[RSX_MAX_RANKS], cursor[RSX_MAX_RANKS], caps[RSX_MAX_RANKS];
			for (int d = 0; d < world; ++d) {
				base[d] = (uint64_t)(uintptr_t)recv_peers[d];
				cursor[d] = base[d] + coff;
				caps[d] = cap_eff;
			}
			if (cudaMemsetAsync(static_cast<unsigned char *>(recv) + coff, 0, 16, static_cast<cudaStream_t>(stream)) != cudaSuccess)
				return RSX_ERR_CUDA;
			if ((r = comm->barrier(comm->ctx))) // every cursor is zero, nobody still sorts out of its receive buffer
				return r;
			uint32_t overflow = 0;
			if (n && (r = rsx_scatter_pass_append(src, n, L, rep->key_range ? -1 : route.routing_column, route.owner, splitters,
			                                      nsplit, base, cursor, caps, world, &overflow, stream)))
				return r;
			if ((r = comm->barrier(comm->ctx))) // all runs have landed
				return r;
			uint64_t got = 0;
			if (cudaMemcpy(&got, static_cast<unsigned char *>(recv) + coff, sizeof(got), cudaMemcpyDeviceToHost) != cudaSuccess)
				return RSX_ERR_CUDA;
			uint64_t mine2[2] = {overflow, got}, all2[2 * RSX_MAX_RANKS];
			if ((r = comm->allgather(comm->ctx, mine2, all2, sizeof(mine2))))
				return r;
			uint64_t any_overflow = 0, got_max = 0, got_sum = 0;
			for (int q = 0; q < world; ++q) {
				any_overflow |= all2[2 * q];
				got_max = std::max(got_max, all2[2 * q + 1]);
				got_sum += all2[2 * q + 1];
			}
			if (any_overflow || got_sum != n_total) { // a run did not fit (src is intact): ask for more room
				rep->needed_capacity = (uint64_t)((double)std::max(got_max, rep->needed_capacity) * 1.25) + 4096;
				return RSX_ERR_WORKSPACE;
			}
			rep->fused = 1;
			rep->append = 1;
			rep->imbalance = (double)got_max / std::max(1.0, (double)n_total / world);
			rep->seconds_exchange = since(t0);
			t0 = clk::now();
			*n_out = (size_t)got;
			*result = recv;
			if (got > 1 && (r = ops.sort(octx, recv, src, (size_t)got, L, result, stream)))
				return r;
			rep->seconds_local_sort = since(t0);
			return RSX_OK;
		}
		t0 = clk::now();
	}

	// 1-2. histograms of every rank, visible to every rank.  Routing only needs the TOP column, and
	// counting one column runs at HBM speed (one atomic per record instead of key_bytes), so that
	// is tried first; only when the top column is constant over all ranks (keys with constant high
	// bytes) every column is counted to find the highest live one.
	// (the all-gather also carries every rank's buffer capacity: the too-small verdict below must
	// be the same on every rank even when the ranks' buffers differ)
	const size_t hwords = (size_t)cols * 256;
	std::vector<uint64_t> hist_all(hwords * world, 0);
	uint64_t min_capacity = ~0ULL;
	bool have_route_hist = false;
	if (ops.hist_column && cols > 1 && !(flags & RSX_MULTI_FULL_HISTOGRAM)) {
		std::vector<uint64_t> h(256 + 1), g((256 + 1) * (size_t)world);
		if ((r = ops.hist_column(octx, src, n, L, cols - 1, h.data(), stream)))
			return r;
		h[256] = capacity;
		if ((r = comm->allgather(comm->ctx, h.data(), g.data(), h.size() * sizeof(uint64_t))))
			return r;
		uint64_t tot[256] = {}, n_total = 0, mx = 0;
		for (int gq = 0; gq < world; ++gq) {
			for (int b = 0; b < 256; ++b)
				tot[b] += g[(size_t)gq * 257 + b];
			min_capacity = std::min(min_capacity, g[(size_t)gq * 257 + 256]);
		}
		for (int b = 0; b < 256; ++b) {
			n_total += tot[b];
			mx = std::max(mx, tot[b]);
		}
		if (mx != n_total) { // the top column is live: it is the routing digit
			have_route_hist = true;
			for (int gq = 0; gq < world; ++gq) {
				uint64_t *dst = &hist_all[(size_t)gq * hwords];
				const uint64_t *top = &g[(size_t)gq * 257];
				uint64_t s = 0;
				for (int b = 0; b < 256; ++b) {
					dst[(size_t)(cols - 1) * 256 + b] = top[b];
					s += top[b];
				}
				for (int c = 0; c + 1 < cols; ++c) // unknown columns: any two buckets, "not constant"
					dst[(size_t)c * 256] = s - s / 2, dst[(size_t)c * 256 + 1] = s / 2;
			}
		}
	}
	if (!have_route_hist) {
		std::vector<uint64_t> hist(hwords + 1), gathered((hwords + 1) * world);
		if ((r = ops.hist(octx, src, n, L, hist.data(), stream)))
			return r;
		hist[hwords] = capacity;
		if ((r = comm->allgather(comm->ctx, hist.data(), gathered.data(), (hwords + 1) * sizeof(uint64_t))))
			return r;
		min_capacity = ~0ULL;
		for (int g = 0; g < world; ++g) {
			memcpy(&hist_all[(size_t)g * hwords], &gathered[(size_t)g * (hwords + 1)], hwords * sizeof(uint64_t));
			min_capacity = std::min(min_capacity, gathered[(size_t)g * (hwords + 1) + hwords]);
		}
	}
	rsx_route route;
	const double thr = (flags & RSX_MULTI_NO_KEY_RANGE) ? 1e30 : 1.15;
	if ((r = rsx_multi_route(hist_all.data(), world, cols, rank, thr, &route)))
		return r;
	rep->live_mask = have_route_hist ? (1u << (cols - 1)) : route.live_mask; // top-column shortcut: only that column is known
	rep->n_total = route.n_total;
	rep->seconds_histogram = since(t0);
	t0 = clk::now();

	if (route.routing_column < 0) { // constant keys everywhere, or a single rank: local sort only
		*n_out = n;
		rep->imbalance = 1.0;
		*result = src;
		if (n > 1 && (r = ops.sort(octx, src, recv, n, L, result, stream)))
			return r;
		rep->seconds_local_sort = since(t0);
		return RSX_OK;
	}

	// 3. routing table: bucket ranges, or key-range splitters for skewed keys
	uint64_t splitters[RSX_MAX_RANKS] = {};
	int nsplit = 0;
	uint64_t send_counts[RSX_MAX_RANKS], recv_counts[RSX_MAX_RANKS], dest_offset[RSX_MAX_RANKS];
	uint64_t my_n_out = route.n_out, max_n_out = route.max_n_out;
	if (route.key_range && world - 1 <= 15) {
		constexpr size_t kSamples = 2048;
		std::vector<uint64_t> mine(1 + kSamples, ~0ULL), all((1 + kSamples) * world);
		const size_t cnt = std::min(kSamples, n);
		mine[0] = cnt;
		if (cnt && (r = ops.sample(octx, src, n, L, cnt, mine.data() + 1, stream)))
			return r;
		if ((r = comm->allgather(comm->ctx, mine.data(), all.data(), mine.size() * sizeof(uint64_t))))
			return r;
		std::vector<uint64_t> pooled;
		for (int g = 0; g < world; ++g) {
			const uint64_t *p = &all[(size_t)g * (1 + kSamples)];
			pooled.insert(pooled.end(), p + 1, p + 1 + p[0]);
		}
		nsplit = world - 1;
		rsx_multi_splitters(pooled.data(), pooled.size(), world, splitters);
		uint64_t cnts[RSX_MAX_RANKS] = {}, cnts_all[RSX_MAX_RANKS * RSX_MAX_RANKS];
		if (n && (r = ops.split_counts(octx, src, n, L, splitters, nsplit, cnts, stream)))
			return r;
		if ((r = comm->allgather(comm->ctx, cnts, cnts_all, sizeof(uint64_t) * RSX_MAX_RANKS)))
			return r;
		max_n_out = 0;
		for (int d = 0; d < world; ++d) {
			uint64_t tot = 0;
			dest_offset[d] = 0;
			for (int g = 0; g < world; ++g) {
				const uint64_t c = cnts_all[(size_t)g * RSX_MAX_RANKS + d];
				if (g < rank)
					dest_offset[d] += c;
				if (d == rank)
					recv_counts[g] = c;
				tot += c;
			}
			send_counts[d] = cnts[d];
			if (d == rank)
				my_n_out = tot;
			max_n_out = std::max(max_n_out, tot);
		}
		rep->key_range = 1;
	} else {
		memcpy(send_counts, route.send_counts, sizeof(send_counts));
		memcpy(recv_counts, route.recv_counts, sizeof(recv_counts));
		memcpy(dest_offset, route.dest_offset, sizeof(dest_offset));
		rep->routing_column = route.routing_column;
	}
	rep->imbalance = (double)max_n_out / std::max(1.0, (double)route.n_total / world);
	rep->needed_capacity = std::max<uint64_t>(max_n_out, *std::max_element(route.n_in, route.n_in + world));
	if (rep->needed_capacity > min_capacity)
		return RSX_ERR_WORKSPACE; // same verdict on every rank; src is still untouched
	rep->seconds_routing = since(t0);
	t0 = clk::now();

	// 4. partition + exchange
	void *sort_src, *sort_aux;
	const bool fused = recv_peers != nullptr && !(flags & RSX_MULTI_NO_FUSED);
	if (fused) {
		uint64_t base[RSX_MAX_RANKS];
		for (int d = 0; d < world; ++d)
			base[d] = (uint64_t)(uintptr_t)recv_peers[d] + dest_offset[d] * rb;
		if ((r = comm->barrier(comm->ctx))) // nobody still sorts out of its receive buffer from a previous call
			return r;
		if ((r = ops.partition_to(octx, src, n, L, rep->key_range ? -1 : route.routing_column, route.owner, splitters, nsplit,
		                          base, world, stream)))
			return r;
		if ((r = comm->barrier(comm->ctx))) // all remote stores have landed
			return r;
		sort_src = recv;
		sort_aux = src;
		rep->fused = 1;
	} else {
		if (!comm->alltoallv)
			return RSX_ERR_INVALID;
		// local stable partition into `recv` (destinations become contiguous ranges), then the
		// caller's all-to-all (NCCL under torchrun) back into `src`
		uint64_t base[RSX_MAX_RANKS], acc = 0, sbytes[RSX_MAX_RANKS], rbytes[RSX_MAX_RANKS];
		for (int d = 0; d < world; ++d) {
			base[d] = (uint64_t)(uintptr_t)recv + acc * rb;
			acc += send_counts[d];
			sbytes[d] = send_counts[d] * rb;
			rbytes[d] = recv_counts[d] * rb;
		}
		if ((r = ops.partition_to(octx, src, n, L, rep->key_range ? -1 : route.routing_column, route.owner, splitters, nsplit,
		                          base, world, stream)))
			return r;
		if ((r = comm->alltoallv(comm->ctx, recv, sbytes, src, rbytes)))
			return r;
		sort_src = src;
		sort_aux = recv;
	}
	rep->seconds_exchange = since(t0);
	t0 = clk::now();

	// 5. local LSD sort of the received records (chunks arrived in source-rank order: stable)
	*n_out = (size_t)my_n_out;
	*result = sort_src;
	if (my_n_out > 1 && (r = ops.sort(octx, sort_src, sort_aux, (size_t)my_n_out, L, result, stream)))
		return r;
	rep->seconds_local_sort = since(t0);
	return RSX_OK;
}

} // extern "C"

// ---- single process, one host thread per GPU -----------------------------------------------------------
namespace {

struct ThreadGroup {
	int world;
	std::mutex mu;
	std::condition_variable cv;
	int arrived = 0;
	unsigned long long generation = 0;
	std::vector<const void *> send;
	bool aborted = false; // a rank failed outside a collective: release everyone who waits
	explicit ThreadGroup(int w) : world(w), send(w) {}
	bool wait() {
		std::unique_lock<std::mutex> lk(mu);
		if (aborted)
			return false;
		const unsigned long long gen = generation;
		if (++arrived == world) {
			arrived = 0;
			++generation;
			cv.notify_all();
		} else {
			cv.wait(lk, [&] { return generation != gen || aborted; });
		}
		return !aborted;
	}
	void abort() {
		std::lock_guard<std::mutex> lk(mu);
		aborted = true;
		cv.notify_all();
	}
};
struct ThreadComm {
	ThreadGroup *grp;
	int rank;
	cudaStream_t stream;
};
int th_allgather(void *ctx, const void *send, void *recv, size_t bytes) {
	ThreadComm *c = static_cast<ThreadComm *>(ctx);
	c->grp->send[c->rank] = send;
	if (!c->grp->wait())
		return RSX_ERR_CUDA;
	for (int g = 0; g < c->grp->world; ++g)
		memcpy(static_cast<unsigned char *>(recv) + (size_t)g * bytes, c->grp->send[g], bytes);
	return c->grp->wait() ? RSX_OK : RSX_ERR_CUDA; // nobody overwrites its send buffer before everyone has copied
}
int th_barrier(void *ctx) {
	ThreadComm *c = static_cast<ThreadComm *>(ctx);
	const cudaError_t e = cudaStreamSynchronize(c->stream); // this device's stores are complete ...
	const bool ok = c->grp->wait();                         // ... on every device
	return (e == cudaSuccess && ok) ? RSX_OK : RSX_ERR_CUDA;
}

} // namespace

extern "C" int rsx_sort_multi(int ngpus, const int *devices, void *const *src, void *const *aux, const size_t *n,
                              size_t capacity, const rsx_layout *layout, uint32_t flags, void **result, size_t *n_out,
                              rsx_multi_report *reports) {
	if (ngpus < 1 || ngpus > RSX_MAX_RANKS || !devices || !src || !aux || !n || !layout || !result || !n_out)
		return RSX_ERR_INVALID;
	for (int g = 0; g < ngpus; ++g)
		if (n[g] > capacity || (capacity && (!src[g] || !aux[g])))
			return RSX_ERR_INVALID;
	int caller_device = 0;
	cudaGetDevice(&caller_device); // restored before returning: the calling thread's device is not ours to change
	// peer mappings: every device stores into every other device's receive buffer
	bool peers_ok = true;
	for (int g = 0; g < ngpus && peers_ok; ++g) {
		if (cudaSetDevice(devices[g]) != cudaSuccess) {
			cudaSetDevice(caller_device);
			return RSX_ERR_NO_DEVICE;
		}
		for (int d = 0; d < ngpus; ++d) {
			if (d == g || devices[d] == devices[g])
				continue;
			int can = 0;
			cudaDeviceCanAccessPeer(&can, devices[g], devices[d]);
			if (!can) {
				peers_ok = false;
				break;
			}
			const cudaError_t e = cudaDeviceEnablePeerAccess(devices[d], 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
				peers_ok = false;
			(void)cudaGetLastError();
		}
	}
	cudaSetDevice(caller_device);
	if (!peers_ok)
		return RSX_ERR_CUDA; // no NVLink / P2P between the devices: use the torchrun + NCCL path (dist.py)
	ThreadGroup grp(ngpus);
	std::vector<int> status(ngpus, RSX_OK);
	std::vector<std::thread> th;
	for (int g = 0; g < ngpus; ++g) {
		th.emplace_back([&, g] {
			cudaSetDevice(devices[g]);
			cudaStream_t st = nullptr;
			cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
			ThreadComm tc{&grp, g, st};
			rsx_comm comm{g, ngpus, th_allgather, th_barrier, nullptr, &tc};
			status[g] = rsx_sort_shard(&comm, nullptr, src[g], n[g], aux[g], aux, capacity, layout, flags & ~RSX_MULTI_NO_FUSED,
			                           &result[g], &n_out[g], reports ? &reports[g] : nullptr, st);
			if (status[g] != RSX_OK && status[g] != RSX_ERR_WORKSPACE)
				grp.abort(); // never leave the other threads waiting in a collective (WORKSPACE is collective)
			cudaStreamSynchronize(st);
			cudaStreamDestroy(st);
		});
	}
	for (auto &t : th)
		t.join();
	for (int g = 0; g < ngpus; ++g)
		if (status[g] != RSX_OK)
			return status[g];
	return RSX_OK;
}
