/*
 * rsx_oracle.c -- CPU restatement of the reference's 8-bit-digit LSD radix sort.
 *
 * TEST INFRASTRUCTURE ONLY (see rsx_oracle.h).  Type-erased plain C: records are moved with
 * memcpy, the key is `key_bytes` little-endian bytes at `key_offset`, and the reference's
 * KeyFunc is described by (kdf_kind, flags) instead of a C++ callable.
 *
 * What is deliberately NOT restated: the reference picks the histogram counter width from n
 * (radix_sort.hpp:100-113: u8 / u16 / u32 / u64).  The width never changes a result (counts
 * are <= n and n fits the chosen type), so 64-bit counters are used throughout.
 */
#include "rsx_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ---- key derivation ---------------------------------------------------------------- */

static inline uint64_t width_mask(uint32_t key_bytes) {
	return key_bytes >= 8 ? ~0ULL : ((1ULL << (8u * key_bytes)) - 1ULL);
}

/* radix_sort_basic_kdf.hpp:19-23 (unsigned: identity), :26-30 (signed: value ^ highbit<T>()),
 * :32-38 / :40-46 (float / double: bits ^ (-(bits >> msb) | 1 << msb)); descending order is
 * the bitwise complement of the derived key (README.md:564-574, radix_tests.cpp:111-113,175-177). */
uint64_t orc_kdf(const void *rec, const orc_layout *L) {
	uint64_t k = 0;
	memcpy(&k, (const unsigned char *)rec + L->key_offset, L->key_bytes);
	const uint64_t top = 1ULL << (8u * L->key_bytes - 1u);
	const uint64_t m = width_mask(L->key_bytes);
	if (L->kdf_kind == ORC_KDF_SIGNED) {
		k ^= top;
	} else if (L->kdf_kind == ORC_KDF_FLOAT) {
		k ^= (k & top) ? m : top;
	}
	if (L->flags & ORC_FLAG_INVERT)
		k = ~k & m;
	return k;
}

static inline unsigned digit_of(uint64_t key, unsigned col) {
	return (unsigned)((key >> (8u * col)) & 0xFFu);
}

/* Shared front half of rs_sort_main / rs_sort_rank: one pass over the input that fills all
 * wc histograms and counts ordered neighbours (radix_sort.hpp:46-58, radix_sort_rank.hpp:41-53),
 * then the column probe (radix_sort.hpp:64-70) and the exclusive scans (:72-80). */
static void survey_columns(const unsigned char *src, size_t n, const orc_layout *L,
                           uint64_t *hist /* wc*256, zeroed */, orc_report *rep,
                           uint64_t *hist_copy) {
	const unsigned wc = L->key_bytes;
	const size_t rb = L->record_bytes;
	uint64_t n_unsorted = n;
	uint64_t cur = orc_kdf(src, L);
	for (size_t i = 0; i < n; ++i) {
		uint64_t nxt = 0;
		if (i + 1 < n) {
			nxt = orc_kdf(src + (i + 1) * rb, L);
			if (cur <= nxt)
				--n_unsorted;
		}
		for (unsigned c = 0; c < wc; ++c)
			++hist[256u * c + digit_of(cur, c)];
		cur = nxt;
	}
	if (hist_copy)
		memcpy(hist_copy, hist, sizeof(uint64_t) * 256u * wc);

	rep->n_unsorted = n_unsorted;
	rep->ncols = 0;
	rep->result_in_aux = 0;
	rep->early_exit = n_unsorted < 2;
	if (rep->early_exit)
		return;

	const uint64_t first = orc_kdf(src, L);
	for (unsigned c = 0; c < wc; ++c) {
		if (hist[256u * c + digit_of(first, c)] != n)
			rep->cols[rep->ncols++] = c;
	}
	for (unsigned i = 0; i < rep->ncols; ++i) {
		uint64_t *h = hist + 256u * rep->cols[i];
		uint64_t run = 0;
		for (unsigned d = 0; d < 256u; ++d) {
			uint64_t cnt = h[d];
			h[d] = run;
			run += cnt;
		}
	}
}

/* ---- value sort: rs_sort_main, radix_sort.hpp:31-93 ---------------------------------- */

void *orc_radix_sort(void *src_v, void *aux_v, size_t n, const orc_layout *L,
                     uint64_t *hist_out, orc_report *rep) {
	orc_report local;
	if (!rep)
		rep = &local;
	memset(rep, 0, sizeof(*rep));
	const unsigned wc = L->key_bytes;
	if (hist_out)
		memset(hist_out, 0, sizeof(uint64_t) * 256u * wc);
	if (n < 2) { /* radix_sort.hpp:37-38,100-101 */
		rep->early_exit = 1;
		rep->n_unsorted = n;
		return src_v; /* the reference never touches the histograms here */
	}
	uint64_t hist[8 * 256];
	memset(hist, 0, sizeof(hist));
	unsigned char *src = (unsigned char *)src_v;
	unsigned char *aux = (unsigned char *)aux_v;
	const size_t rb = L->record_bytes;

	survey_columns(src, n, L, hist, rep, hist_out);
	if (rep->early_exit)
		return src; /* radix_sort.hpp:60-62: nothing was written */

	/* radix_sort.hpp:82-90: one stable counting-sort pass per live column, LSD first. */
	for (unsigned i = 0; i < rep->ncols; ++i) {
		const unsigned col = rep->cols[i];
		uint64_t *offs = hist + 256u * col;
		for (size_t j = 0; j < n; ++j) {
			const unsigned char *r = src + j * rb;
			uint64_t dst = offs[digit_of(orc_kdf(r, L), col)]++;
			memcpy(aux + dst * rb, r, rb);
		}
		unsigned char *t = src;
		src = aux;
		aux = t;
	}
	rep->result_in_aux = rep->ncols & 1u; /* radix_sort.hpp:89,92 */
	return src;
}

/* ---- rank sort: rs_sort_rank, radix_sort_rank.hpp:22-92 ------------------------------ */

static inline uint64_t idx_load(const void *base, size_t i, int idx_bytes) {
	uint64_t v = 0;
	memcpy(&v, (const unsigned char *)base + i * (size_t)idx_bytes, (size_t)idx_bytes);
	return v;
}
static inline void idx_store(void *base, size_t i, int idx_bytes, uint64_t v) {
	memcpy((unsigned char *)base + i * (size_t)idx_bytes, &v, (size_t)idx_bytes);
}

void *orc_radix_sort_rank(const void *src_v, void *index_buffer, size_t n, const orc_layout *L,
                          int idx_bytes, int as_shipped, orc_report *rep) {
	orc_report local;
	if (!rep)
		rep = &local;
	memset(rep, 0, sizeof(*rep));
	if (n < 2) { /* radix_sort_rank.hpp:28-32 */
		if (n)
			idx_store(index_buffer, 0, idx_bytes, 0);
		rep->early_exit = 1;
		rep->n_unsorted = n;
		return index_buffer;
	}
	uint64_t hist[8 * 256];
	memset(hist, 0, sizeof(hist));
	const unsigned char *src = (const unsigned char *)src_v;
	const size_t rb = L->record_bytes;

	survey_columns(src, n, L, hist, rep, NULL);
	/* radix_sort_rank.hpp:52: identity permutation written during the histogram loop
	 * (truncated to IdxType exactly like the reference's implicit conversion). */
	for (size_t i = 0; i < n; ++i)
		idx_store(index_buffer, i, idx_bytes, (uint64_t)i);
	if (rep->early_exit)
		return index_buffer; /* :55-57 */

	unsigned char *from = (unsigned char *)index_buffer;                      /* :77 */
	unsigned char *to = (unsigned char *)index_buffer + n * (size_t)idx_bytes; /* :78 */
	for (unsigned i = 0; i < rep->ncols; ++i) {
		const unsigned col = rep->cols[i];
		uint64_t *offs = hist + 256u * col;
		for (size_t j = 0; j < n; ++j) {
			const uint64_t id = idx_load(from, j, idx_bytes);
			/* :82 reads src[j]; the intended algorithm (radix_sort_u32_ranks.c:98,104) reads
			 * the element the index points at. */
			const unsigned char *r = src + (as_shipped ? j : (size_t)id) * rb;
			uint64_t dst = offs[digit_of(orc_kdf(r, L), col)]++;
			idx_store(to, dst, idx_bytes, id);
		}
		unsigned char *t = from;
		from = to;
		to = t;
	}
	rep->result_in_aux = rep->ncols & 1u; /* :91 */
	return from;
}

/* ---- independent cross-checks: bottom-up merge sort ---------------------------------- */

void orc_stable_sort(void *data_v, void *tmp_v, size_t n, const orc_layout *L) {
	unsigned char *a = (unsigned char *)data_v, *b = (unsigned char *)tmp_v;
	const size_t rb = L->record_bytes;
	int flipped = 0;
	for (size_t w = 1; w < n; w *= 2) {
		for (size_t lo = 0; lo < n; lo += 2 * w) {
			size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
			size_t i = lo, j = mid, o = lo;
			while (i < mid && j < hi) {
				/* take from the right run only when strictly smaller: ties keep input order */
				if (orc_kdf(a + j * rb, L) < orc_kdf(a + i * rb, L))
					memcpy(b + (o++) * rb, a + (j++) * rb, rb);
				else
					memcpy(b + (o++) * rb, a + (i++) * rb, rb);
			}
			if (i < mid)
				memcpy(b + o * rb, a + i * rb, (mid - i) * rb), o += mid - i;
			if (j < hi)
				memcpy(b + o * rb, a + j * rb, (hi - j) * rb);
		}
		unsigned char *t = a;
		a = b;
		b = t;
		flipped ^= 1;
	}
	if (flipped)
		memcpy(data_v, a, n * rb);
}

void orc_stable_argsort(const void *src_v, uint64_t *idx, uint64_t *tmp, size_t n,
                        const orc_layout *L) {
	const unsigned char *src = (const unsigned char *)src_v;
	const size_t rb = L->record_bytes;
	uint64_t *a = idx, *b = tmp;
	for (size_t i = 0; i < n; ++i)
		a[i] = i;
	int flipped = 0;
	for (size_t w = 1; w < n; w *= 2) {
		for (size_t lo = 0; lo < n; lo += 2 * w) {
			size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
			size_t i = lo, j = mid, o = lo;
			while (i < mid && j < hi) {
				if (orc_kdf(src + a[j] * rb, L) < orc_kdf(src + a[i] * rb, L))
					b[o++] = a[j++];
				else
					b[o++] = a[i++];
			}
			while (i < mid)
				b[o++] = a[i++];
			while (j < hi)
				b[o++] = a[j++];
		}
		uint64_t *t = a;
		a = b;
		b = t;
		flipped ^= 1;
	}
	if (flipped)
		memcpy(idx, a, n * sizeof(uint64_t));
}
