/*
	8xW-bit LSD rank (argsort) radix sort on a B200 -- drop-in for the reference's
	radix_sort_rank.hpp (radix_sort_rank.hpp:97-112):

	    IdxType* radix_sort_rank(const T* src, IdxType* index_buffer, size_t n[, kf]);

	src is never written.  index_buffer must hold 2n entries (README.md:520-526); the returned
	pointer is index_buffer or index_buffer + n by the parity of the live columns
	(radix_sort_rank.hpp:77-91); on early exit the identity permutation is in the first half
	(:52-57).  IdxType may be any 1-, 2-, 4- or 8-byte unsigned integer.

	The ranks R satisfy src[R[0]] <= src[R[1]] <= ... by derived key with ties in input order --
	the semantics the reference documents (README.md:485-490) and its listing
	radix_sort_u32_ranks.c implements.  NOTE: the reference header as shipped only does so for
	inputs with at most one non-trivial column (radix_sort_rank.hpp:82 looks the key up by
	position instead of through the index); see DESIGN.md.
*/
#pragma once

#include "radix_sort.hpp"

template <typename T, typename IdxType, typename KeyFunc = basic_kdfs::ascending,
          int passes = sizeof(std::invoke_result_t<std::remove_reference_t<KeyFunc> &, const T &>)>
IdxType *radix_sort_rank(const T *RESTRICT src, IdxType *RESTRICT index_buffer, size_t n, KeyFunc &&kf = KeyFunc{}) {
	using F = std::remove_cv_t<std::remove_reference_t<KeyFunc>>;
	static_assert(rsx_detail::is_descriptor<F, T>::value,
	              "radix_sort_rank on the GPU needs a basic_kdfs descriptor functor as key derivation "
	              "(include/radix_sort_basic_kdf.hpp); there is no CPU fallback");
	static_assert(std::is_integral_v<IdxType> && std::is_unsigned_v<IdxType>, "IdxType must be an unsigned integer");
	(void)kf;
	const rsx_layout layout = F::template layout<T>();
	void *result = nullptr;
	const int st = rsx_sort_rank(src, index_buffer, n, &layout, (int)sizeof(IdxType), &result, nullptr,
	                             rsx_detail::stream_slot());
	rsx_detail::last_status() = st;
	return st == RSX_OK ? static_cast<IdxType *>(result) : nullptr;
}
