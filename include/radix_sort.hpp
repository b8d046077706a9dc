/*
	8xW-bit LSD radix sort on a B200 -- drop-in for the reference's radix_sort.hpp.

	Same call syntax and return convention as the reference (radix_sort.hpp:98-115):

	    T* radix_sort(T* src, T* aux, size_t n);              // default KDF for T
	    T* radix_sort(T* src, T* aux, size_t n, kf);          // kf: a basic_kdfs descriptor functor

	src and aux hold n elements each and are both clobbered; the returned pointer aliases src if
	the number of non-trivial 8-bit columns is even (or the input was already ordered), else aux
	(radix_sort.hpp:60-62,89-92).  The result is bit-identical to the reference's: a stable sort
	by derived key.

	What differs, by necessity:
	  * the work runs in librsx.so (CUDA, sm_100a) through the C ABI in rsx.h; device / managed
	    pointers are sorted in place on the GPU, plain host pointers are staged H2D/D2H;
	  * `kf` must be describable to the device: basic_kdfs::ascending (default), ::descending,
	    ::by_member<&Rec::key[, basic_kdfs::desc]> (radix_sort_basic_kdf.hpp).  Any other callable is a
	    compile-time error -- there is no CPU fallback;
	  * the reference cannot fail; this returns nullptr if the device call fails and
	    radix_sort_last_status() tells why.
*/
#pragma once

#include <cstddef>
#include <type_traits>
#include <utility>

#include "radix_sort_basic_kdf.hpp"
#include "rsx.h"

#ifndef RESTRICT
#define RESTRICT __restrict__
#endif

namespace rsx_detail {
template <typename F, typename T, typename = void> struct is_descriptor : std::false_type {};
template <typename F, typename T>
struct is_descriptor<F, T, std::void_t<decltype(F::template layout<T>())>> : std::true_type {};

inline int &last_status() {
	static thread_local int status = RSX_OK;
	return status;
}
inline void *&stream_slot() {
	static thread_local void *stream = nullptr;
	return stream;
}
} // namespace rsx_detail

// rsx_status of the last radix_sort / radix_sort_rank call on this thread.
inline int radix_sort_last_status() { return rsx_detail::last_status(); }
// cudaStream_t used by subsequent calls on this thread (default: the null stream).
inline void radix_sort_set_stream(void *cuda_stream) { rsx_detail::stream_slot() = cuda_stream; }

template <typename T, typename KeyFunc = basic_kdfs::ascending,
          int passes = sizeof(std::invoke_result_t<std::remove_reference_t<KeyFunc> &, const T &>)>
T *radix_sort(T *RESTRICT src, T *RESTRICT aux, size_t n, KeyFunc &&kf = KeyFunc{}) {
	using F = std::remove_cv_t<std::remove_reference_t<KeyFunc>>;
	static_assert(rsx_detail::is_descriptor<F, T>::value,
	              "radix_sort on the GPU needs a key-derivation the device can evaluate: use "
	              "basic_kdfs::ascending / descending / by_member<&Rec::key> instead of an arbitrary callable "
	              "(include/radix_sort_basic_kdf.hpp); there is no CPU fallback");
	static_assert(passes >= 1 && passes <= 8, "KeyType must be 64 bits or less (radix_sort.hpp:34)");
	(void)kf;
	const rsx_layout layout = F::template layout<T>();
	void *result = nullptr;
	const int st = rsx_sort(src, aux, n, &layout, &result, nullptr, rsx_detail::stream_slot());
	rsx_detail::last_status() = st;
	return st == RSX_OK ? static_cast<T *>(result) : nullptr;
}
