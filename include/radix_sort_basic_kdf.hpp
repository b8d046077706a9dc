/*
	Key-derivation functions for the B200 radix sort -- drop-in for the reference's
	radix_sort_basic_kdf.hpp (same namespace, same `kdf` overload set, same `highbit`).

	The reference passes the KDF as an arbitrary callable that is inlined into the sort loops
	(radix_sort.hpp:31-35).  A CUDA kernel cannot call a host lambda, so in addition to the
	host-callable overloads (which behave exactly like the reference's, radix_sort_basic_kdf.hpp:13-46)
	this header provides *descriptor functors*: small callable types that compute the same derived
	key on the host AND describe it to the device path as an rsx_layout.  They cover every
	derivation the reference ships, documents or tests:

	    basic_kdfs::ascending            default order, any arithmetic T     (radix_sort_basic_kdf.hpp:19-46)
	    basic_kdfs::descending           ~kdf(v)                              (README.md:564-574, radix_tests.cpp:175-177)
	    basic_kdfs::by_member<&R::key>   record sorted by one of its members  (radix_tests.cpp:41-43)
	    basic_kdfs::by_member<&R::key, basic_kdfs::desc>                      (radix_tests.cpp:111-113, by value)

	A callable the device cannot interpret (e.g. one that dereferences host pointers,
	radix_tests.cpp:111-113) is rejected at compile time; there is no silent CPU fallback.
*/
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "rsx.h"

namespace basic_kdfs {

namespace detail {
// std::make_unsigned<T> is ill-formed for non-integral T even inside enable_if, so select lazily
template <typename T, bool = std::is_integral_v<T> && !std::is_same_v<T, bool>> struct unsigned_of { using type = void; };
template <typename T> struct unsigned_of<T, true> { using type = std::make_unsigned_t<T>; };
template <typename T> using unsigned_of_t = typename unsigned_of<T>::type;
} // namespace detail

// Unsigned T with only the most significant bit set (reference: radix_sort_basic_kdf.hpp:13-17).
template <typename T>
constexpr std::enable_if_t<std::is_integral_v<T>, detail::unsigned_of_t<T>> highbit(void) {
	return static_cast<detail::unsigned_of_t<T>>(detail::unsigned_of_t<T>(1) << (sizeof(T) * 8 - 1));
}

// unsigned integers: the key is the value (reference :19-23); bool is excluded like upstream
template <typename T>
std::enable_if_t<std::is_integral_v<T> && std::is_unsigned_v<T> && !std::is_same_v<T, bool>, T> kdf(const T &value) {
	return value;
}

// signed integers: flip the sign bit (reference :26-30)
template <typename T>
std::enable_if_t<std::is_integral_v<T> && std::is_signed_v<T> && !std::is_same_v<T, bool>, detail::unsigned_of_t<T>>
kdf(const T &value) {
	return static_cast<detail::unsigned_of_t<T>>(static_cast<detail::unsigned_of_t<T>>(value) ^ highbit<T>());
}

// float / double: negative -> all bits flipped, positive -> sign bit set (reference :32-46)
template <typename T> std::enable_if_t<std::is_same_v<T, float>, uint32_t> kdf(const T &value) {
	uint32_t bits;
	std::memcpy(&bits, &value, sizeof bits);
	return bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
template <typename T> std::enable_if_t<std::is_same_v<T, double>, uint64_t> kdf(const T &value) {
	uint64_t bits;
	std::memcpy(&bits, &value, sizeof bits);
	return bits ^ ((bits >> 63) ? ~0ULL : (1ULL << 63));
}

// ---- descriptor functors -------------------------------------------------------------------------

namespace detail {
template <typename K> constexpr uint32_t kdf_kind_of() {
	static_assert(std::is_arithmetic_v<K> && !std::is_same_v<K, bool>, "key must be an integer, float or double");
	if constexpr (std::is_floating_point_v<K>) {
		static_assert(sizeof(K) == 4 || sizeof(K) == 8, "only float and double have a KDF (like the reference)");
		return RSX_KDF_FLOAT;
	} else {
		return std::is_signed_v<K> ? RSX_KDF_SIGNED : RSX_KDF_UNSIGNED;
	}
}
template <typename M> struct member_traits;
template <typename R, typename K> struct member_traits<K R::*> {
	using record = R;
	using key = K;
};
} // namespace detail

enum order { asc = 0, desc = 1 };

// Whole element is the key, reference's default order.
struct ascending {
	template <typename T> auto operator()(const T &v) const { return kdf(v); }
	template <typename T> static rsx_layout layout() {
		return rsx_layout{(uint32_t)sizeof(T), 0u, (uint32_t)sizeof(T), detail::kdf_kind_of<T>(), 0u};
	}
};

// Whole element is the key, descending: the complement of the derived key (README.md:564-574).
struct descending {
	template <typename T> auto operator()(const T &v) const {
		auto k = kdf(v);
		return static_cast<decltype(k)>(~k);
	}
	template <typename T> static rsx_layout layout() {
		return rsx_layout{(uint32_t)sizeof(T), 0u, (uint32_t)sizeof(T), detail::kdf_kind_of<T>(), RSX_FLAG_INVERT};
	}
};

// Record sorted by one member: basic_kdfs::by_member<&sortrec::key>{}.
template <auto Member, order Order = asc> struct by_member {
	using R = typename detail::member_traits<decltype(Member)>::record;
	using K = typename detail::member_traits<decltype(Member)>::key;
	auto operator()(const R &r) const {
		auto k = kdf(r.*Member);
		return Order == desc ? static_cast<decltype(k)>(~k) : k;
	}
	template <typename T> static rsx_layout layout() {
		static_assert(std::is_same_v<T, R>, "by_member<> used with a different record type");
		static_assert(std::is_trivially_copyable_v<R>, "records are moved bytewise on the device");
		alignas(R) static unsigned char probe[sizeof(R)];
		const R *r = reinterpret_cast<const R *>(probe);
		const size_t off = reinterpret_cast<const unsigned char *>(&(r->*Member)) - probe;
		return rsx_layout{(uint32_t)sizeof(R), (uint32_t)off, (uint32_t)sizeof(K), detail::kdf_kind_of<K>(),
		                  Order == desc ? RSX_FLAG_INVERT : 0u};
	}
};

} // namespace basic_kdfs
