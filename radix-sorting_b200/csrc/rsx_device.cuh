// rsx_device.cuh -- device-side helpers: record types, key extraction, key derivation, digits.
#pragma once

#include "rsx_internal.cuh"

namespace rsx {

// ---- record storage types -------------------------------------------------------------------
template <int ES> struct Rec;
template <> struct Rec<1> { using type = uint8_t; };
template <> struct Rec<2> { using type = uint16_t; };
template <> struct Rec<4> { using type = uint32_t; };
template <> struct Rec<8> { using type = unsigned long long; };
template <> struct Rec<16> { using type = ulonglong2; };

template <int PL> struct Payload;
template <> struct Payload<0> { using type = uint32_t; }; // unused
template <> struct Payload<4> { using type = uint32_t; };
template <> struct Payload<8> { using type = unsigned long long; };

// The aligned 64-bit word (or the whole record, zero-extended) that contains the key.
template <int ES>
__device__ __forceinline__ unsigned long long key_word(const typename Rec<ES>::type &r, uint32_t word_sel) {
	if constexpr (ES == 16)
		return word_sel ? r.y : r.x;
	else
		return (unsigned long long)r;
}

__device__ __forceinline__ unsigned long long width_mask(uint32_t key_bytes) {
	return key_bytes >= 8 ? ~0ULL : ((1ULL << (8u * key_bytes)) - 1ULL);
}

// Derived sort key.  radix_sort_basic_kdf.hpp:19-23 (unsigned), :26-30 (signed, ^ highbit),
// :32-46 (float/double: ^ (-(b >> msb) | 1 << msb)); descending = complement (README.md:564-574).
__device__ __forceinline__ unsigned long long derive_key(unsigned long long word, const KeyDesc &kd) {
	const unsigned long long m = width_mask(kd.key_bytes);
	unsigned long long k = (word >> kd.key_shift) & m;
	const unsigned long long top = 1ULL << (8u * kd.key_bytes - 1u);
	if (kd.kdf_kind == RSX_KDF_SIGNED)
		k ^= top;
	else if (kd.kdf_kind == RSX_KDF_FLOAT)
		k ^= (k & top) ? m : top;
	if (kd.invert)
		k = ~k & m;
	return k;
}

// ---- per-pass digit extraction ----------------------------------------------------------------
// For one column the derivation collapses to: digit = raw_byte ^ xor_const [^ sign-dependent mask].
//   unsigned : xor_const = 0
//   signed   : xor_const = 0x80 on the top column, else 0
//   float    : sign clear -> like signed; sign set -> raw_byte ^ 0xFF
//   invert   : xor_const ^= 0xFF
struct DigitDesc {
	uint32_t word_sel;   // 16 B records: which 8-byte word
	uint32_t bit_shift;  // bit position of the column's byte inside the 64-bit key word
	uint32_t xor_const;
	uint32_t float_mask; // 0 if not a float KDF, else (0xFF ^ top_const): applied when sign set
	uint32_t sign_shift; // bit position of the key's sign bit inside the 64-bit key word
};

__host__ __device__ inline DigitDesc make_digit_desc(const KeyDesc &kd, int col) {
	DigitDesc dd;
	dd.word_sel = kd.word_sel;
	dd.bit_shift = kd.key_shift + 8u * (uint32_t)col;
	const bool top = (uint32_t)col == kd.key_bytes - 1u;
	const uint32_t top_const = (kd.kdf_kind != RSX_KDF_UNSIGNED && top) ? 0x80u : 0u;
	dd.xor_const = top_const ^ (kd.invert ? 0xFFu : 0u);
	dd.float_mask = kd.kdf_kind == RSX_KDF_FLOAT ? (0xFFu ^ top_const) : 0u;
	dd.sign_shift = kd.key_shift + 8u * kd.key_bytes - 1u;
	return dd;
}

template <int ES, bool FLOAT>
__device__ __forceinline__ uint32_t digit_of(const typename Rec<ES>::type &r, const DigitDesc &dd) {
	uint32_t d;
	if constexpr (ES <= 4) {
		const uint32_t w = (uint32_t)r;
		d = ((w >> dd.bit_shift) & 0xFFu) ^ dd.xor_const;
		if constexpr (FLOAT) {
			const uint32_t s = (uint32_t)((int32_t)(w << (31u - dd.sign_shift)) >> 31);
			d ^= s & dd.float_mask;
		}
	} else {
		const unsigned long long w = key_word<ES>(r, dd.word_sel);
		d = ((uint32_t)(w >> dd.bit_shift) & 0xFFu) ^ dd.xor_const;
		if constexpr (FLOAT) {
			const uint32_t hi = (uint32_t)(w >> (dd.sign_shift & 32u)); // the half holding the sign
			const uint32_t s = (uint32_t)((int32_t)(hi << (31u - (dd.sign_shift & 31u))) >> 31);
			d ^= s & dd.float_mask;
		}
	}
	return d;
}

// ---- small utilities --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() {
	uint32_t l;
	asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
	return l;
}
__device__ __forceinline__ uint32_t lanemask_lt() {
	uint32_t m;
	asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
	z ^= z >> 30;
	z *= 0xBF58476D1CE4E5B9ULL;
	z ^= z >> 27;
	z *= 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return z;
}

} // namespace rsx
