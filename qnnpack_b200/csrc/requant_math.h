// Fixed-point (Q31) requantization, the epilogue of every q8 kernel in this library.
//
// Specification (reference, paths relative to its root):
//   src/qnnpack/requantization.h:464-480  qnnp_q31_requantize  (what every micro-kernel test checks)
//   src/requantization/q31-scalar.c:17-138, src/q8gemm/4x4c2-sse2.c:178-278 (inlined SSE2 form)
//   parameter derivation: src/qnnpack/requantization.h:122-198
//
//     q = (int32) ((int64(n) * multiplier + 2^30) >> 31)            (rounds ties up)
//     r = (q & mask) - (n < 0);  y = (q >> shift) + (r > mask >> 1)  (rounds ties away from zero)
//     out = clamp(y, qmin - zp, qmax - zp) + zp
//
// B200 form.  The epilogue is ALU-bound on write-heavy layers (DESIGN.md §epilogue budget), so the
// two roundings are collapsed into ONE 64-bit multiply-add and ONE shift, exactly:
//
//     floor((floor(P / 2^31) + h) / 2^s) == floor((P + h * 2^31) / 2^(31+s))       for integer h
//   with P = n*multiplier + 2^30 and h = (s > 0 ? 2^(s-1) : 0) - (s > 0 && n < 0), which is the
//   "add half, minus one for negatives" form of round-half-away-from-zero.  Adding zp * 2^(31+s) to
//   the addend folds the "+ zero_point" in as well.  Hence
//
//     out = clamp( (int64(n) * multiplier + (n < 0 ? c_neg : c_pos)) >> (31 + s), qmin, qmax )
//
//   c_pos = 2^30 + [s>0] 2^(30+s) + zp * 2^(31+s),   c_neg = c_pos - [s>0] 2^31.
// The 64-bit sum cannot overflow for s <= 23 (|n*mult| < 2^62, zp*2^(31+s) <= 255*2^54); for larger
// shifts (scale < 2^-24, outputs pinned next to the zero point) q8_requant_exact_slow() is used.
// tests/test_requant_math.py checks both forms against the oracle over edge and random values.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define Q8_HD __host__ __device__ __forceinline__
#else
#define Q8_HD static inline
#endif

struct Q8Requant {
  int32_t multiplier;  // [2^30, 2^31 - 128]
  int32_t shift;       // s in [0, 31]
  int32_t zero_point;
  int32_t qmin;        // output_min (already includes the zero point, i.e. a uint8 code)
  int32_t qmax;
  int64_t c_pos;       // fused addends (valid when shift <= 23)
  int64_t c_neg;
  int32_t fused;       // 1 if the fused form may be used
  int32_t u_ok;        // 1 if the "U" form below may be used (needs a bound on |n|, see q8_requant_enable_u)
  uint32_t u_m2;       // 2 * multiplier
  int32_t u_sm;        // 2^(32 - shift): the final arithmetic shift as a multiply-high
  uint64_t u_k2;       // 2*c_pos - 2^32*multiplier - 2^32   (mod 2^64)
};

// Host-side parameter derivation from the fp32 scale's bit pattern (requantization.h:131-141).
Q8_HD Q8Requant q8_make_requant(uint32_t scale_bits, uint8_t zero_point, uint8_t qmin, uint8_t qmax) {
  Q8Requant r;
  r.multiplier = (int32_t) (((scale_bits & 0x007FFFFFu) | 0x00800000u) << 7);
  r.shift = 127 + 31 - 32 - (int32_t) (scale_bits >> 23);
  r.zero_point = zero_point;
  r.qmin = qmin;
  r.qmax = qmax;
  r.fused = (r.shift >= 0 && r.shift <= 23) ? 1 : 0;
  r.c_pos = 0;
  r.c_neg = 0;
  if (r.fused) {
    const int s = r.shift;
    r.c_pos = (int64_t(1) << 30) + (s > 0 ? (int64_t(1) << (30 + s)) : 0) + (int64_t(zero_point) << (31 + s));
    r.c_neg = r.c_pos - (s > 0 ? (int64_t(1) << 31) : 0);
  }
  r.u_ok = 0;
  r.u_m2 = 0;
  r.u_sm = 0;
  r.u_k2 = 0;
  return r;
}

// "U" form: the cheapest exact evaluation (4 integer instructions per value on the GPU: XOR, IMAD.WIDE.U32 with a
// CONSTANT 64-bit addend, LEA.HI, IMAD.HI), available when the kernel can bound its accumulators, |n| <= nmax.
//   X = n*mult + c_pos - [n<0]*2^31 is what the fused form shifts right by 31+s.  Write n = nu - 2^31 with
//   nu = n XOR 2^31 (unsigned) and b = nu >> 31 = [n >= 0]:
//       2X = nu*(2 mult) + b*2^32 + K2,     K2 = 2 c_pos - 2^32 mult - 2^32
//       floor(X / 2^31) = hi32(nu*(2 mult) + K2) + b          (b*2^32 is a multiple of 2^32; K2 taken mod 2^64)
//       out = floor(X / 2^31) >> s  =  mulhi(floor(X / 2^31), 2^(32-s))          for s >= 2
//   The sign-dependent rounding term became the carry-free "+ b", and the 64-bit addend no longer depends on n.
//   hi32 is taken modulo 2^32, which is exact when floor(X / 2^31) fits in int32 — guaranteed by the nmax check.
Q8_HD void q8_requant_enable_u(Q8Requant& r, int64_t nmax) {
  r.u_ok = 0;
  if (!r.fused || r.shift < 2 || nmax < 0 || nmax > 0x7FFFFFFFll) return;
  const int64_t bound = ((nmax * (int64_t) r.multiplier) >> 31) + (r.c_pos >> 31) + 2;
  if (bound >= 0x7FFFFFFFll) return;
  r.u_m2 = (uint32_t) r.multiplier << 1;
  r.u_sm = (int32_t) (1u << (32 - r.shift));
  r.u_k2 = 2ull * (uint64_t) r.c_pos - ((uint64_t) (uint32_t) r.multiplier << 32) - (1ull << 32);
  r.u_ok = 1;
}

// nu = n XOR 0x80000000.  Returns y + zero_point, unclamped.
Q8_HD int32_t q8_requant_u_unclamped(uint32_t nu, uint32_t m2, uint64_t k2, int32_t sm) {
  const uint32_t hi = (uint32_t) (((uint64_t) nu * m2 + k2) >> 32);
  const int32_t t = (int32_t) (hi + (nu >> 31));
#if defined(__CUDA_ARCH__)
  return __mulhi(t, sm);
#else
  return (int32_t) (((int64_t) t * (int64_t) sm) >> 32);
#endif
}

// Straight transcription of the specification; used for shift > 23 and as the in-library cross-check.
Q8_HD int32_t q8_requant_exact_slow(int32_t n, const Q8Requant& p) {
  const int64_t product = (int64_t) n * (int64_t) p.multiplier;
  const int32_t q31 = (int32_t) (uint32_t) ((uint64_t) (product + 0x40000000ll) >> 31);
  const uint32_t mask = (p.shift >= 32) ? 0xFFFFFFFFu : ((1u << p.shift) - 1u);
  const int32_t rem = (int32_t) ((uint32_t) q31 & mask) - (int32_t) (n < 0);
  const int32_t thr = (int32_t) (mask >> 1);
  // arithmetic shift without UB (src/qnnpack/scalar-utils.h:41-60)
  const int32_t sh = p.shift > 31 ? 31 : p.shift;
  const int32_t asr = q31 >= 0 ? (int32_t) ((uint32_t) q31 >> sh) : (int32_t) ~(~(uint32_t) q31 >> sh);
  int32_t y = asr + (int32_t) (rem > thr);
  const int32_t lo = p.qmin - p.zero_point, hi = p.qmax - p.zero_point;
  y = y < lo ? lo : y;
  y = y > hi ? hi : y;
  return y + p.zero_point;
}

// Fused form for shift >= 2, WITHOUT the final clamp: returns y + zero_point as int32 (may lie outside
// [0,255]; cannot overflow because the high word of the 64-bit sum fits in int32).
// The sign-dependent addend is built without a predicate: c_neg = c_pos - 2^31 and, for s >= 2, the low
// word of c_pos is exactly 0x40000000 (bit 31 clear), so
//     addend = { hi: c_pos_hi + (n >> 31),  lo: c_pos_lo | (n & 0x80000000) }
// i.e. one logic op and one add instead of compare + two selects.
Q8_HD int32_t q8_requant_fused_unclamped(int32_t n, int32_t multiplier, int64_t c_pos, int32_t shift_m1) {
  const uint32_t lo = (uint32_t) c_pos | ((uint32_t) n & 0x80000000u);
  const int32_t hi = (int32_t) (c_pos >> 32) + (n >> 31);
  const int64_t c = (int64_t) (((uint64_t) (uint32_t) hi << 32) | lo);
  const int64_t p = (int64_t) n * (int64_t) multiplier + c;
  // (p >> (32 + (s-1))): only the high word matters
  return ((int32_t) (p >> 32)) >> shift_m1;
}

// shift == 1: here it is c_neg whose low word is 0x40000000, so 2^31 is ADDED for n >= 0 instead.
Q8_HD int32_t q8_requant_fused_shift1_unclamped(int32_t n, int32_t multiplier, int64_t c_neg) {
  const uint32_t lo = (uint32_t) c_neg | (~(uint32_t) n & 0x80000000u);
  const int64_t c = (int64_t) (((uint64_t) (c_neg >> 32) << 32) | lo);
  const int64_t p = (int64_t) n * (int64_t) multiplier + c;
  return (int32_t) (p >> 32);
}

// shift == 0 (scale in [0.5, 1)): the second rounding is the identity, and y + zp could exceed
// int32 for |n| near 2^31, so the zero point is added after the upper clamp.  Fully clamped result.
Q8_HD int32_t q8_requant_shift0(int32_t n, int32_t multiplier, int32_t zero_point, int32_t qmin, int32_t qmax) {
  const int64_t p = (int64_t) n * (int64_t) multiplier + 0x40000000ll;
  int32_t y = (int32_t) (p >> 31);
  const int32_t hi = qmax - zero_point;
  y = y > hi ? hi : y;
  y += zero_point;
  return y < qmin ? qmin : y;
}

// 0: fused shift in [2,23], clamp provided by the u8 saturating pack; 1: same + explicit clamp;
// 2: shift == 0; 3: exact slow form (shift > 23); 4: fused shift == 1;
// 5: "U" form, clamp provided by the u8 saturating pack; 6: "U" form + explicit clamp
Q8_HD int q8_requant_mode(const Q8Requant& p) {
  if (!p.fused) return 3;
  if (p.u_ok) return (p.qmin == 0 && p.qmax == 255) ? 5 : 6;
  if (p.shift == 0) return 2;
  if (p.shift == 1) return 4;
  return (p.qmin == 0 && p.qmax == 255) ? 0 : 1;
}

Q8_HD int32_t q8_requant(int32_t n, const Q8Requant& p) {
  if (!p.fused) return q8_requant_exact_slow(n, p);
  if (p.shift == 0) return q8_requant_shift0(n, p.multiplier, p.zero_point, p.qmin, p.qmax);
  int32_t y = p.u_ok ? q8_requant_u_unclamped((uint32_t) n ^ 0x80000000u, p.u_m2, p.u_k2, p.u_sm) : p.shift == 1 ? q8_requant_fused_shift1_unclamped(n, p.multiplier, p.c_neg)
                           : q8_requant_fused_unclamped(n, p.multiplier, p.c_pos, p.shift - 1);
  y = y < p.qmin ? p.qmin : y;
  y = y > p.qmax ? p.qmax : y;
  return y;
}
