// Device form of the fused Q31 requantisation (see requant_math.h for the derivation and the host-checked forms).
#pragma once
#include <stdint.h>

#include "requant_math.h"

namespace q8 {

// RQ 0/1: fused, shift in [2,23]:  y = ( hi32(n*mult + {c_hi, c_lo | sign(n)}) + (n >> 31) ) >> (shift-1)
//         written so that the compiler emits LOP3 + IMAD.HI (64-bit addend pair) + LEA.HI + shift per value;
//         the shift runs as a multiply-high (FMA pipe) when shift_mul = 2^(33-shift) != 0.
//    2  : shift == 0      3: exact slow form (shift > 23)      4: fused, shift == 1
//    5/6: "U" form (requant_math.h): XOR + IMAD.WIDE.U32 (constant addend) + LEA.HI + IMAD.HI; 6 adds the clamp
// RQ 0 relies on the caller's u8-saturating pack for the clamp (qmin = 0, qmax = 255).
template <int RQ>
__device__ __forceinline__ int32_t requant_dev(int32_t n, const Q8Requant& rq, int32_t shift_mul) {
  if constexpr (RQ == 0 || RQ == 1) {
    const uint32_t lo = (uint32_t) rq.c_pos | ((uint32_t) n & 0x80000000u);
    const int64_t addend = (int64_t) (((uint64_t) (uint32_t) (rq.c_pos >> 32) << 32) | lo);
    const int32_t hi = (int32_t) (((int64_t) n * (int64_t) rq.multiplier + addend) >> 32);
    int32_t y = shift_mul != 0 ? __mulhi(hi + (n >> 31), shift_mul) : ((hi + (n >> 31)) >> (rq.shift - 1));
    if constexpr (RQ == 1) {
      y = max(y, rq.qmin);
      y = min(y, rq.qmax);
    }
    return y;
  } else if constexpr (RQ == 5 || RQ == 6) {
    int32_t y = q8_requant_u_unclamped((uint32_t) n ^ 0x80000000u, rq.u_m2, rq.u_k2, rq.u_sm);
    if constexpr (RQ == 6) {
      y = max(y, rq.qmin);
      y = min(y, rq.qmax);
    }
    return y;
  } else if constexpr (RQ == 2) {
    return q8_requant_shift0(n, rq.multiplier, rq.zero_point, rq.qmin, rq.qmax);
  } else if constexpr (RQ == 4) {
    int32_t y = q8_requant_fused_shift1_unclamped(n, rq.multiplier, rq.c_neg);
    y = max(y, rq.qmin);
    return min(y, rq.qmax);
  } else {
    return q8_requant(n, rq);  // RQ 3 = generic: picks the fused / shift-0 / shift-1 / exact form at run time
  }
}

// Generic requantise + pack of four values as ONE out-of-line function: the rarely used forms must not be inlined 32x
// into every epilogue instance (they made up most of the kernels' ~200 KB of code).
static __device__ __noinline__ uint32_t requant_pack4_generic(int32_t a, int32_t b, int32_t c, int32_t d, const Q8Requant& rq) {
  const int32_t ya = q8_requant(a, rq), yb = q8_requant(b, rq), yc = q8_requant(c, rq), yd = q8_requant(d, rq);
  return (uint32_t) ya | ((uint32_t) yb << 8) | ((uint32_t) yc << 16) | ((uint32_t) yd << 24);  // already clamped to u8
}

// host helper: 2^(33 - shift) when the final shift may run as a multiply-high, else 0
inline int32_t requant_shift_mul(const Q8Requant& rq) {
  return (rq.fused && rq.shift >= 3) ? (int32_t) (1u << (33 - rq.shift)) : 0;
}

}  // namespace q8
