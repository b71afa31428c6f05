/*
 * TEST INFRASTRUCTURE — CPU restatement of the reference's q8 hot path.  NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object; the product (qnnpack_b200/csrc) never links or calls it.
 *
 * Parity pin: this restatement is checked byte-for-byte against the UNMODIFIED reference compiled
 * into oracle/_ref/libqnnpack_ref.so (tests/test_oracle.py, wherever oracle/_ref has been built) and against the committed fixtures in tests/golden/ that were generated from that
 * compiled reference (tests/golden/make_golden.py).  The reference has no on-disk golden vectors
 * of its own (all its tests draw from std::random_device); its deterministic known-answer tests
 * for Q31 (test/requantization-tester.h:84-246) are restated in tests/test_oracle.py.
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Arithmetic is written in plain scalar C, one output element at a time, in the form the
 * reference's micro-kernels use:   acc = packed_bias + sum_k a_k * (w_k - kernel_zero_point)
 * with padded taps reading the byte `input_zero_point`, then the Q31 requantization.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define Q8O_API __attribute__((visibility("default")))

/* The .scalar view of the reference's parameter union: src/qnnpack/params.h:128-138 */
struct q8o_requant_params {
  int32_t multiplier;
  int32_t remainder_mask;
  int32_t remainder_threshold;
  uint32_t shift;
  int32_t min_less_zero_point;
  int32_t max_less_zero_point;
  int32_t zero_point;
};

static inline uint32_t f32_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, sizeof u);
  return u;
}

/* arithmetic shift right without relying on implementation-defined behaviour
 * (src/qnnpack/scalar-utils.h:41-60) */
static inline int32_t asr32(int32_t x, uint32_t n) {
  return x >= 0 ? (int32_t) ((uint32_t) x >> n) : (int32_t) ~(~(uint32_t) x >> n);
}

/* src/qnnpack/requantization.h:122-198 (scalar branch :184-196) and :22-54.
 * Only the bit pattern of `scale` is used: 24-bit mantissa << 7 is the Q31 multiplier,
 * the exponent gives the post-shift. */
Q8O_API void q8o_compute_requant_params(
    float scale, uint8_t zero_point, uint8_t qmin, uint8_t qmax, struct q8o_requant_params* p) {
  const uint32_t bits = f32_bits(scale);
  p->multiplier = (int32_t) (((bits & UINT32_C(0x007FFFFF)) | UINT32_C(0x00800000)) << 7);
  const int32_t shift = 127 + 31 - 32 - (int32_t) (bits >> 23);
  const uint32_t mask = (UINT32_C(1) << shift) - UINT32_C(1);
  p->remainder_mask = (int32_t) mask;
  p->remainder_threshold = (int32_t) (mask >> 1);
  p->shift = (uint32_t) shift;
  p->min_less_zero_point = (int32_t) qmin - (int32_t) zero_point;
  p->max_less_zero_point = (int32_t) qmax - (int32_t) zero_point;
  p->zero_point = (int32_t) zero_point;
}

/* src/qnnpack/requantization.h:464-480 — the function every micro-kernel test uses as its
 * expected value (test/gemm-microkernel-tester.h:259,376; test/dwconv-microkernel-tester.h:249). */
Q8O_API uint8_t q8o_q31_requantize(int32_t n, const struct q8o_requant_params* p) {
  const int64_t product = (int64_t) n * (int64_t) p->multiplier;
  const int32_t q31 = (int32_t) (uint32_t) ((uint64_t) (product + INT64_C(0x40000000)) >> 31);
  const int32_t rem = (q31 & p->remainder_mask) - (int32_t) (n < 0);
  int32_t y = asr32(q31, p->shift) + (int32_t) (rem > p->remainder_threshold);
  if (y < p->min_less_zero_point) y = p->min_less_zero_point;
  if (y > p->max_less_zero_point) y = p->max_less_zero_point;
  return (uint8_t) (y + p->zero_point);
}

/* Array form with the signature of the stand-alone requantizers
 * (src/qnnpack/requantization-stubs.h:22-29; src/requantization/q31-scalar.c:17-138). */
Q8O_API void q8o_requantize_q31(
    size_t n, const int32_t* input, float scale, uint8_t zero_point, uint8_t qmin, uint8_t qmax,
    uint8_t* output) {
  struct q8o_requant_params p;
  q8o_compute_requant_params(scale, zero_point, qmin, qmax, &p);
  for (size_t i = 0; i < n; i++) output[i] = q8o_q31_requantize(input[i], &p);
}

/* src/convolution.c:29-37 */
static inline size_t out_dim(size_t padded_in, size_t k, size_t dil, size_t stride) {
  const size_t eff = (k - 1) * dil + 1;
  return (padded_in - eff) / stride + 1;
}

Q8O_API size_t q8o_output_dim(size_t in, size_t pad_a, size_t pad_b, size_t k, size_t dil, size_t stride) {
  return out_dim(pad_a + in + pad_b, k, dil, stride);
}

/*
 * Whole-operator restatement of qnnp_create_convolution2d_nhwc_q8 + setup + run.
 *
 *  - bias folding: packed_bias[oc] = bias[oc] + K*izp*kzp - izp*sum_k w[oc][k], in int32
 *      (src/qnnpack/pack.h:24,43 gemm; :63,84 conv; :146,159 depthwise), K = KH*KW*GIC.
 *  - tap -> input coordinate and the unsigned "in bounds" test: src/indirection.c:56-63 (conv),
 *      :105-116 (depthwise); padded taps read a buffer filled with input_zero_point
 *      (src/convolution.c:336, src/indirection.c:64,71).
 *  - accumulate a*(w - kzp) in int32: src/q8gemm/4x4c2-sse2.c:47-109, src/q8conv/4x4c2-sse2.c:33-138,
 *      src/q8dwconv/up8x9-sse2.c:43-125.
 *  - kernel layout [G][GOC][KH][KW][GIC] (src/qnnpack/pack.h:79; depthwise [C][KH][KW] :158),
 *      output index ((n*OH+oy)*OW+ox)*out_stride + g*GOC + oc  (src/operator-run.c:60-69,206-216).
 *  - bytes of an output pixel beyond G*GOC are never written (SURVEY.md §7 trap 7).
 *
 * The reference picks one of three micro-kernel families by shape (src/convolution.c:180-189);
 * all three compute the same integers, so one loop nest restates all of them.
 * Unsigned arithmetic is used for the accumulator so that wrap-around (unreachable for supported K)
 * is defined behaviour.
 * Returns 0 on success, 1 if the requantization scale is outside [2^-32, 1).
 */
Q8O_API int q8o_convolution2d_nhwc_q8(
    size_t batch, size_t in_h, size_t in_w,
    uint32_t pad_top, uint32_t pad_right, uint32_t pad_bottom, uint32_t pad_left,
    uint32_t kh, uint32_t kw, uint32_t stride_h, uint32_t stride_w, uint32_t dil_h, uint32_t dil_w,
    uint32_t groups, size_t gic, size_t goc,
    uint8_t izp, float input_scale, uint8_t kzp, float kernel_scale,
    const uint8_t* kernel, const int32_t* bias,
    uint8_t ozp, float output_scale, uint8_t qmin, uint8_t qmax,
    const uint8_t* input, size_t in_stride, uint8_t* output, size_t out_stride) {
  /* fp32, in exactly this order: src/convolution.c:161 */
  const float scale = input_scale * kernel_scale / output_scale;
  if (!(scale < 1.0f) || !(scale >= 0x1.0p-32f)) return 1;
  struct q8o_requant_params rp;
  q8o_compute_requant_params(scale, ozp, qmin, qmax, &rp);

  const size_t out_h = out_dim(pad_top + in_h + pad_bottom, kh, dil_h, stride_h);
  const size_t out_w = out_dim(pad_left + in_w + pad_right, kw, dil_w, stride_w);
  const size_t ks = (size_t) kh * kw;
  const int32_t boff = (int32_t) ((uint32_t) (ks * gic) * (uint32_t) izp * (uint32_t) kzp);

  for (size_t g = 0; g < groups; g++) {
    for (size_t oc = 0; oc < goc; oc++) {
      const uint8_t* w = kernel + (g * goc + oc) * ks * gic;
      uint32_t wsum = 0;
      for (size_t i = 0; i < ks * gic; i++) wsum += w[i];
      const uint32_t packed_bias = (uint32_t) bias[g * goc + oc] + (uint32_t) boff - wsum * (uint32_t) izp;

      for (size_t n = 0; n < batch; n++) {
        for (size_t oy = 0; oy < out_h; oy++) {
          for (size_t ox = 0; ox < out_w; ox++) {
            uint32_t acc = packed_bias;
            for (size_t ky = 0; ky < kh; ky++) {
              const size_t iy = oy * stride_h + ky * dil_h - pad_top; /* wraps when in the padding */
              for (size_t kx = 0; kx < kw; kx++) {
                const size_t ix = ox * stride_w + kx * dil_w - pad_left;
                const uint8_t* a = (iy < in_h && ix < in_w)
                    ? input + ((n * in_h + iy) * in_w + ix) * in_stride + g * gic
                    : NULL;
                const uint8_t* wk = w + (ky * kw + kx) * gic;
                for (size_t ic = 0; ic < gic; ic++) {
                  const int32_t av = a ? (int32_t) a[ic] : (int32_t) izp;
                  acc += (uint32_t) (av * ((int32_t) wk[ic] - (int32_t) kzp));
                }
              }
            }
            output[((n * out_h + oy) * out_w + ox) * out_stride + g * goc + oc] =
                q8o_q31_requantize((int32_t) acc, &rp);
          }
        }
      }
    }
  }
  return 0;
}

/* qnnp_create_fully_connected_nc_q8 + setup + run: src/fully-connected.c:25-161 —
 * a gemm operator with groups=1, M=batch (:149-158); kernel is [OC][IC] (pack.h:36). */
Q8O_API int q8o_fully_connected_nc_q8(
    size_t batch, size_t ic, size_t oc,
    uint8_t izp, float input_scale, uint8_t kzp, float kernel_scale,
    const uint8_t* kernel, const int32_t* bias,
    uint8_t ozp, float output_scale, uint8_t qmin, uint8_t qmax,
    const uint8_t* input, size_t in_stride, uint8_t* output, size_t out_stride) {
  return q8o_convolution2d_nhwc_q8(
      1, batch, 1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, ic, oc, izp, input_scale, kzp, kernel_scale, kernel,
      bias, ozp, output_scale, qmin, qmax, input, in_stride, output, out_stride);
}

/* int32 accumulators only (no requantization): what test/gemm-microkernel-tester.h:213-226 calls
 * `acc` — used by tests that derive the output scale from the accumulator range (:236-241). */
Q8O_API void q8o_gemm_accumulators(
    size_t m, size_t n, size_t k, const uint8_t* a, size_t a_stride, const uint8_t* b /* [n][k] */,
    const int32_t* bias, uint8_t azp, uint8_t bzp, int32_t* acc /* [m][n] */) {
  for (size_t i = 0; i < m; i++)
    for (size_t j = 0; j < n; j++) {
      uint32_t s = (uint32_t) bias[j];
      for (size_t t = 0; t < k; t++)
        s += (uint32_t) (((int32_t) a[i * a_stride + t] - (int32_t) azp) * ((int32_t) b[j * k + t] - (int32_t) bzp));
      acc[i * n + j] = (int32_t) s;
    }
}
