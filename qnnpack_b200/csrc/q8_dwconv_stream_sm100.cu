// q8dwconv 3x3, streaming dp4a kernel for sm_100a (stride 1 or 2, dilation 1, channels % 4 == 0).
//
// Replaces q8dwconv_ukernel_up8x9__sse2 (reference src/q8dwconv/up8x9-sse2.c:14-372) driven by
// src/operator-run.c:647-710; the indirection buffer (src/indirection.c:81-132) is never built.
//
// Same integers as the reference:  acc[c] = bias'[c] + sum_taps a_tap[c] * (w_tap[c] - kzp), padded taps read izp.
//
// Why this shape.  Depthwise layers move 2 bytes per output (stride 1) with 9 MACs each: at HBM speed an SM must
// retire ~11 outputs per clock, i.e. it has ~11 issue slots per output.  One IMAD per tap (9) plus byte
// extraction does not fit.  Here a thread owns 4 channels x 4 output columns and streams down the input rows:
//   * the 4 channel bytes of up to 9 neighbouring pixels are transposed (PRMT) into "3 taps of one channel"
//     words, so that ONE dp4a does the three horizontal taps of a kernel row;  (w - kzp) is 9-bit, so the
//     weights are split once at create time into two s8 halves (w-kzp = A + B) — or one operand when it fits;
//   * each input row is loaded once per thread and feeds the (up to) three output rows it belongs to, whose
//     partial sums rotate through registers; a finished row is requantised (fused Q31) and stored.
#include <cuda_runtime.h>
#include <stdint.h>

#include "q8_dwconv_sm100.cuh"
#include "requant_dev.cuh"
#include "sm100_ptx.cuh"

namespace q8 {

namespace {

constexpr int TX = 4;  // output columns per thread

// u8 activations x (s8 | u8) weights, 4-way dot product accumulate
template <bool W_UNSIGNED>
__device__ __forceinline__ int32_t dot4(uint32_t a, uint32_t w, int32_t acc) {
  int32_t r;
  if constexpr (W_UNSIGNED) {
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(w), "r"(acc));
  } else {
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(w), "r"(acc));
  }
  return r;
}

// 4x4 byte transpose: in: a,b,c,d = 4 pixels x 4 channels; out: t[ch] = (a[ch], b[ch], c[ch], d[ch])
__device__ __forceinline__ void transpose4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t (&t)[4]) {
  const uint32_t ab_lo = __byte_perm(a, b, 0x5140);  // a0 b0 a1 b1
  const uint32_t ab_hi = __byte_perm(a, b, 0x7362);  // a2 b2 a3 b3
  const uint32_t cd_lo = __byte_perm(c, d, 0x5140);
  const uint32_t cd_hi = __byte_perm(c, d, 0x7362);
  t[0] = __byte_perm(ab_lo, cd_lo, 0x5410);  // a0 b0 c0 d0
  t[1] = __byte_perm(ab_lo, cd_lo, 0x7632);  // a1 b1 c1 d1
  t[2] = __byte_perm(ab_hi, cd_hi, 0x5410);
  t[3] = __byte_perm(ab_hi, cd_hi, 0x7632);
}

// Load the NC = (TX-1)*S + 3 input pixels (4 channels each) one input row contributes to this thread's strip.
// rp = address of (row iy, column ix0) — may lie outside the image for padded rows/columns, in which case it is
// never dereferenced; colmask bit j = column ix0 + j is inside the image.
template <int S>
__device__ __forceinline__ void load_cols(const uint8_t* rp, const int (&coff)[(TX - 1) * S + 3], bool rowok, uint32_t colmask,
                                          uint32_t fill, uint32_t (&col)[(TX - 1) * S + 3]) {
  constexpr int NC = (TX - 1) * S + 3;
  constexpr uint32_t kAll = (1u << NC) - 1;
  if (rowok && colmask == kAll) {
#pragma unroll
    for (int j = 0; j < NC; j++) col[j] = __ldg(reinterpret_cast<const uint32_t*>(rp + coff[j]));
  } else {
#pragma unroll
    for (int j = 0; j < NC; j++)
      col[j] = (rowok && ((colmask >> j) & 1u)) ? __ldg(reinterpret_cast<const uint32_t*>(rp + coff[j])) : fill;
  }
}

// Windows of one input row: win[x][c] = bytes (pixel x*S + 0, +1, +2, <don't care>) of channel c.
template <int S>
__device__ __forceinline__ void build_windows(const uint32_t (&col)[(TX - 1) * S + 3], uint32_t (&win)[TX][4]) {
  uint32_t t0[4];
  transpose4(col[0], col[1], col[2], col[3], t0);
  if constexpr (S == 1) {
    // pixels 4,5 only as a pair per channel: u01 = (p4c0, p5c0, p4c1, p5c1), u23 = (p4c2, p5c2, p4c3, p5c3)
    const uint32_t u01 = __byte_perm(col[4], col[5], 0x5140);
    const uint32_t u23 = __byte_perm(col[4], col[5], 0x7362);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint32_t u = (c < 2) ? u01 : u23;
      win[0][c] = t0[c];                                                       // p0 p1 p2 (p3)
      win[1][c] = __byte_perm(t0[c], u, (c & 1) ? 0x6321 : 0x4321);            // p1 p2 p3 (p4)
      win[2][c] = __byte_perm(t0[c], u, (c & 1) ? 0x7632 : 0x5432);            // p2 p3 p4 p5
      win[3][c] = __byte_perm(t0[c], u, (c & 1) ? 0x0763 : 0x0543);            // p3 p4 p5 (x)
    }
  } else {
    uint32_t t1[4];
    transpose4(col[4], col[5], col[6], col[7], t1);
#pragma unroll
    for (int c = 0; c < 4; c++) {
      win[0][c] = t0[c];                                                       // p0 p1 p2 (p3)
      win[1][c] = __byte_perm(t0[c], t1[c], 0x0432);                           // p2 p3 p4 (x)
      win[2][c] = t1[c];                                                       // p4 p5 p6 (p7)
      win[3][c] = __byte_perm(t1[c], col[8], 0x0032 | ((4 + c) << 8));         // p6 p7 p8 (x)
    }
  }
}

// INIT: this is the first kernel row of a new output row -> start from the folded bias instead of accumulating
template <int WMODE, bool INIT>
__device__ __forceinline__ void accumulate_row(const uint32_t (&win)[TX][4], const uint32_t (&wa)[4], const uint32_t (&wb)[4],
                                               int32_t (&acc)[TX][4], const int4 bias) {
  const int32_t b[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
  for (int x = 0; x < TX; x++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      acc[x][c] = dot4<WMODE == 1>(win[x][c], wa[c], INIT ? b[c] : acc[x][c]);
      if constexpr (WMODE == 2) acc[x][c] = dot4<false>(win[x][c], wb[c], acc[x][c]);
    }
  }
}

// NEG: the weight operand holds kzp - w (WMODE 3), so the slots hold the NEGATED accumulators (started from the negated bias)
template <int RQ, bool NEG>
__device__ __forceinline__ void finish_row(const DwStreamParams& p, uint8_t* obase, int oy, int oy_end, int ox0,
                                           int32_t (&acc_in)[TX][4]) {
  if (oy >= 0 && oy < oy_end) {
    uint8_t* orow = obase + (size_t) oy * p.out_w * p.out_stride;
#pragma unroll
    for (int x = 0; x < TX; x++) {
      if (ox0 + x < p.out_w) {
        int32_t acc[TX][4];
#pragma unroll
        for (int c = 0; c < 4; c++) acc[x][c] = NEG ? -acc_in[x][c] : acc_in[x][c];
        uint32_t packed;
        if constexpr (RQ == 5 || RQ == 6) {
          auto rq1 = [&](int32_t nu) -> int32_t {
            int32_t y = q8_requant_u_unclamped((uint32_t) nu, p.rq.u_m2, p.rq.u_k2, p.rq.u_sm);
            if constexpr (RQ == 6) y = min(max(y, p.rq.qmin), p.rq.qmax);
            return y;
          };
          packed = pack_sat_u8x4(rq1(acc[x][0]), rq1(acc[x][1]), rq1(acc[x][2]), rq1(acc[x][3]));
        } else {
          packed = requant_pack4_generic(acc[x][0], acc[x][1], acc[x][2], acc[x][3], p.rq);
        }
        *reinterpret_cast<uint32_t*>(orow + (size_t) (ox0 + x) * p.out_stride) = packed;
      }
    }
  }
}

template <int S, int WMODE, int RQ>
__global__ void __launch_bounds__(128, 3) q8_dwconv3x3_stream_kernel(const __grid_constant__ DwStreamParams p) {
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total_threads) return;
  const int cg = (int) (idx % p.cgroups);
  long long r = idx / p.cgroups;
  const int xs = (int) (r % p.xstrips);
  r /= p.xstrips;
  const int yc = (int) (r % p.ychunks);
  const long long n = r / p.ychunks;
  const int c0 = cg * 4;
  const int ox0 = xs * TX;
  const int oy0 = yc * p.tyc;
  const int rows = min(p.tyc, p.out_h - oy0);

  // per-channel packed taps: word = (w[ky][0], w[ky][1], w[ky][2], 0) for channel c, operand A (and B)
  uint32_t wa[3][4], wb[3][4];
#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.wa + (size_t) ky * p.c_pad + c0));
    wa[ky][0] = a.x, wa[ky][1] = a.y, wa[ky][2] = a.z, wa[ky][3] = a.w;
    if constexpr (WMODE == 2) {
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.wb + (size_t) ky * p.c_pad + c0));
      wb[ky][0] = b.x, wb[ky][1] = b.y, wb[ky][2] = b.z, wb[ky][3] = b.w;
    } else {
      wb[ky][0] = wb[ky][1] = wb[ky][2] = wb[ky][3] = 0;
    }
  }
  int4 bias = __ldg(reinterpret_cast<const int4*>(p.bias + c0));
  if constexpr (RQ == 5 || RQ == 6) {  // "U" requantisation consumes n + 2^31 (mod 2^32): the offset rides on the bias
    bias.x ^= 0x80000000, bias.y ^= 0x80000000, bias.z ^= 0x80000000, bias.w ^= 0x80000000;
  }
  constexpr bool NEG = WMODE == 3;  // negated weight operand: accumulate -(bias + sum), negate once per output
  if constexpr (NEG) bias.x = -bias.x, bias.y = -bias.y, bias.z = -bias.z, bias.w = -bias.w;

  uint8_t* obase = p.out + (size_t) n * p.out_h * p.out_w * p.out_stride + c0;
  const uint32_t fill = (uint32_t) p.izp * 0x01010101u;
  const int ix0 = ox0 * S - p.pad_left;
  const int iy0 = oy0 * S - p.pad_top;
  constexpr int NC = (TX - 1) * S + 3;
  int coff[NC];
  uint32_t colmask = 0;
#pragma unroll
  for (int j = 0; j < NC; j++) {
    coff[j] = j * (int) p.in_stride;
    colmask |= ((unsigned) (ix0 + j) < (unsigned) p.in_w ? 1u : 0u) << j;
  }
  const long long row_pitch = (long long) p.in_w * p.in_stride;
  // address of (row iy0, column ix0); advanced by one row pitch per input row
  const uint8_t* rp = p.in + ((long long) n * p.in_h + iy0) * row_pitch + (long long) ix0 * p.in_stride + c0;
  const int oy_end = oy0 + rows;
  const int T = (rows - 1) * S + 3;  // input rows touched by this thread

  if constexpr (S == 1) {
    // input row t feeds output rows o = t - ky; o lives in slot o % 3; o completes at t = o + 2.
    int32_t acc[3][TX][4];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
      for (int x = 0; x < TX; x++) acc[s][x][0] = bias.x, acc[s][x][1] = bias.y, acc[s][x][2] = bias.z, acc[s][x][3] = bias.w;
    for (int t3 = 0; t3 < T; t3 += 3) {
#pragma unroll
      for (int v = 0; v < 3; v++) {
        const int t = t3 + v;
        if (t < T) {
          uint32_t col[NC];
          load_cols<1>(rp + (long long) t * row_pitch, coff, (unsigned) (iy0 + t) < (unsigned) p.in_h, colmask, fill, col);
          uint32_t win[TX][4];
          build_windows<1>(col, win);
          // kernel row 0 opens output row t (its slot restarts from the bias); rows 1, 2 continue rows t-1, t-2.
          // (For t < 2 the "continued" rows do not exist: their slots collect garbage that is overwritten by the
          //  next INIT before it could ever be stored.)
          accumulate_row<WMODE, true>(win, wa[0], wb[0], acc[v], bias);
          accumulate_row<WMODE, false>(win, wa[1], wb[1], acc[(v + 2) % 3], bias);
          accumulate_row<WMODE, false>(win, wa[2], wb[2], acc[(v + 1) % 3], bias);
          finish_row<RQ, NEG>(p, obase, t < 2 ? -1 : oy0 + t - 2, oy_end, ox0, acc[(v + 1) % 3]);
        }
      }
    }
  } else {
    // stride 2: even input row t = 2u feeds ky=0 of o=u and ky=2 of o=u-1 (which completes); odd t = 2u+1 feeds ky=1 of o=u.
    int32_t acc[2][TX][4];
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int x = 0; x < TX; x++) acc[s][x][0] = bias.x, acc[s][x][1] = bias.y, acc[s][x][2] = bias.z, acc[s][x][3] = bias.w;
    for (int t4 = 0; t4 < T; t4 += 4) {
#pragma unroll
      for (int v = 0; v < 4; v++) {
        const int t = t4 + v;
        if (t < T) {
          uint32_t col[NC];
          load_cols<2>(rp + (long long) t * row_pitch, coff, (unsigned) (iy0 + t) < (unsigned) p.in_h, colmask, fill, col);
          uint32_t win[TX][4];
          build_windows<2>(col, win);
          if ((v & 1) == 0) {
            accumulate_row<WMODE, true>(win, wa[0], wb[0], acc[v >> 1], bias);
            accumulate_row<WMODE, false>(win, wa[2], wb[2], acc[1 - (v >> 1)], bias);
            const int o = (t >> 1) - 1;
            finish_row<RQ, NEG>(p, obase, o < 0 ? -1 : oy0 + o, oy_end, ox0, acc[1 - (v >> 1)]);
          } else {
            accumulate_row<WMODE, false>(win, wa[1], wb[1], acc[v >> 1], bias);
          }
        }
      }
    }
  }
}

template <int S, int WMODE>
cudaError_t launch_rq(const DwStreamParams& p, cudaStream_t stream) {
  const int threads = 128;
  const unsigned blocks = (unsigned) ((p.total_threads + threads - 1) / threads);
  switch (p.rq_mode) {
    case 5: q8_dwconv3x3_stream_kernel<S, WMODE, 5><<<blocks, threads, 0, stream>>>(p); break;
    case 6: q8_dwconv3x3_stream_kernel<S, WMODE, 6><<<blocks, threads, 0, stream>>>(p); break;
    default: q8_dwconv3x3_stream_kernel<S, WMODE, 3><<<blocks, threads, 0, stream>>>(p); break;
  }
  return cudaGetLastError();
}

template <int S>
cudaError_t launch_wmode(const DwStreamParams& p, cudaStream_t stream) {
  switch (p.wmode) {
    case 0: return launch_rq<S, 0>(p, stream);
    case 1: return launch_rq<S, 1>(p, stream);
    case 3: return launch_rq<S, 3>(p, stream);
    default: return launch_rq<S, 2>(p, stream);
  }
}

}  // namespace

// Requirements (checked by the caller): 3x3, dilation 1, stride_h == stride_w in {1, 2}, channels % 4 == 0,
// input/output base and pixel strides multiples of 4.
cudaError_t launch_q8_dwconv3x3_stream(DwStreamParams p, cudaStream_t stream) {
  p.cgroups = p.channels / 4;
  p.xstrips = (p.out_w + TX - 1) / TX;
  // rows per thread: enough to amortise the 2-row halo, small enough to keep every SM busy on small images
  p.tyc = p.out_h >= 56 ? 16 : (p.out_h >= 14 ? 14 : p.out_h);
  p.ychunks = (p.out_h + p.tyc - 1) / p.tyc;
  p.total_threads = (long long) p.batch * p.ychunks * p.xstrips * p.cgroups;
  if (p.total_threads == 0) return cudaSuccess;
  return p.stride == 1 ? launch_wmode<1>(p, stream) : launch_wmode<2>(p, stream);
}

}  // namespace q8
