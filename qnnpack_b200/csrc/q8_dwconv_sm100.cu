// q8dwconv (3x3 depthwise) and the generic direct convolution for sm_100a — CUDA-core, HBM-streaming kernels.
//
// Replaces (reference, paths relative to its root):
//   src/operator-run.c:647-710 + :238-254  dwconv case -> q8dwconv_ukernel_up8x9__sse2 (src/q8dwconv/up8x9-sse2.c:14-372)
//   src/indirection.c:81-132 (never materialised: tap -> address is computed in registers)
//   src/q8dwconv/mp8x25-sse2.c (5x5) and grouped q8conv (src/operator-run.c:805-844 with groups > 1) via the
//   direct kernel at the bottom.
//
// Arithmetic: acc[c] = bias'[c] + sum_taps a_tap[c] * (w_tap[c] - kzp), bias' = b + 9*izp*kzp - izp*sum w
// (src/qnnpack/pack.h:146-159), a padded tap reading the byte izp (src/convolution.c:336) — the
// reference's own form, so the integers are identical — followed by the fused Q31 epilogue.
//
// These layers have 3.6-9 int-ops per byte: no tensor cores, the job is to stream NHWC rows with
// coalesced accesses (a warp covers 128 consecutive channels = 128 B per tap) and to amortise tap
// loads over a strip of TX adjacent output pixels.
#include <cuda_runtime.h>
#include <stdint.h>

#include "q8_dwconv_sm100.cuh"
#include "sm100_ptx.cuh"

namespace q8 {

// RQ 5 / 6: "U" requantisation without / with clamp; anything else: the generic run-time form
template <int RQ>
__device__ __forceinline__ int32_t requant_one(int32_t n, const Q8Requant& rq) {
  if constexpr (RQ == 5 || RQ == 6) {
    int32_t t = q8_requant_u_unclamped((uint32_t) n ^ 0x80000000u, rq.u_m2, rq.u_k2, rq.u_sm);
    if constexpr (RQ == 6) t = min(max(t, rq.qmin), rq.qmax);
    return t;
  } else {
    return q8_requant(n, rq);
  }
}

// CV: channels per thread (4 -> 32-bit loads, 1 -> byte loads).
// SW: compile-time horizontal stride for the strip path (1 or 2, dilation_w == 1), 0 = fully generic.
// TX: output pixels per thread along x.
template <int CV, int SW, int TX, int RQ>
__global__ void __launch_bounds__(256) q8_dwconv3x3_kernel(const __grid_constant__ DwParams p) {
  constexpr int NCOL = (SW == 0) ? 3 : (TX - 1) * SW + 3;
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total_threads) return;
  const int cg = (int) (idx % p.cgroups);
  long long r = idx / p.cgroups;
  const int xs = (int) (r % p.xstrips);
  r /= p.xstrips;
  const int oy = (int) (r % p.out_h);
  const long long n = r / p.out_h;
  const int c0 = cg * CV;
  const int ox0 = xs * TX;

  // weights (w - kzp) as int32 [9][C], folded bias [C]
  int32_t w[9][CV], acc[TX][CV];
  if constexpr (CV == 4) {
#pragma unroll
    for (int t = 0; t < 9; t++) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(p.w32 + (size_t) t * p.c_pad + c0));
      w[t][0] = v.x, w[t][1] = v.y, w[t][2] = v.z, w[t][3] = v.w;
    }
    const int4 b = __ldg(reinterpret_cast<const int4*>(p.bias + c0));
#pragma unroll
    for (int x = 0; x < TX; x++) acc[x][0] = b.x, acc[x][1] = b.y, acc[x][2] = b.z, acc[x][3] = b.w;
  } else {
#pragma unroll
    for (int t = 0; t < 9; t++) w[t][0] = __ldg(p.w32 + (size_t) t * p.c_pad + c0);
    const int32_t b = __ldg(p.bias + c0);
#pragma unroll
    for (int x = 0; x < TX; x++) acc[x][0] = b;
  }

  const uint8_t* img = p.in + (size_t) n * p.in_h * p.in_w * p.in_stride + c0;
  const uint32_t fill = (uint32_t) p.izp * 0x01010101u;
  const int ix_base = ox0 * p.stride_w - p.pad_left;

#pragma unroll
  for (int ky = 0; ky < 3; ky++) {
    const int iy = oy * p.stride_h + ky * p.dil_h - p.pad_top;
    const bool rowok = (unsigned) iy < (unsigned) p.in_h;
    const uint8_t* rowp = img + (size_t) (rowok ? iy : 0) * p.in_w * p.in_stride;
    if constexpr (SW != 0) {
      // strip path: column j of the window serves tap (x, kx) with j = x*SW + kx
      uint32_t col[NCOL];
#pragma unroll
      for (int j = 0; j < NCOL; j++) {
        const int ix = ix_base + j;
        const bool ok = rowok && (unsigned) ix < (unsigned) p.in_w;
        if constexpr (CV == 4) {
          col[j] = ok ? __ldg(reinterpret_cast<const uint32_t*>(rowp + (size_t) ix * p.in_stride)) : fill;
        } else {
          col[j] = ok ? (uint32_t) __ldg(rowp + (size_t) ix * p.in_stride) : (uint32_t) p.izp;
        }
      }
#pragma unroll
      for (int x = 0; x < TX; x++) {
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const uint32_t v = col[x * SW + kx];
#pragma unroll
          for (int c = 0; c < CV; c++) acc[x][c] += (int32_t) ((v >> (8 * c)) & 0xFFu) * w[ky * 3 + kx][c];
        }
      }
    } else {
      // generic stride / dilation: TX == 1, every tap loaded on its own
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int ix = ix_base + kx * p.dil_w;
        const bool ok = rowok && (unsigned) ix < (unsigned) p.in_w;
        uint32_t v;
        if constexpr (CV == 4) {
          v = ok ? __ldg(reinterpret_cast<const uint32_t*>(rowp + (size_t) ix * p.in_stride)) : fill;
        } else {
          v = ok ? (uint32_t) __ldg(rowp + (size_t) ix * p.in_stride) : (uint32_t) p.izp;
        }
#pragma unroll
        for (int c = 0; c < CV; c++) acc[0][c] += (int32_t) ((v >> (8 * c)) & 0xFFu) * w[ky * 3 + kx][c];
      }
    }
  }

  uint8_t* orow = p.out + (((size_t) n * p.out_h + oy) * p.out_w) * p.out_stride + c0;
#pragma unroll
  for (int x = 0; x < TX; x++) {
    const int ox = ox0 + x;
    if (ox < p.out_w) {
      if constexpr (CV == 4) {
        const uint32_t packed = pack_sat_u8x4(requant_one<RQ>(acc[x][0], p.rq), requant_one<RQ>(acc[x][1], p.rq),
                                              requant_one<RQ>(acc[x][2], p.rq), requant_one<RQ>(acc[x][3], p.rq));
        *reinterpret_cast<uint32_t*>(orow + (size_t) ox * p.out_stride) = packed;
      } else {
        int32_t y = requant_one<RQ>(acc[x][0], p.rq);
        y = min(max(y, 0), 255);
        orow[(size_t) ox * p.out_stride] = (uint8_t) y;
      }
    }
  }
}

template <int CV, int SW, int TX>
static cudaError_t launch_dw_rq(const DwParams& p, cudaStream_t stream) {
  const int threads = 256;
  const long long blocks = (p.total_threads + threads - 1) / threads;
  switch (p.rq_mode) {
    case 5: q8_dwconv3x3_kernel<CV, SW, TX, 5><<<(unsigned) blocks, threads, 0, stream>>>(p); break;
    case 6: q8_dwconv3x3_kernel<CV, SW, TX, 6><<<(unsigned) blocks, threads, 0, stream>>>(p); break;
    default: q8_dwconv3x3_kernel<CV, SW, TX, 3><<<(unsigned) blocks, threads, 0, stream>>>(p); break;
  }
  return cudaGetLastError();
}

// Chooses the variant; fills the derived fields of `p` (cgroups, xstrips, total_threads).
cudaError_t launch_q8_dwconv3x3(DwParams p, int cv, cudaStream_t stream) {
  const bool strip = p.dil_w == 1 && (p.stride_w == 1 || p.stride_w == 2);
  const int tx = strip ? 4 : 1;
  p.cgroups = p.channels / cv;
  p.xstrips = (p.out_w + tx - 1) / tx;
  p.total_threads = (long long) p.batch * p.out_h * p.xstrips * p.cgroups;
  if (p.total_threads == 0) return cudaSuccess;
  if (cv == 4) {
    if (strip && p.stride_w == 1) return launch_dw_rq<4, 1, 4>(p, stream);
    if (strip && p.stride_w == 2) return launch_dw_rq<4, 2, 4>(p, stream);
    return launch_dw_rq<4, 0, 1>(p, stream);
  }
  if (strip && p.stride_w == 1) return launch_dw_rq<1, 1, 4>(p, stream);
  if (strip && p.stride_w == 2) return launch_dw_rq<1, 2, 4>(p, stream);
  return launch_dw_rq<1, 0, 1>(p, stream);
}

// ------------------------------------------------------------------------------------------------
// Generic direct convolution: one thread per output element.  Covers what has no fast path yet
// (grouped non-depthwise convolutions, 5x5 depthwise).  Same integers as every other path.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) q8_direct_conv_kernel(const __grid_constant__ DirectParams p) {
  const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.total) return;
  const int oc_all = p.groups * p.goc;
  const int och = (int) (idx % oc_all);
  long long r = idx / oc_all;
  const int ox = (int) (r % p.out_w);
  r /= p.out_w;
  const int oy = (int) (r % p.out_h);
  const long long n = r / p.out_h;
  const int g = och / p.goc;
  const uint8_t* wrow = p.w + (size_t) och * p.kh * p.kw * p.gic;
  int32_t acc = __ldg(p.bias + och);  // folded bias: padded taps read izp
  for (int ky = 0; ky < p.kh; ky++) {
    int iy = oy * p.stride_h + ky * p.dil_h - p.pad_top;
    bool oky = true;
    if (p.deconv) {  // src/indirection.c:165-178: y = oy + pad_top - ky*dil must be a non-negative multiple of the stride
      const int y = oy + p.pad_top - ky * p.dil_h;
      iy = y / p.stride_h;
      oky = y >= 0 && iy * p.stride_h == y;
    }
    for (int kx = 0; kx < p.kw; kx++) {
      int ix = ox * p.stride_w + kx * p.dil_w - p.pad_left;
      bool okx = true;
      if (p.deconv) {
        const int x = ox + p.pad_left - kx * p.dil_w;
        ix = x / p.stride_w;
        okx = x >= 0 && ix * p.stride_w == x;
      }
      const bool ok = oky && okx && (unsigned) iy < (unsigned) p.in_h && (unsigned) ix < (unsigned) p.in_w;
      const uint8_t* a = p.in + (((size_t) n * p.in_h + (ok ? iy : 0)) * p.in_w + (ok ? ix : 0)) * p.in_stride +
          (size_t) g * p.gic;
      const uint8_t* wk = wrow + (size_t) (ky * p.kw + kx) * p.gic;
      for (int c = 0; c < p.gic; c++) {
        const int32_t av = ok ? (int32_t) __ldg(a + c) : p.izp;
        acc += av * ((int32_t) __ldg(wk + c) - p.kzp);
      }
    }
  }
  int32_t y = q8_requant(acc, p.rq);
  p.out[(((size_t) n * p.out_h + oy) * p.out_w + ox) * p.out_stride + och] = (uint8_t) y;
}

cudaError_t launch_q8_direct_conv(const DirectParams& p, cudaStream_t stream) {
  if (p.total == 0) return cudaSuccess;
  const int threads = 256;
  const long long blocks = (p.total + threads - 1) / threads;
  q8_direct_conv_kernel<<<(unsigned) blocks, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stand-alone Q31 requantization (int32 -> uint8): the epilogue as its own kernel, used to pin the
// device arithmetic against the reference's requantization tests (test/requantization.cc Q31 cases).
// ------------------------------------------------------------------------------------------------
__global__ void q8_requantize_kernel(const int32_t* __restrict__ in, uint8_t* __restrict__ out, long long n, Q8Requant rq) {
  const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint8_t) q8_requant(in[i], rq);
}

cudaError_t launch_q8_requantize(const int32_t* in, uint8_t* out, long long n, const Q8Requant& rq, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int threads = 256;
  q8_requantize_kernel<<<(unsigned) ((n + threads - 1) / threads), threads, 0, stream>>>(in, out, n, rq);
  return cudaGetLastError();
}

}  // namespace q8
