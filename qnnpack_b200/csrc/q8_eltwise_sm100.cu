// Element-wise and pooling operators of qnnpack.h for sm_100a: HBM-streaming CUDA-core kernels with the reference's
// fixed-point arithmetic restated exactly (every result is checked byte for byte against the compiled reference).
//
// Replaces (reference, paths relative to its root):
//   add_nc_q8                     src/add.c:22-149,  ukernel src/q8vadd/sse2.c,  spec qnnp_add_quantize requantization.h:500-522
//   global_average_pooling_nwc_q8 src/global-average-pooling.c:22-147, src/q8gavgpool/*, spec qnnp_avgpool_quantize :482-498
//   average_pooling2d_nhwc_q8     src/average-pooling.c:36-275, src/q8avgpool/*  (same quantisation, padding reads izp)
//   max_pooling2d_nhwc_u8         src/max-pooling.c:36-223, src/u8maxpool/*      (max over the taps inside the image, clamp)
//   clamp_nc_u8                   src/clamp.c, src/u8clamp/*
//   sigmoid_nc_q8 / leaky_relu_nc_q8   src/sigmoid.c, src/leaky-relu.c -> 256-entry table, src/x8lut/scalar.c
//   softargmax_nc_q8              src/softargmax.c, src/operator-run.c:625-637, src/u8rmax/*, src/u8lut32norm/scalar.c
//   channel_shuffle_nc_x8         src/channel-shuffle.c, src/x8zip/*
// None of these has data reuse or a contraction: the roofline is HBM bandwidth (bytes in + bytes out), the design rule is
// 16-byte coalesced accesses whenever base, strides and channel count allow, 4-byte or 1-byte pieces otherwise.
#include <cuda_runtime.h>
#include <stdint.h>

#include "q8_eltwise_sm100.cuh"

namespace q8 {
namespace {

constexpr int kThreads = 256;

inline unsigned blocks_for(long long work) { return (unsigned) ((work + kThreads - 1) / kThreads); }

// ---- fixed-point helpers -------------------------------------------------------------------------------------------
// requantization.h:500-522 (scalar form; the SSE2 ukernel is tested bit-exact against it, test/q8vadd.cc)
__device__ __forceinline__ uint32_t add_quantize(uint32_t a, uint32_t b, const AddParams& p) {
  int32_t acc = (int32_t) ((uint32_t) p.zero_point_product + a * p.a_multiplier + b * p.b_multiplier);
  const int32_t rem = (acc & p.remainder_mask) - (int32_t) (acc < 0);
  acc = (acc >> p.shift) + (int32_t) (rem > p.remainder_threshold);
  int32_t y = acc + p.y_zero_point;
  y = y >= p.y_max ? p.y_max : y;
  y = y <= p.y_min ? p.y_min : y;
  return (uint32_t) y;
}

// The same value with the rounding written as one add and one shift, valid while acc + 2^(shift-1) cannot overflow
// (shift <= 22: |acc| <= 255 * (2^22 - 1) * 2 < 2^31 - 2^21):  (acc >> s) + [rem > threshold] == (acc + 2^(s-1) - [acc < 0]) >> s.
// Returns acc' + zero point, unclamped (the caller saturates four of them into a word, then clamps the word).
__device__ __forceinline__ int32_t add_quantize_fast(uint32_t a, uint32_t b, const AddParams& p) {
  const int32_t acc = (int32_t) ((uint32_t) p.zero_point_product + a * p.a_multiplier + b * p.b_multiplier);
  return ((acc + p.remainder_threshold + 1 + (acc >> 31)) >> p.shift) + p.y_zero_point;
}

__device__ __forceinline__ uint32_t pack_sat4(int32_t a, int32_t b, int32_t c, int32_t d) {
  uint32_t hi, out;
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(d), "r"(c));
  asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(out) : "r"(b), "r"(a), "r"(hi));
  return out;
}

// requantization.h:482-498
__device__ __forceinline__ uint32_t avgpool_quantize(int32_t n, const AvgQuant& q) {
  const int64_t product = (int64_t) n * (int64_t) q.multiplier;
  const int64_t adjusted = product - (int64_t) (n < 0);
  int32_t y = (int32_t) ((adjusted + q.rounding) >> q.right_shift);
  y = y < q.min_less_zp ? q.min_less_zp : y;
  y = y > q.max_less_zp ? q.max_less_zp : y;
  return (uint32_t) (y + q.zero_point);
}

template <int VEC>
__device__ __forceinline__ void load_bytes(const uint8_t* p, uint32_t (&w)[VEC >= 4 ? VEC / 4 : 1]) {
  if constexpr (VEC == 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
  } else if constexpr (VEC == 4) {
    w[0] = *reinterpret_cast<const uint32_t*>(p);
  } else {
    w[0] = *p;
  }
}
template <int VEC>
__device__ __forceinline__ void store_bytes(uint8_t* p, const uint32_t (&w)[VEC >= 4 ? VEC / 4 : 1]) {
  if constexpr (VEC == 16) {
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  } else if constexpr (VEC == 4) {
    *reinterpret_cast<uint32_t*>(p) = w[0];
  } else {
    *p = (uint8_t) w[0];
  }
}

// work item -> (row, byte offset inside the row)
__device__ __forceinline__ void split_item(long long i, int pieces_per_row, int vec, long long& row, int& off) {
  row = i / pieces_per_row;
  off = (int) (i - row * pieces_per_row) * vec;
}

// ---- add ------------------------------------------------------------------------------------------------------------
template <int VEC, bool FAST>
__global__ void __launch_bounds__(kThreads) q8_add_kernel(const __grid_constant__ AddParams p) {
  const long long i = (long long) blockIdx.x * kThreads + threadIdx.x;
  if (i >= p.rows * p.pieces_per_row) return;
  long long row;
  int off;
  split_item(i, p.pieces_per_row, VEC, row, off);
  constexpr int NW = VEC >= 4 ? VEC / 4 : 1;
  uint32_t a[NW], b[NW], y[NW];
  load_bytes<VEC>(p.a + row * p.a_stride + off, a);
  load_bytes<VEC>(p.b + row * p.b_stride + off, b);
#pragma unroll
  for (int w = 0; w < NW; w++) {
    if constexpr (VEC == 1) {
      y[w] = add_quantize(a[w], b[w], p);
    } else if constexpr (FAST) {
      int32_t r[4];
#pragma unroll
      for (int k = 0; k < 4; k++) r[k] = add_quantize_fast(__byte_perm(a[w], 0, 0x4440 + k), __byte_perm(b[w], 0, 0x4440 + k), p);
      // saturate to [0, 255] while packing, then clamp the four bytes at once: identical to clamping each int32 to
      // [y_min, y_max] because 0 <= y_min < y_max <= 255
      y[w] = __vminu4(__vmaxu4(pack_sat4(r[0], r[1], r[2], r[3]), (uint32_t) p.y_min * 0x01010101u), (uint32_t) p.y_max * 0x01010101u);
    } else {
      y[w] = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) y[w] |= add_quantize((a[w] >> (8 * k)) & 0xFF, (b[w] >> (8 * k)) & 0xFF, p) << (8 * k);
    }
  }
  store_bytes<VEC>(p.y + row * p.y_stride + off, y);
}

// ---- byte-wise maps: clamp, lookup table -------------------------------------------------------------------------------
template <int VEC, bool LUT>
__global__ void __launch_bounds__(kThreads) q8_map_kernel(const __grid_constant__ MapParams p) {
  __shared__ uint8_t lut[256];
  if constexpr (LUT) {
    lut[threadIdx.x] = p.lut[threadIdx.x];  // kThreads == 256
    __syncthreads();
  }
  const long long i = (long long) blockIdx.x * kThreads + threadIdx.x;
  if (i >= p.rows * p.pieces_per_row) return;
  long long row;
  int off;
  split_item(i, p.pieces_per_row, VEC, row, off);
  constexpr int NW = VEC >= 4 ? VEC / 4 : 1;
  uint32_t x[NW], y[NW];
  load_bytes<VEC>(p.x + row * p.x_stride + off, x);
#pragma unroll
  for (int w = 0; w < NW; w++) {
    if constexpr (LUT) {
      if constexpr (VEC == 1) {
        y[w] = lut[x[w]];
      } else {
        y[w] = (uint32_t) lut[x[w] & 0xFF] | ((uint32_t) lut[(x[w] >> 8) & 0xFF] << 8) | ((uint32_t) lut[(x[w] >> 16) & 0xFF] << 16) |
            ((uint32_t) lut[x[w] >> 24] << 24);
      }
    } else {
      if constexpr (VEC == 1) {
        y[w] = min(max(x[w], p.lo), p.hi);
      } else {
        y[w] = __vminu4(__vmaxu4(x[w], p.lo * 0x01010101u), p.hi * 0x01010101u);
      }
    }
  }
  store_bytes<VEC>(p.y + row * p.y_stride + off, y);
}

// ---- channel shuffle: y[row][c * groups + g] = x[row][g * group_channels + c]  (src/x8zip) ---------------------------
__global__ void __launch_bounds__(kThreads) q8_shuffle_kernel(const __grid_constant__ ShuffleParams p) {
  const long long i = (long long) blockIdx.x * kThreads + threadIdx.x;
  const int channels = p.groups * p.group_channels;
  if (i >= p.rows * channels) return;
  const long long row = i / channels;
  const int o = (int) (i - row * channels);
  const int c = o / p.groups, g = o - c * p.groups;
  p.y[row * p.y_stride + o] = p.x[row * p.x_stride + (long long) g * p.group_channels + c];
}

// ---- softargmax: one warp per row ------------------------------------------------------------------------------------
// operator-run.c:625-637: t' = t + (255 - max(x));  u8lut32norm: y = min(255, ((t'[x] << 8) + sum/2) / sum), sum = sum t'[x]
__global__ void __launch_bounds__(kThreads) q8_softargmax_kernel(const __grid_constant__ SoftargmaxParams p) {
  const long long row = (long long) blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int lane = threadIdx.x & 31;
  const uint8_t* x = p.x + row * p.x_stride;
  uint8_t* y = p.y + row * p.y_stride;
  uint32_t mx = 0;
  for (int c = lane; c < p.channels; c += 32) mx = max(mx, (uint32_t) x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const uint32_t* t = p.table + (mx ^ 255u);
  uint32_t sum = 0;
  for (int c = lane; c < p.channels; c += 32) sum += __ldg(t + x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const uint32_t rounding = sum >> 1;
  for (int c = lane; c < p.channels; c += 32) {
    const uint32_t q = ((__ldg(t + x[c]) << 8) + rounding) / sum;
    y[c] = (uint8_t) (q > 255u ? 255u : q);
  }
}

// ---- global average pooling: [batch][width][channels] -> [batch][channels] -------------------------------------------
template <int CV>
__global__ void __launch_bounds__(kThreads) q8_gavgpool_kernel(const __grid_constant__ GavgParams p) {
  const long long i = (long long) blockIdx.x * kThreads + threadIdx.x;
  const int cgroups = p.channels / CV;
  if (i >= p.batch * cgroups) return;
  const long long n = i / cgroups;
  const int c0 = (int) (i - n * cgroups) * CV;
  const uint8_t* x = p.x + n * p.width * p.x_stride + c0;
  int32_t acc[CV];
#pragma unroll
  for (int k = 0; k < CV; k++) acc[k] = p.bias;
  long long w = 0;
  if constexpr (CV == 4) {
    // seven pixels per step, loads first: seven independent requests in flight per thread (a 7x7 map is seven steps)
    for (; w + 7 <= p.width; w += 7) {
      uint32_t v[7];
#pragma unroll
      for (int i = 0; i < 7; i++) v[i] = *reinterpret_cast<const uint32_t*>(x + (w + i) * p.x_stride);
#pragma unroll
      for (int i = 0; i < 7; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[k] = __dp4a(v[i], (uint32_t) (1u << (8 * k)), (uint32_t) acc[k]);  // + byte k
    }
  }
  for (; w < p.width; w++) {
    if constexpr (CV == 4) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(x + w * p.x_stride);
#pragma unroll
      for (int k = 0; k < 4; k++) acc[k] = __dp4a(v, (uint32_t) (1u << (8 * k)), (uint32_t) acc[k]);
    } else {
      acc[0] += x[w * p.x_stride];
    }
  }
  uint8_t* y = p.y + n * p.y_stride + c0;
  if constexpr (CV == 4) {
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) o |= avgpool_quantize(acc[k], p.q) << (8 * k);
    *reinterpret_cast<uint32_t*>(y) = o;
  } else {
    y[0] = (uint8_t) avgpool_quantize(acc[0], p.q);
  }
}

// ---- 2-D pooling (NHWC): thread = (image, output pixel, CV channels) -------------------------------------------------
template <int CV, bool MAX>
__global__ void __launch_bounds__(kThreads) q8_pool2d_kernel(const __grid_constant__ PoolParams p) {
  const long long i = (long long) blockIdx.x * kThreads + threadIdx.x;
  const int cgroups = p.channels / CV;
  if (i >= p.batch * p.out_h * p.out_w * cgroups) return;
  long long r = i / cgroups;
  const int c0 = (int) (i - r * cgroups) * CV;
  const int ox = (int) (r % p.out_w);
  r /= p.out_w;
  const int oy = (int) (r % p.out_h);
  const long long n = r / p.out_h;
  int32_t acc[CV];
#pragma unroll
  for (int k = 0; k < CV; k++) acc[k] = MAX ? 0 : p.bias;
  for (int ky = 0; ky < p.kh; ky++) {
    int iy = oy * p.stride_h + ky * p.dil_h - p.pad_top;
    if constexpr (MAX) {
      // src/indirection.c:218-224: a tap outside the image reads the nearest edge pixel (doz(), then min with size - 1)
      iy = iy < 0 ? 0 : (iy > p.in_h - 1 ? p.in_h - 1 : iy);
    } else if ((unsigned) iy >= (unsigned) p.in_h) {
      continue;  // src/average-pooling.c:143-150: a padded tap reads the zero buffer = izp, i.e. contributes izp - izp = 0
    }
    for (int kx = 0; kx < p.kw; kx++) {
      int ix = ox * p.stride_w + kx * p.dil_w - p.pad_left;
      if constexpr (MAX) {
        ix = ix < 0 ? 0 : (ix > p.in_w - 1 ? p.in_w - 1 : ix);
      } else if ((unsigned) ix >= (unsigned) p.in_w) {
        continue;
      }
      const uint8_t* x = p.x + ((n * p.in_h + iy) * p.in_w + ix) * p.x_stride + c0;
      if constexpr (CV == 4) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(x);
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int32_t b = (int32_t) ((v >> (8 * k)) & 0xFF);
          acc[k] = MAX ? max(acc[k], b) : acc[k] + b - p.izp;
        }
      } else {
        const int32_t b = x[0];
        acc[0] = MAX ? max(acc[0], b) : acc[0] + b - p.izp;
      }
    }
  }
  uint8_t* y = p.y + ((n * p.out_h + oy) * p.out_w + ox) * p.y_stride + c0;
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < CV; k++) {
    uint32_t v;
    if constexpr (MAX) {
      v = (uint32_t) min(max(acc[k], p.lo), p.hi);
    } else {
      v = avgpool_quantize(acc[k], p.q);
    }
    o |= v << (8 * k);
  }
  if constexpr (CV == 4) {
    *reinterpret_cast<uint32_t*>(y) = o;
  } else {
    y[0] = (uint8_t) o;
  }
}

}  // namespace

cudaError_t launch_q8_add(const AddParams& p, int vec, cudaStream_t stream) {
  const long long work = p.rows * p.pieces_per_row;
  if (work == 0) return cudaSuccess;
  const bool fast = p.shift <= 22;
  if (vec == 16) {
    if (fast) q8_add_kernel<16, true><<<blocks_for(work), kThreads, 0, stream>>>(p);
    else q8_add_kernel<16, false><<<blocks_for(work), kThreads, 0, stream>>>(p);
  } else if (vec == 4) {
    if (fast) q8_add_kernel<4, true><<<blocks_for(work), kThreads, 0, stream>>>(p);
    else q8_add_kernel<4, false><<<blocks_for(work), kThreads, 0, stream>>>(p);
  } else {
    q8_add_kernel<1, false><<<blocks_for(work), kThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_q8_map(const MapParams& p, int vec, bool lut, cudaStream_t stream) {
  const long long work = p.rows * p.pieces_per_row;
  if (work == 0) return cudaSuccess;
  const unsigned g = blocks_for(work);
  if (lut) {
    if (vec == 16) q8_map_kernel<16, true><<<g, kThreads, 0, stream>>>(p);
    else if (vec == 4) q8_map_kernel<4, true><<<g, kThreads, 0, stream>>>(p);
    else q8_map_kernel<1, true><<<g, kThreads, 0, stream>>>(p);
  } else {
    if (vec == 16) q8_map_kernel<16, false><<<g, kThreads, 0, stream>>>(p);
    else if (vec == 4) q8_map_kernel<4, false><<<g, kThreads, 0, stream>>>(p);
    else q8_map_kernel<1, false><<<g, kThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_q8_shuffle(const ShuffleParams& p, cudaStream_t stream) {
  const long long work = p.rows * p.groups * p.group_channels;
  if (work == 0) return cudaSuccess;
  q8_shuffle_kernel<<<blocks_for(work), kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_q8_softargmax(const SoftargmaxParams& p, cudaStream_t stream) {
  if (p.rows == 0) return cudaSuccess;
  q8_softargmax_kernel<<<(unsigned) ((p.rows + kThreads / 32 - 1) / (kThreads / 32)), kThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_q8_gavgpool(const GavgParams& p, int cv, cudaStream_t stream) {
  const long long work = p.batch * (p.channels / cv);
  if (work == 0) return cudaSuccess;
  if (cv == 4) {
    q8_gavgpool_kernel<4><<<blocks_for(work), kThreads, 0, stream>>>(p);
  } else {
    q8_gavgpool_kernel<1><<<blocks_for(work), kThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

cudaError_t launch_q8_pool2d(const PoolParams& p, int cv, bool is_max, cudaStream_t stream) {
  const long long work = p.batch * p.out_h * p.out_w * (p.channels / cv);
  if (work == 0) return cudaSuccess;
  const unsigned g = blocks_for(work);
  if (is_max) {
    if (cv == 4) q8_pool2d_kernel<4, true><<<g, kThreads, 0, stream>>>(p);
    else q8_pool2d_kernel<1, true><<<g, kThreads, 0, stream>>>(p);
  } else {
    if (cv == 4) q8_pool2d_kernel<4, false><<<g, kThreads, 0, stream>>>(p);
    else q8_pool2d_kernel<1, false><<<g, kThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace q8
