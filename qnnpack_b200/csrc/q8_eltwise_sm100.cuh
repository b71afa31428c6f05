// Parameters of the element-wise / pooling kernels (q8_eltwise_sm100.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace q8 {

// rows x pieces_per_row work items of `vec` bytes; row r of tensor t starts at t + r * t_stride
struct AddParams {
  const uint8_t* a;
  const uint8_t* b;
  uint8_t* y;
  long long rows, a_stride, b_stride, y_stride;
  int pieces_per_row;
  // reference qnnp_compute_add_quantization_params, scalar member (src/qnnpack/requantization.h:327-414)
  int32_t zero_point_product;
  uint32_t a_multiplier, b_multiplier;
  int32_t shift, remainder_mask, remainder_threshold, y_zero_point, y_min, y_max;
};

struct MapParams {  // clamp (lo/hi) or 256-entry lookup table (sigmoid, leaky ReLU)
  const uint8_t* x;
  uint8_t* y;
  const uint8_t* lut;  // device, 256 bytes (LUT variant)
  long long rows, x_stride, y_stride;
  int pieces_per_row;
  uint32_t lo, hi;
};

struct ShuffleParams {
  const uint8_t* x;
  uint8_t* y;
  long long rows, x_stride, y_stride;
  int groups, group_channels;
};

struct SoftargmaxParams {
  const uint8_t* x;
  uint8_t* y;
  const uint32_t* table;  // device, 256 + 255 entries: the row's table is table + (255 - max)  (zero-extended at the end)
  long long rows, x_stride, y_stride;
  int channels;
};

// reference qnnp_compute_avgpool_quantization_params, scalar member (requantization.h:200-298)
struct AvgQuant {
  int32_t multiplier;
  int64_t rounding;
  uint32_t right_shift;
  int32_t min_less_zp, max_less_zp, zero_point;
};

struct GavgParams {
  const uint8_t* x;
  uint8_t* y;
  long long batch, width, x_stride, y_stride;
  int channels;
  int32_t bias;  // -width * input_zero_point
  AvgQuant q;
};

struct PoolParams {
  const uint8_t* x;
  uint8_t* y;
  long long batch, x_stride, y_stride;
  int channels, in_h, in_w, out_h, out_w, kh, kw, stride_h, stride_w, dil_h, dil_w, pad_top, pad_left;
  int32_t izp, bias;  // average pooling: sum of (x - izp) over the taps inside the image; bias = 0
  int32_t lo, hi;     // max pooling clamp
  AvgQuant q;
};

cudaError_t launch_q8_add(const AddParams& p, int vec, cudaStream_t stream);
cudaError_t launch_q8_map(const MapParams& p, int vec, bool lut, cudaStream_t stream);
cudaError_t launch_q8_shuffle(const ShuffleParams& p, cudaStream_t stream);
cudaError_t launch_q8_softargmax(const SoftargmaxParams& p, cudaStream_t stream);
cudaError_t launch_q8_gavgpool(const GavgParams& p, int cv, cudaStream_t stream);
cudaError_t launch_q8_pool2d(const PoolParams& p, int cv, bool is_max, cudaStream_t stream);

}  // namespace q8
