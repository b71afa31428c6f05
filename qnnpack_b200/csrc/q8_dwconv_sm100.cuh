// Parameters of the CUDA-core kernels: 3x3 depthwise, generic direct convolution, stand-alone requantizer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "requant_math.h"

namespace q8 {

struct DwParams {
  const uint8_t* in;
  uint8_t* out;
  const int32_t* w32;   // [9][c_pad]  (w - kzp), tap index = ky*3 + kx
  const int32_t* bias;  // [c_pad] folded bias (reference pack.h:146-159)
  long long in_stride, out_stride;
  long long total_threads;
  int batch, channels, c_pad;
  int in_h, in_w, out_h, out_w;
  int stride_h, stride_w, dil_h, dil_w, pad_top, pad_left;
  int cgroups, xstrips;
  int izp;
  int rq_mode;
  Q8Requant rq;
};

// streaming dp4a kernel (q8_dwconv_stream_sm100.cu)
struct DwStreamParams {
  const uint8_t* in;
  uint8_t* out;
  const uint32_t* wa;   // [3][c_pad]: per channel and kernel row the three taps packed as bytes (operand A)
  const uint32_t* wb;   // [3][c_pad]: operand B when (w - kzp) does not fit 8 bits (wmode 2)
  const int32_t* bias;  // [c_pad] folded bias
  long long in_stride, out_stride;
  long long total_threads;
  int batch, channels, c_pad;
  int in_h, in_w, out_h, out_w;
  int stride, pad_top, pad_left;
  int cgroups, xstrips, ychunks, tyc;
  int wmode;            // 0: one s8 operand, 1: one u8 operand (kzp == 0), 2: two s8 operands (w - kzp = A + B)
  int izp;
  int rq_mode;
  int shift_mul;        // see requant_dev.cuh
  Q8Requant rq;
};

struct DirectParams {
  const uint8_t* in;
  uint8_t* out;
  const uint8_t* w;     // original layout [G][GOC][KH][KW][GIC]
  const int32_t* bias;  // [G*GOC] folded bias
  long long in_stride, out_stride;
  long long total;      // batch*out_h*out_w*groups*goc
  int groups, gic, goc;
  int in_h, in_w, out_h, out_w, kh, kw;
  int stride_h, stride_w, dil_h, dil_w, pad_top, pad_left;
  int izp, kzp;
  Q8Requant rq;
};

cudaError_t launch_q8_dwconv3x3(DwParams p, int cv, cudaStream_t stream);
cudaError_t launch_q8_dwconv3x3_stream(DwStreamParams p, cudaStream_t stream);
cudaError_t launch_q8_direct_conv(const DirectParams& p, cudaStream_t stream);
cudaError_t launch_q8_requantize(const int32_t* in, uint8_t* out, long long n, const Q8Requant& rq, cudaStream_t stream);

}  // namespace q8
