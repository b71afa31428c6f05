/*
 * qnnpack.h — C ABI of the B200-native q8 inference library (drop-in for pytorch/QNNPACK's public header).
 *
 * Every declaration below has the same name, argument order, argument types and status codes as the
 * reference's include/qnnpack.h (cited per entry point as "ref :first-last"), so code written against
 * QNNPACK links against libqnnpack.so unchanged.  What differs is behind the ABI:
 *
 *   - the q8gemm / q8conv / q8dwconv hot path (convolution2d_nhwc_q8, fully_connected_nc_q8,
 *     run_operator) executes hand-written sm_100a kernels; there is NO CPU path — on a machine
 *     without a compute-capability-10.x GPU qnnp_initialize() returns qnnp_status_unsupported_hardware;
 *   - `input` / `output` given to qnnp_setup_* may be host pointers (copied to and from the device
 *     inside qnnp_run_operator, which stays synchronous like the reference) or device pointers
 *     (zero-copy; see qnnpack_cuda.h for streams and asynchronous runs);
 *   - `pthreadpool_t` arguments are accepted and ignored (NULL is legal, as in the reference's tests);
 *   - operators outside the hot path (the 20 entry points marked "not on the q8 hot path") link but
 *     return qnnp_status_unsupported_parameter.
 */
#pragma once

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

/* The reference includes <pthreadpool.h> only for this opaque handle type (ref :15). */
#ifndef QNNP_HAVE_PTHREADPOOL_H
typedef struct pthreadpool* pthreadpool_t;
#else
#include <pthreadpool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ref :24-32 */
enum qnnp_status {
  qnnp_status_success = 0,
  qnnp_status_uninitialized = 1,
  qnnp_status_invalid_parameter = 2,
  qnnp_status_unsupported_parameter = 3,
  qnnp_status_unsupported_hardware = 4,
  qnnp_status_out_of_memory = 5,
};

/* ref :38 — opaque */
typedef struct qnnp_operator* qnnp_operator_t;

/* ref :34-36.  Idempotent and thread-safe (src/init.c:244-258); selects the CUDA device
 * (QNNP_CUDA_DEVICE or the current device) and requires compute capability 10.x. */
enum qnnp_status qnnp_initialize(void);
enum qnnp_status qnnp_deinitialize(void);

/* ---- q8 hot path ------------------------------------------------------------------------------ */

/* ref :40-65 (src/convolution.c:39-378).  kernel: [groups][group_output_channels][kh][kw][group_input_channels],
 * bias: [groups*group_output_channels]; both are packed and copied to the device — the caller may free them. */
enum qnnp_status qnnp_create_convolution2d_nhwc_q8(
    uint32_t input_padding_top, uint32_t input_padding_right, uint32_t input_padding_bottom, uint32_t input_padding_left,
    uint32_t kernel_height, uint32_t kernel_width, uint32_t subsampling_height, uint32_t subsampling_width,
    uint32_t dilation_height, uint32_t dilation_width, uint32_t groups, size_t group_input_channels,
    size_t group_output_channels, uint8_t input_zero_point, float input_scale, uint8_t kernel_zero_point,
    float kernel_scale, const uint8_t* kernel, const int32_t* bias, uint8_t output_zero_point, float output_scale,
    uint8_t output_min, uint8_t output_max, uint32_t flags, qnnp_operator_t* convolution);

/* ref :67-76 (src/convolution.c:380-492).  Strides are in bytes per pixel, >= channels. */
enum qnnp_status qnnp_setup_convolution2d_nhwc_q8(
    qnnp_operator_t convolution, size_t batch_size, size_t input_height, size_t input_width, const uint8_t* input,
    size_t input_stride, uint8_t* output, size_t output_stride, pthreadpool_t threadpool);

/* ref :118-132 (src/fully-connected.c:25-129).  kernel: [output_channels][input_channels]. */
enum qnnp_status qnnp_create_fully_connected_nc_q8(
    size_t input_channels, size_t output_channels, uint8_t input_zero_point, float input_scale,
    uint8_t kernel_zero_point, float kernel_scale, const uint8_t* kernel, const int32_t* bias,
    uint8_t output_zero_point, float output_scale, uint8_t output_min, uint8_t output_max, uint32_t flags,
    qnnp_operator_t* fully_connected);

/* ref :134-140 (src/fully-connected.c:131-161) */
enum qnnp_status qnnp_setup_fully_connected_nc_q8(
    qnnp_operator_t fully_connected, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

/* ref :327-329 (src/operator-run.c:639-1153).  Synchronous: the output is complete on return. */
enum qnnp_status qnnp_run_operator(qnnp_operator_t op, pthreadpool_t threadpool);

/* ref :331-332 (src/operator-delete.c:15-28).  NULL -> qnnp_status_invalid_parameter. */
enum qnnp_status qnnp_delete_operator(qnnp_operator_t op);

/* ---- not on the q8 hot path: link-compatible, return qnnp_status_unsupported_parameter ---------- */

/* ref :78-116 */
enum qnnp_status qnnp_create_deconvolution2d_nhwc_q8(
    uint32_t input_padding_top, uint32_t input_padding_right, uint32_t input_padding_bottom, uint32_t input_padding_left,
    uint32_t adjustment_height, uint32_t adjustment_width, uint32_t kernel_height, uint32_t kernel_width,
    uint32_t stride_height, uint32_t stride_width, uint32_t dilation_height, uint32_t dilation_width, uint32_t groups,
    size_t group_input_channels, size_t group_output_channels, uint8_t input_zero_point, float input_scale,
    uint8_t kernel_zero_point, float kernel_scale, const uint8_t* kernel, const int32_t* bias,
    uint8_t output_zero_point, float output_scale, uint8_t output_min, uint8_t output_max, uint32_t flags,
    qnnp_operator_t* deconvolution);
enum qnnp_status qnnp_setup_deconvolution2d_nhwc_q8(
    qnnp_operator_t deconvolution, size_t batch_size, size_t input_height, size_t input_width, const uint8_t* input,
    size_t input_stride, uint8_t* output, size_t output_stride, pthreadpool_t threadpool);

/* ref :142-160 */
enum qnnp_status qnnp_create_global_average_pooling_nwc_q8(
    size_t channels, uint8_t input_zero_point, float input_scale, uint8_t output_zero_point, float output_scale,
    uint8_t output_min, uint8_t output_max, uint32_t flags, qnnp_operator_t* global_average_pooling);
enum qnnp_status qnnp_setup_global_average_pooling_nwc_q8(
    qnnp_operator_t global_average_pooling, size_t batch_size, size_t width, const uint8_t* input, size_t input_stride,
    uint8_t* output, size_t output_stride);

/* ref :162-190 */
enum qnnp_status qnnp_create_average_pooling2d_nhwc_q8(
    uint32_t input_padding_top, uint32_t input_padding_right, uint32_t input_padding_bottom, uint32_t input_padding_left,
    uint32_t pooling_height, uint32_t pooling_width, uint32_t stride_height, uint32_t stride_width, size_t channels,
    uint8_t input_zero_point, float input_scale, uint8_t output_zero_point, float output_scale, uint8_t output_min,
    uint8_t output_max, uint32_t flags, qnnp_operator_t* average_pooling);
enum qnnp_status qnnp_setup_average_pooling2d_nhwc_q8(
    qnnp_operator_t average_pooling, size_t batch_size, size_t input_height, size_t input_width, const uint8_t* input,
    size_t input_stride, uint8_t* output, size_t output_stride, pthreadpool_t threadpool);

/* ref :192-218 */
enum qnnp_status qnnp_create_max_pooling2d_nhwc_u8(
    uint32_t input_padding_top, uint32_t input_padding_right, uint32_t input_padding_bottom, uint32_t input_padding_left,
    uint32_t pooling_height, uint32_t pooling_width, uint32_t stride_height, uint32_t stride_width,
    uint32_t dilation_height, uint32_t dilation_width, size_t channels, uint8_t output_min, uint8_t output_max,
    uint32_t flags, qnnp_operator_t* max_pooling);
enum qnnp_status qnnp_setup_max_pooling2d_nhwc_u8(
    qnnp_operator_t max_pooling, size_t batch_size, size_t input_height, size_t input_width, const uint8_t* input,
    size_t input_stride, uint8_t* output, size_t output_stride, pthreadpool_t threadpool);

/* ref :220-232 */
enum qnnp_status qnnp_create_channel_shuffle_nc_x8(
    size_t groups, size_t group_channels, uint32_t flags, qnnp_operator_t* channel_shuffle);
enum qnnp_status qnnp_setup_channel_shuffle_nc_x8(
    qnnp_operator_t channel_shuffle, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

/* ref :234-255 */
enum qnnp_status qnnp_create_add_nc_q8(
    size_t channels, uint8_t a_zero_point, float a_scale, uint8_t b_zero_point, float b_scale, uint8_t sum_zero_point,
    float sum_scale, uint8_t sum_min, uint8_t sum_max, uint32_t flags, qnnp_operator_t* add);
enum qnnp_status qnnp_setup_add_nc_q8(
    qnnp_operator_t add, size_t batch_size, const uint8_t* a, size_t a_stride, const uint8_t* b, size_t b_stride,
    uint8_t* sum, size_t sum_stride);

/* ref :257-270 */
enum qnnp_status qnnp_create_clamp_nc_u8(
    size_t channels, uint8_t output_min, uint8_t output_max, uint32_t flags, qnnp_operator_t* clamp);
enum qnnp_status qnnp_setup_clamp_nc_u8(
    qnnp_operator_t clamp, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

/* ref :272-289 */
enum qnnp_status qnnp_create_sigmoid_nc_q8(
    size_t channels, uint8_t input_zero_point, float input_scale, uint8_t output_zero_point, float output_scale,
    uint8_t output_min, uint8_t output_max, uint32_t flags, qnnp_operator_t* sigmoid);
enum qnnp_status qnnp_setup_sigmoid_nc_q8(
    qnnp_operator_t sigmoid, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

/* ref :291-309 */
enum qnnp_status qnnp_create_leaky_relu_nc_q8(
    size_t channels, float negative_slope, uint8_t input_zero_point, float input_scale, uint8_t output_zero_point,
    float output_scale, uint8_t output_min, uint8_t output_max, uint32_t flags, qnnp_operator_t* leaky_relu);
enum qnnp_status qnnp_setup_leaky_relu_nc_q8(
    qnnp_operator_t leaky_relu, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

/* ref :311-325 */
enum qnnp_status qnnp_create_softargmax_nc_q8(
    size_t channels, float input_scale, uint8_t output_zero_point, float output_scale, uint32_t flags,
    qnnp_operator_t* softargmax);
enum qnnp_status qnnp_setup_softargmax_nc_q8(
    qnnp_operator_t softargmax, size_t batch_size, const uint8_t* input, size_t input_stride, uint8_t* output,
    size_t output_stride);

#ifdef __cplusplus
} /* extern "C" */
#endif
