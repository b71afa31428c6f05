"""ctypes signatures for the qnnpack.h C ABI (include/qnnpack.h in this repo; reference
include/qnnpack.h:24-332).  The same binder works for any shared object that exports that ABI,
which is the point of a drop-in: tests bind the compiled reference with it, the product binds
libqnnpack.so with it."""
from __future__ import annotations

import ctypes as C

STATUS_NAMES = {
    0: "success",
    1: "uninitialized",
    2: "invalid_parameter",
    3: "unsupported_parameter",
    4: "unsupported_hardware",
    5: "out_of_memory",
}

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
op_t = C.c_void_p
size_t = C.c_size_t
u32 = C.c_uint32
u8 = C.c_uint8
f32 = C.c_float


# Every symbol include/qnnpack.h declares; tests check that the product library exports all of them.
ALL_QNNPACK_H_SYMBOLS = [
    "qnnp_initialize", "qnnp_deinitialize",
    "qnnp_create_convolution2d_nhwc_q8", "qnnp_setup_convolution2d_nhwc_q8",
    "qnnp_create_deconvolution2d_nhwc_q8", "qnnp_setup_deconvolution2d_nhwc_q8",
    "qnnp_create_fully_connected_nc_q8", "qnnp_setup_fully_connected_nc_q8",
    "qnnp_create_global_average_pooling_nwc_q8", "qnnp_setup_global_average_pooling_nwc_q8",
    "qnnp_create_average_pooling2d_nhwc_q8", "qnnp_setup_average_pooling2d_nhwc_q8",
    "qnnp_create_max_pooling2d_nhwc_u8", "qnnp_setup_max_pooling2d_nhwc_u8",
    "qnnp_create_channel_shuffle_nc_x8", "qnnp_setup_channel_shuffle_nc_x8",
    "qnnp_create_add_nc_q8", "qnnp_setup_add_nc_q8",
    "qnnp_create_clamp_nc_u8", "qnnp_setup_clamp_nc_u8",
    "qnnp_create_sigmoid_nc_q8", "qnnp_setup_sigmoid_nc_q8",
    "qnnp_create_leaky_relu_nc_q8", "qnnp_setup_leaky_relu_nc_q8",
    "qnnp_create_softargmax_nc_q8", "qnnp_setup_softargmax_nc_q8",
    "qnnp_run_operator", "qnnp_delete_operator",
]


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach argtypes/restype for the entry points of the q8 hot path."""
    lib.qnnp_initialize.argtypes = []
    lib.qnnp_initialize.restype = C.c_int
    lib.qnnp_deinitialize.argtypes = []
    lib.qnnp_deinitialize.restype = C.c_int

    # include/qnnpack.h:40-65
    lib.qnnp_create_convolution2d_nhwc_q8.argtypes = [
        u32, u32, u32, u32,  # padding top, right, bottom, left
        u32, u32,            # kernel h, w
        u32, u32,            # subsampling h, w
        u32, u32,            # dilation h, w
        u32, size_t, size_t,  # groups, group_input_channels, group_output_channels
        u8, f32, u8, f32,    # input zp/scale, kernel zp/scale
        C.c_void_p, C.c_void_p,  # kernel, bias
        u8, f32, u8, u8,     # output zp/scale, min, max
        u32, C.POINTER(op_t),
    ]
    lib.qnnp_create_convolution2d_nhwc_q8.restype = C.c_int
    # include/qnnpack.h:67-76
    lib.qnnp_setup_convolution2d_nhwc_q8.argtypes = [
        op_t, size_t, size_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t, C.c_void_p,
    ]
    lib.qnnp_setup_convolution2d_nhwc_q8.restype = C.c_int
    # include/qnnpack.h:118-132
    lib.qnnp_create_fully_connected_nc_q8.argtypes = [
        size_t, size_t, u8, f32, u8, f32, C.c_void_p, C.c_void_p, u8, f32, u8, u8, u32, C.POINTER(op_t),
    ]
    lib.qnnp_create_fully_connected_nc_q8.restype = C.c_int
    # include/qnnpack.h:134-140
    lib.qnnp_setup_fully_connected_nc_q8.argtypes = [op_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t]
    lib.qnnp_setup_fully_connected_nc_q8.restype = C.c_int
    # include/qnnpack.h:327-332
    # the operators beside the convolution path (include/qnnpack.h:78-116, 142-325)
    nc_setup = [op_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t]
    lib.qnnp_create_deconvolution2d_nhwc_q8.argtypes = [u32] * 13 + [size_t, size_t, u8, f32, u8, f32, C.c_void_p, C.c_void_p,
                                                                     u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_deconvolution2d_nhwc_q8.argtypes = [op_t, size_t, size_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t,
                                                       C.c_void_p]
    lib.qnnp_create_add_nc_q8.argtypes = [size_t, u8, f32, u8, f32, u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_add_nc_q8.argtypes = [op_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t, C.c_void_p, size_t]
    lib.qnnp_create_global_average_pooling_nwc_q8.argtypes = [size_t, u8, f32, u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_global_average_pooling_nwc_q8.argtypes = [op_t, size_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t]
    lib.qnnp_create_average_pooling2d_nhwc_q8.argtypes = [u32] * 8 + [size_t, u8, f32, u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_average_pooling2d_nhwc_q8.argtypes = [op_t, size_t, size_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t,
                                                         C.c_void_p]
    lib.qnnp_create_max_pooling2d_nhwc_u8.argtypes = [u32] * 10 + [size_t, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_max_pooling2d_nhwc_u8.argtypes = [op_t, size_t, size_t, size_t, C.c_void_p, size_t, C.c_void_p, size_t,
                                                     C.c_void_p]
    lib.qnnp_create_channel_shuffle_nc_x8.argtypes = [size_t, size_t, u32, C.POINTER(op_t)]
    lib.qnnp_setup_channel_shuffle_nc_x8.argtypes = nc_setup
    lib.qnnp_create_clamp_nc_u8.argtypes = [size_t, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_clamp_nc_u8.argtypes = nc_setup
    lib.qnnp_create_sigmoid_nc_q8.argtypes = [size_t, u8, f32, u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_sigmoid_nc_q8.argtypes = nc_setup
    lib.qnnp_create_leaky_relu_nc_q8.argtypes = [size_t, f32, u8, f32, u8, f32, u8, u8, u32, C.POINTER(op_t)]
    lib.qnnp_setup_leaky_relu_nc_q8.argtypes = nc_setup
    lib.qnnp_create_softargmax_nc_q8.argtypes = [size_t, f32, u8, f32, u32, C.POINTER(op_t)]
    lib.qnnp_setup_softargmax_nc_q8.argtypes = nc_setup
    for name in ALL_QNNPACK_H_SYMBOLS:
        getattr(lib, name).restype = C.c_int
    lib.qnnp_run_operator.argtypes = [op_t, C.c_void_p]
    lib.qnnp_run_operator.restype = C.c_int
    lib.qnnp_delete_operator.argtypes = [op_t]
    lib.qnnp_delete_operator.restype = C.c_int
    return lib


