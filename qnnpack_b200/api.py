"""Python mirror of the qnnpack.h operator interface over a shared object exporting that C ABI.

``QnnpackLibrary`` drives  create -> setup -> run -> delete  exactly as the reference's operator
testers do (test/convolution-operator-tester.h:416-447, test/fully-connected-operator-tester.h), with
NumPy arrays (host pointers: the library stages the copies) or raw device pointers (zero-copy).
``load()`` returns the product library (qnnpack_b200/lib/libqnnpack.so, sm_100a kernels); it raises
if the extension has not been built or no B200-class GPU is present — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
# QNNP_LIB_PATH selects another build of the same library (A/B runs of compile-time variants)
PRODUCT_LIB = os.environ.get("QNNP_LIB_PATH") or os.path.join(HERE, "lib", "libqnnpack.so")


class QnnpackError(RuntimeError):
    def __init__(self, what: str, status: int):
        super().__init__(f"{what} -> qnnp_status_{_capi.STATUS_NAMES.get(status, status)}")
        self.status = status


def _ptr(a):
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw (device) address


def output_dim(in_dim: int, pad_a: int, pad_b: int, k: int, dil: int, stride: int) -> int:
    """reference src/convolution.c:29-37"""
    return (pad_a + in_dim + pad_b - ((k - 1) * dil + 1)) // stride + 1


class QnnpackLibrary:
    def __init__(self, path: str, threads: int = 0):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} is missing: build it first (python -m qnnpack_b200.build)")
        self.path = path
        self.lib = _capi.bind(C.CDLL(path))
        st = self.lib.qnnp_initialize()
        if st != 0:
            raise QnnpackError("qnnp_initialize", st)
        self.pool = None
        self.threads = 1
        if threads and threads > 1 and hasattr(self.lib, "pthreadpool_create"):
            self.lib.pthreadpool_create.argtypes = [C.c_size_t]
            self.lib.pthreadpool_create.restype = C.c_void_p
            self.lib.pthreadpool_destroy.argtypes = [C.c_void_p]
            self.pool = C.c_void_p(self.lib.pthreadpool_create(threads))
            self.threads = threads
        self.is_cuda = hasattr(self.lib, "qnnp_cuda_launch_count")
        if self.is_cuda:
            L = self.lib
            L.qnnp_cuda_launch_count.restype = C.c_ulonglong
            L.qnnp_cuda_debug_dw_umma_launch_count.restype = C.c_ulonglong
            L.qnnp_cuda_set_stream.argtypes = [C.c_void_p]
            L.qnnp_cuda_get_stream.restype = C.c_void_p
            L.qnnp_cuda_run_operator_async.argtypes = [C.c_void_p]
            L.qnnp_cuda_operator_packed_weights.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
            L.qnnp_cuda_operator_packed_bias.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
            L.qnnp_cuda_requantize_q31.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8,
                                                   C.c_void_p]
            L.qnnp_cuda_debug_set_accumulator_dump.argtypes = [C.c_void_p]
            L.qnnp_cuda_debug_set_accumulator_dump.restype = None
            L.qnnp_cuda_operator_kernel_name.argtypes = [C.c_void_p]
            L.qnnp_cuda_operator_kernel_name.restype = C.c_char_p

    def close(self):
        if self.pool is not None:
            self.lib.pthreadpool_destroy(self.pool)
            self.pool = None

    # -- operator objects -------------------------------------------------------------------
    def create_convolution(self, kernel, bias, *, pad=(0, 0, 0, 0), ksize=(1, 1), stride=(1, 1), dilation=(1, 1),
                           groups=1, gic, goc, izp, input_scale, kzp, kernel_scale, ozp, output_scale,
                           qmin=0, qmax=255):
        """pad = (top, right, bottom, left) as in qnnp_create_convolution2d_nhwc_q8. Returns (status, op)."""
        kernel = np.ascontiguousarray(kernel, dtype=np.uint8)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        op = _capi.op_t()
        st = self.lib.qnnp_create_convolution2d_nhwc_q8(
            pad[0], pad[1], pad[2], pad[3], ksize[0], ksize[1], stride[0], stride[1], dilation[0], dilation[1],
            groups, gic, goc, izp, float(np.float32(input_scale)), kzp, float(np.float32(kernel_scale)),
            _ptr(kernel), _ptr(bias), ozp, float(np.float32(output_scale)), qmin, qmax, 0, C.byref(op))
        return st, op

    def create_fully_connected(self, kernel, bias, *, izp, input_scale, kzp, kernel_scale, ozp, output_scale,
                               qmin=0, qmax=255):
        kernel = np.ascontiguousarray(kernel, dtype=np.uint8)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        oc, ic = kernel.shape
        op = _capi.op_t()
        st = self.lib.qnnp_create_fully_connected_nc_q8(
            ic, oc, izp, float(np.float32(input_scale)), kzp, float(np.float32(kernel_scale)), _ptr(kernel),
            _ptr(bias), ozp, float(np.float32(output_scale)), qmin, qmax, 0, C.byref(op))
        return st, op

    def setup_convolution(self, op, batch, in_h, in_w, x, in_stride, out, out_stride):
        """x / out: NumPy arrays (host) or integer device addresses."""
        return self.lib.qnnp_setup_convolution2d_nhwc_q8(op, batch, in_h, in_w, _ptr(x), in_stride, _ptr(out),
                                                         out_stride, self.pool)

    def setup_fully_connected(self, op, batch, x, in_stride, out, out_stride):
        return self.lib.qnnp_setup_fully_connected_nc_q8(op, batch, _ptr(x), in_stride, _ptr(out), out_stride)

    # -- the operators beside the convolution path: one generic driver ---------------------------------
    def create(self, name: str, *args):
        """qnnp_create_<name>(*args, flags=0, &op) -> (status, op)"""
        op = _capi.op_t()
        fn = getattr(self.lib, f"qnnp_create_{name}")
        cargs = [float(np.float32(a)) if isinstance(a, (float, np.floating)) else (_ptr(a) if isinstance(a, np.ndarray) else a)
                 for a in args]
        return fn(*cargs, 0, C.byref(op)), op

    def setup(self, name: str, op, *args, threadpool=False):
        """qnnp_setup_<name>(op, *args[, threadpool]); NumPy arrays become pointers, device addresses are passed as
        ctypes.c_void_p(address)."""
        fn = getattr(self.lib, f"qnnp_setup_{name}")
        cargs = [_ptr(a) if isinstance(a, np.ndarray) else a for a in args]
        if threadpool:
            cargs.append(self.pool)
        return fn(op, *cargs)

    def run(self, op):
        return self.lib.qnnp_run_operator(op, self.pool)

    def run_async(self, op):
        return self.lib.qnnp_cuda_run_operator_async(op)

    def delete(self, op):
        return self.lib.qnnp_delete_operator(op)

    def kernel_name(self, op) -> str:
        return self.lib.qnnp_cuda_operator_kernel_name(op).decode() if self.is_cuda else "reference"

    def launch_count(self) -> int:
        return int(self.lib.qnnp_cuda_launch_count()) if self.is_cuda else 0

    def dw_umma_launch_count(self) -> int:
        return int(self.lib.qnnp_cuda_debug_dw_umma_launch_count()) if self.is_cuda else 0

    def packed_weights(self, op):
        p, n = C.c_void_p(), C.c_size_t()
        st = self.lib.qnnp_cuda_operator_packed_weights(op, C.byref(p), C.byref(n))
        if st != 0:
            raise QnnpackError("qnnp_cuda_operator_packed_weights", st)
        return p.value, n.value

    def packed_bias(self, op):
        p, n = C.c_void_p(), C.c_size_t()
        st = self.lib.qnnp_cuda_operator_packed_bias(op, C.byref(p), C.byref(n))
        if st != 0:
            raise QnnpackError("qnnp_cuda_operator_packed_bias", st)
        return p.value, n.value

    def measure_int8_peak(self, iters: int = 2000, reps: int = 5):
        """-> (tera-ops/s, ms per launch): smem-resident tcgen05 kind::i8 loop on every SM (q8_peak_sm100.cu)."""
        tops, ms = C.c_double(), C.c_double()
        self.lib.qnnp_cuda_measure_int8_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        st = self.lib.qnnp_cuda_measure_int8_peak(iters, reps, C.byref(tops), C.byref(ms))
        if st != 0:
            raise QnnpackError("qnnp_cuda_measure_int8_peak", st)
        return tops.value, ms.value

    def set_stream(self, cuda_stream: int):
        st = self.lib.qnnp_cuda_set_stream(C.c_void_p(cuda_stream))
        if st != 0:
            raise QnnpackError("qnnp_cuda_set_stream", st)

    # -- one-shot helpers on host arrays ------------------------------------------------------
    def convolution(self, x, kernel, bias, *, out_stride=None, out_fill=0xA5, lead_in=16, **kw):
        """x: uint8 [N,H,W,in_stride] on the host; returns uint8 [N,OH,OW,out_stride] (bytes of a pixel
        beyond groups*goc keep ``out_fill``).  ``lead_in`` spare bytes precede the input because the
        reference's SSE2 tails read up to 7 bytes before a row (src/q8gemm/4x4c2-sse2.c:111-121)."""
        st, op = self.create_convolution(kernel, bias, **kw)
        if st != 0:
            raise QnnpackError("qnnp_create_convolution2d_nhwc_q8", st)
        try:
            pad, ksize = kw.get("pad", (0, 0, 0, 0)), kw.get("ksize", (1, 1))
            stride, dil = kw.get("stride", (1, 1)), kw.get("dilation", (1, 1))
            n, h, w, in_stride = x.shape
            oh = output_dim(h, pad[0], pad[2], ksize[0], dil[0], stride[0])
            ow = output_dim(w, pad[3], pad[1], ksize[1], dil[1], stride[1])
            out_stride = kw.get("groups", 1) * kw["goc"] if out_stride is None else out_stride
            buf = np.zeros(lead_in + x.size + 16, dtype=np.uint8)
            xin = buf[lead_in:lead_in + x.size].reshape(x.shape)
            xin[...] = x
            out = np.full((n, oh, ow, out_stride), out_fill, dtype=np.uint8)
            st = self.setup_convolution(op, n, h, w, xin, in_stride, out, out_stride)
            if st != 0:
                raise QnnpackError("qnnp_setup_convolution2d_nhwc_q8", st)
            st = self.run(op)
            if st != 0:
                raise QnnpackError("qnnp_run_operator", st)
            return out
        finally:
            self.delete(op)

    def fully_connected(self, x, kernel, bias, *, out_stride=None, out_fill=0xA5, lead_in=16, **kw):
        st, op = self.create_fully_connected(kernel, bias, **kw)
        if st != 0:
            raise QnnpackError("qnnp_create_fully_connected_nc_q8", st)
        try:
            b, in_stride = x.shape
            oc = kernel.shape[0]
            out_stride = oc if out_stride is None else out_stride
            buf = np.zeros(lead_in + x.size + 16, dtype=np.uint8)
            xin = buf[lead_in:lead_in + x.size].reshape(x.shape)
            xin[...] = x
            out = np.full((b, out_stride), out_fill, dtype=np.uint8)
            st = self.setup_fully_connected(op, b, xin, in_stride, out, out_stride)
            if st != 0:
                raise QnnpackError("qnnp_setup_fully_connected_nc_q8", st)
            st = self.run(op)
            if st != 0:
                raise QnnpackError("qnnp_run_operator", st)
            return out
        finally:
            self.delete(op)

    def requantize_q31(self, acc: np.ndarray, scale, zp, qmin, qmax) -> np.ndarray:
        """Device counterpart of qnnp_requantize_q31__scalar (src/requantization/q31-scalar.c:17)."""
        acc = np.ascontiguousarray(acc, dtype=np.int32)
        out = np.empty(acc.shape, dtype=np.uint8)
        st = self.lib.qnnp_cuda_requantize_q31(acc.size, _ptr(acc), float(np.float32(scale)), zp, qmin, qmax, _ptr(out))
        if st != 0:
            raise QnnpackError("qnnp_cuda_requantize_q31", st)
        return out


_product = None


def load() -> QnnpackLibrary:
    """The product: libqnnpack.so with the sm_100a kernels.  Raises when it is missing or no B200 is present."""
    global _product
    if _product is None:
        _product = QnnpackLibrary(PRODUCT_LIB)
    return _product
