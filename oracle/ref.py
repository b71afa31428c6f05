"""TEST INFRASTRUCTURE — driver for the UNMODIFIED reference compiled into oracle/_ref/.

``oracle/_ref/libqnnpack_ref.so`` is built by ``make -C oracle ref`` from the sources where they lie
under /root/reference (never copied into this repo).  It is the strongest oracle we have (the
reference's own SSE2 micro-kernels through its own operator API) and the CPU baseline that
``bench.py --impl reference`` times.  The library travels to the GPU box inside the repo snapshot;
/root/reference itself does not, so nothing here reads it at run time.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from qnnpack_b200 import _capi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libqnnpack_ref.so")


def available() -> bool:
    return os.path.exists(REF_LIB)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class QnnpackHost:
    """Drives any host-memory implementation of the qnnpack.h ABI (here: the compiled reference)
    through create -> setup -> run -> delete, exactly as test/convolution-operator-tester.h:416-447
    and test/fully-connected-operator-tester.h do."""

    def __init__(self, path: str = REF_LIB, threads: int = 0):
        self.lib = _capi.bind(C.CDLL(path))
        st = self.lib.qnnp_initialize()
        if st != 0:
            raise RuntimeError(f"qnnp_initialize -> {_capi.STATUS_NAMES.get(st, st)}")
        self.pool = None
        self.threads = 1
        if threads and threads > 1:
            self.lib.pthreadpool_create.argtypes = [C.c_size_t]
            self.lib.pthreadpool_create.restype = C.c_void_p
            self.lib.pthreadpool_destroy.argtypes = [C.c_void_p]
            self.pool = C.c_void_p(self.lib.pthreadpool_create(threads))
            self.threads = threads

    def close(self):
        if self.pool is not None:
            self.lib.pthreadpool_destroy(self.pool)
            self.pool = None

    # -- operator objects -------------------------------------------------------------------
    def create_convolution(self, kernel, bias, *, pad=(0, 0, 0, 0), ksize=(1, 1), stride=(1, 1), dilation=(1, 1),
                           groups=1, gic, goc, izp, input_scale, kzp, kernel_scale, ozp, output_scale,
                           qmin=0, qmax=255):
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

    def setup_convolution(self, op, x: np.ndarray, out: np.ndarray):
        n, h, w, in_stride = x.shape
        return self.lib.qnnp_setup_convolution2d_nhwc_q8(op, n, h, w, _ptr(x), in_stride, _ptr(out),
                                                         out.shape[-1], self.pool)

    def run(self, op):
        return self.lib.qnnp_run_operator(op, self.pool)

    def delete(self, op):
        return self.lib.qnnp_delete_operator(op)

    # -- one-shot helpers ---------------------------------------------------------------------
    def convolution(self, x, kernel, bias, *, out_stride=None, out_fill=0xA5, lead_in=16, **kw):
        """x: uint8 [N,H,W,in_stride]; returns uint8 [N,OH,OW,out_stride].  ``lead_in`` bytes are
        allocated before the input because the reference's SSE2 tails read up to 7 bytes before a
        row (src/q8gemm/4x4c2-sse2.c:111-121; test/gemm-microkernel-tester.h:187,195)."""
        from .q8_oracle import output_dim

        st, op = self.create_convolution(kernel, bias, **kw)
        if st != 0:
            raise RuntimeError(f"create -> {_capi.STATUS_NAMES.get(st, st)}")
        try:
            pad, ksize = kw.get("pad", (0, 0, 0, 0)), kw.get("ksize", (1, 1))
            stride, dil = kw.get("stride", (1, 1)), kw.get("dilation", (1, 1))
            n, h, w, _ = x.shape
            oh = output_dim(h, pad[0], pad[2], ksize[0], dil[0], stride[0])
            ow = output_dim(w, pad[3], pad[1], ksize[1], dil[1], stride[1])
            out_stride = kw.get("groups", 1) * kw["goc"] if out_stride is None else out_stride
            buf = np.zeros(lead_in + x.size + 16, dtype=np.uint8)
            xin = buf[lead_in:lead_in + x.size].reshape(x.shape)
            xin[...] = x
            out = np.full((n, oh, ow, out_stride), out_fill, dtype=np.uint8)
            st = self.setup_convolution(op, xin, out)
            if st != 0:
                raise RuntimeError(f"setup -> {_capi.STATUS_NAMES.get(st, st)}")
            st = self.run(op)
            if st != 0:
                raise RuntimeError(f"run -> {_capi.STATUS_NAMES.get(st, st)}")
            return out
        finally:
            self.delete(op)

    def fully_connected(self, x, kernel, bias, *, out_stride=None, out_fill=0xA5, lead_in=16, **kw):
        st, op = self.create_fully_connected(kernel, bias, **kw)
        if st != 0:
            raise RuntimeError(f"create -> {_capi.STATUS_NAMES.get(st, st)}")
        try:
            b, in_stride = x.shape
            oc = kernel.shape[0]
            out_stride = oc if out_stride is None else out_stride
            buf = np.zeros(lead_in + x.size + 16, dtype=np.uint8)
            xin = buf[lead_in:lead_in + x.size].reshape(x.shape)
            xin[...] = x
            out = np.full((b, out_stride), out_fill, dtype=np.uint8)
            st = self.lib.qnnp_setup_fully_connected_nc_q8(op, b, _ptr(xin), in_stride, _ptr(out), out_stride)
            if st != 0:
                raise RuntimeError(f"setup -> {_capi.STATUS_NAMES.get(st, st)}")
            st = self.run(op)
            if st != 0:
                raise RuntimeError(f"run -> {_capi.STATUS_NAMES.get(st, st)}")
            return out
        finally:
            self.delete(op)

    def requantize_q31(self, acc: np.ndarray, scale, zp, qmin, qmax, variant="scalar") -> np.ndarray:
        """src/requantization/q31-scalar.c:17 / q31-sse2.c (n must be a multiple of 4 / 16)."""
        fn = getattr(self.lib, f"qnnp_requantize_q31__{variant}")
        fn.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
        fn.restype = None
        acc = np.ascontiguousarray(acc, dtype=np.int32)
        assert acc.size % 16 == 0
        out = np.empty(acc.shape, dtype=np.uint8)
        fn(acc.size, _ptr(acc), float(np.float32(scale)), zp, qmin, qmax, _ptr(out))
        return out
