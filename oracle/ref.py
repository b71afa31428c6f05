"""TEST INFRASTRUCTURE — driver for the UNMODIFIED reference compiled into oracle/_ref/.

``oracle/_ref/libqnnpack_ref.so`` is built by ``make -C oracle ref`` from the sources where they lie
under /root/reference (never copied into this repo).  It is the strongest oracle we have (the
reference's own SSE2 micro-kernels through its own operator API) and the CPU baseline that
``bench.py --impl reference`` times.  The library travels to the GPU box inside the repo snapshot;
/root/reference itself does not, so nothing here reads it at run time.

Because the product keeps the qnnpack.h ABI, the same ctypes driver (qnnpack_b200.api.QnnpackLibrary)
serves both libraries.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from qnnpack_b200.api import QnnpackLibrary

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libqnnpack_ref.so")


def available() -> bool:
    return os.path.exists(REF_LIB)


class QnnpackHost(QnnpackLibrary):
    """The compiled reference (host memory only, optional pthreadpool with ``threads`` workers)."""

    def __init__(self, path: str = REF_LIB, threads: int = 0):
        super().__init__(path, threads=threads)

    def requantize_q31(self, acc: np.ndarray, scale, zp, qmin, qmax, variant="scalar") -> np.ndarray:
        """src/requantization/q31-scalar.c:17 / q31-sse2.c (n must be a multiple of 16)."""
        fn = getattr(self.lib, f"qnnp_requantize_q31__{variant}")
        fn.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
        fn.restype = None
        acc = np.ascontiguousarray(acc, dtype=np.int32)
        assert acc.size % 16 == 0
        out = np.empty(acc.shape, dtype=np.uint8)
        fn(acc.size, acc.ctypes.data_as(C.c_void_p), float(np.float32(scale)), zp, qmin, qmax,
           out.ctypes.data_as(C.c_void_p))
        return out
