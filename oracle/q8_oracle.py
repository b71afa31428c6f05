"""TEST INFRASTRUCTURE — CPU oracle for the q8 hot path.  NOT PRODUCT CODE.

Two restatements of the reference's arithmetic live here:

* ``COracle``  — ctypes binding of oracle/q8_oracle.c (scalar C, one element at a time);
* the ``*_np`` functions — a vectorised NumPy restatement for shapes too large to loop in C tests.

Both are pinned against the unmodified compiled reference (oracle/ref.py, oracle/_ref/) by
tests/test_oracle_vs_ref.py and against tests/golden/*.npz.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.

Reference citations (paths relative to the reference root) are on each function.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_C_LIB = os.path.join(HERE, "_build", "libq8oracle.so")


# --------------------------------------------------------------------------------------------
# NumPy restatement
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RequantParams:
    """The .scalar view of union qnnp_conv_quantization_params (src/qnnpack/params.h:128-138)."""
    multiplier: int
    remainder_mask: int
    remainder_threshold: int
    shift: int
    min_less_zero_point: int
    max_less_zero_point: int
    zero_point: int


def requant_scale(input_scale: float, kernel_scale: float, output_scale: float) -> np.float32:
    """fp32, in exactly this order: src/convolution.c:161, src/fully-connected.c:71."""
    return np.float32(np.float32(input_scale) * np.float32(kernel_scale)) / np.float32(output_scale)


def compute_requant_params(scale, zero_point: int, qmin: int, qmax: int) -> RequantParams:
    """src/qnnpack/requantization.h:122-198 (scalar branch) — uses only the bits of the fp32 scale."""
    scale = np.float32(scale)
    assert scale < 1.0 and scale >= np.float32(2.0 ** -32), "requantization scale must be in [2^-32, 1)"
    bits = int(np.array(scale, dtype=np.float32).view(np.uint32))
    multiplier = ((bits & 0x007FFFFF) | 0x00800000) << 7
    shift = 127 + 31 - 32 - (bits >> 23)
    mask = (1 << shift) - 1
    return RequantParams(multiplier, mask, mask >> 1, shift, qmin - zero_point, qmax - zero_point, zero_point)


def q31_requantize_np(acc: np.ndarray, p: RequantParams) -> np.ndarray:
    """src/qnnpack/requantization.h:464-480, element-wise over an int32 array."""
    n = np.asarray(acc, dtype=np.int32).astype(np.int64)
    product = n * np.int64(p.multiplier)
    q31 = ((product + np.int64(0x40000000)) >> np.int64(31)).astype(np.int32).astype(np.int64)
    rem = (q31 & np.int64(p.remainder_mask)) - (n < 0).astype(np.int64)
    y = (q31 >> np.int64(p.shift)) + (rem > p.remainder_threshold).astype(np.int64)
    y = np.clip(y, p.min_less_zero_point, p.max_less_zero_point) + p.zero_point
    return y.astype(np.uint8)


def output_dim(in_dim: int, pad_a: int, pad_b: int, k: int, dil: int, stride: int) -> int:
    """src/convolution.c:29-37."""
    return (pad_a + in_dim + pad_b - ((k - 1) * dil + 1)) // stride + 1


def conv_accumulators_np(x, kernel, bias, *, pad, ksize, stride, dilation, groups, gic, goc, izp, kzp):
    """int32 accumulators  bias + sum (x - izp)(w - kzp)  with out-of-bounds taps contributing
    zero (padded taps read the byte izp: src/convolution.c:336, src/indirection.c:64,71), i.e. what
    test/convolution-operator-tester.h:367-403 computes.  Equal, modulo 2^32, to the reference's
    packed-bias form (src/qnnpack/pack.h:24-43,63-84,146-159).

    x: uint8 [N, H, W, groups*gic] (dense);  kernel: uint8 [groups, goc, KH, KW, gic];
    bias: int32 [groups*goc].  Returns int32 [N, OH, OW, groups*goc].
    """
    pt, pr, pb, pl = pad
    kh, kw = ksize
    sh, sw = stride
    dh, dw = dilation
    n, h, w, _ = x.shape
    oh = output_dim(h, pt, pb, kh, dh, sh)
    ow = output_dim(w, pl, pr, kw, dw, sw)
    xs = x.astype(np.int64) - int(izp)
    ks = kernel.reshape(groups, goc, kh, kw, gic).astype(np.int64) - int(kzp)
    acc = np.zeros((n, oh, ow, groups, goc), dtype=np.int64)
    acc += np.asarray(bias, dtype=np.int64).reshape(1, 1, 1, groups, goc)
    oy = np.arange(oh)
    ox = np.arange(ow)
    for ky in range(kh):
        iy = oy * sh + ky * dh - pt
        vy = (iy >= 0) & (iy < h)
        for kx in range(kw):
            ix = ox * sw + kx * dw - pl
            vx = (ix >= 0) & (ix < w)
            if not vy.any() or not vx.any():
                continue
            patch = xs[:, iy[vy]][:, :, ix[vx]]  # [n, oy', ox', C]
            patch = patch.reshape(n, int(vy.sum()), int(vx.sum()), groups, gic)
            contrib = np.einsum("nyxgc,goc->nyxgo", patch, ks[:, :, ky, kx, :], optimize=True)
            acc[np.ix_(np.arange(n), np.nonzero(vy)[0], np.nonzero(vx)[0])] += contrib
    acc = acc.reshape(n, oh, ow, groups * goc)
    # int32 wrap-around like the reference's accumulators
    return ((acc + 2**31) % 2**32 - 2**31).astype(np.int32)


def convolution2d_nhwc_q8_np(x, kernel, bias, *, pad=(0, 0, 0, 0), ksize=(1, 1), stride=(1, 1),
                             dilation=(1, 1), groups=1, gic=None, goc=None, izp=0, input_scale=1.0,
                             kzp=0, kernel_scale=1.0, ozp=0, output_scale=1.0, qmin=0, qmax=255):
    """create+setup+run of qnnp_convolution2d_nhwc_q8 (src/convolution.c:39-492,
    src/operator-run.c:647-844) on a dense NHWC uint8 tensor; returns dense NHWC uint8."""
    acc = conv_accumulators_np(x, kernel, bias, pad=pad, ksize=ksize, stride=stride, dilation=dilation,
                               groups=groups, gic=gic, goc=goc, izp=izp, kzp=kzp)
    p = compute_requant_params(requant_scale(input_scale, kernel_scale, output_scale), ozp, qmin, qmax)
    return q31_requantize_np(acc, p)


def fully_connected_nc_q8_np(x, kernel, bias, *, izp, input_scale, kzp, kernel_scale, ozp, output_scale,
                             qmin=0, qmax=255):
    """src/fully-connected.c:25-161: x uint8 [batch, IC], kernel uint8 [OC, IC] -> uint8 [batch, OC]."""
    acc = (x.astype(np.int64) - int(izp)) @ (kernel.astype(np.int64) - int(kzp)).T + np.asarray(bias, np.int64)
    acc = ((acc + 2**31) % 2**32 - 2**31).astype(np.int32)
    p = compute_requant_params(requant_scale(input_scale, kernel_scale, output_scale), ozp, qmin, qmax)
    return q31_requantize_np(acc, p)


# --------------------------------------------------------------------------------------------
# C restatement (oracle/q8_oracle.c)
# --------------------------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "q8_oracle.c")
    if force or not os.path.exists(_C_LIB) or os.path.getmtime(_C_LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_C_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-Wall", "-o", _C_LIB, src])
    return _C_LIB


class _CParams(C.Structure):
    _fields_ = [("multiplier", C.c_int32), ("remainder_mask", C.c_int32), ("remainder_threshold", C.c_int32),
                ("shift", C.c_uint32), ("min_less_zero_point", C.c_int32), ("max_less_zero_point", C.c_int32),
                ("zero_point", C.c_int32)]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    """ctypes driver for oracle/q8_oracle.c."""

    def __init__(self):
        self.lib = C.CDLL(build_c_oracle())
        L = self.lib
        L.q8o_compute_requant_params.argtypes = [C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.POINTER(_CParams)]
        L.q8o_q31_requantize.argtypes = [C.c_int32, C.POINTER(_CParams)]
        L.q8o_q31_requantize.restype = C.c_uint8
        L.q8o_requantize_q31.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p]
        L.q8o_convolution2d_nhwc_q8.argtypes = (
            [C.c_size_t] * 3 + [C.c_uint32] * 4 + [C.c_uint32] * 6 + [C.c_uint32, C.c_size_t, C.c_size_t]
            + [C.c_uint8, C.c_float, C.c_uint8, C.c_float, C.c_void_p, C.c_void_p]
            + [C.c_uint8, C.c_float, C.c_uint8, C.c_uint8]
            + [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t])
        L.q8o_convolution2d_nhwc_q8.restype = C.c_int
        L.q8o_gemm_accumulators.argtypes = [C.c_size_t] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                                               C.c_uint8, C.c_uint8, C.c_void_p]

    def requant_params(self, scale, zp, qmin, qmax) -> RequantParams:
        p = _CParams()
        self.lib.q8o_compute_requant_params(float(np.float32(scale)), zp, qmin, qmax, C.byref(p))
        return RequantParams(p.multiplier, p.remainder_mask, p.remainder_threshold, p.shift,
                             p.min_less_zero_point, p.max_less_zero_point, p.zero_point)

    def requantize_q31(self, acc: np.ndarray, scale, zp, qmin, qmax) -> np.ndarray:
        acc = np.ascontiguousarray(acc, dtype=np.int32)
        out = np.empty(acc.shape, dtype=np.uint8)
        self.lib.q8o_requantize_q31(acc.size, _ptr(acc), float(np.float32(scale)), zp, qmin, qmax, _ptr(out))
        return out

    def convolution(self, x, kernel, bias, *, pad=(0, 0, 0, 0), ksize=(1, 1), stride=(1, 1), dilation=(1, 1),
                    groups=1, gic, goc, izp, input_scale, kzp, kernel_scale, ozp, output_scale, qmin=0, qmax=255,
                    in_stride=None, out_stride=None, out_fill=0xA5):
        """x: uint8 [N, H, W, in_stride] (only the first groups*gic bytes of a pixel are read).
        Returns uint8 [N, OH, OW, out_stride], untouched bytes keep ``out_fill``."""
        n, h, w, xs = x.shape
        in_stride = xs if in_stride is None else in_stride
        assert in_stride == xs
        out_stride = groups * goc if out_stride is None else out_stride
        oh = output_dim(h, pad[0], pad[2], ksize[0], dilation[0], stride[0])
        ow = output_dim(w, pad[3], pad[1], ksize[1], dilation[1], stride[1])
        x = np.ascontiguousarray(x)
        kernel = np.ascontiguousarray(kernel, dtype=np.uint8)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        out = np.full((n, oh, ow, out_stride), out_fill, dtype=np.uint8)
        rc = self.lib.q8o_convolution2d_nhwc_q8(
            n, h, w, pad[0], pad[1], pad[2], pad[3], ksize[0], ksize[1], stride[0], stride[1], dilation[0],
            dilation[1], groups, gic, goc, izp, float(np.float32(input_scale)), kzp, float(np.float32(kernel_scale)),
            _ptr(kernel), _ptr(bias), ozp, float(np.float32(output_scale)), qmin, qmax,
            _ptr(x), in_stride, _ptr(out), out_stride)
        if rc != 0:
            raise ValueError("requantization scale outside [2^-32, 1)")
        return out

    def fully_connected(self, x, kernel, bias, *, izp, input_scale, kzp, kernel_scale, ozp, output_scale,
                        qmin=0, qmax=255, out_stride=None, out_fill=0xA5):
        b, xs = x.shape
        oc, ic = kernel.shape
        y = self.convolution(x.reshape(1, b, 1, xs), kernel, bias, gic=ic, goc=oc, izp=izp, input_scale=input_scale,
                             kzp=kzp, kernel_scale=kernel_scale, ozp=ozp, output_scale=output_scale, qmin=qmin,
                             qmax=qmax, out_stride=out_stride, out_fill=out_fill)
        return y.reshape(b, -1)

    def gemm_accumulators(self, a, b, bias, azp, bzp) -> np.ndarray:
        m, k = a.shape
        n = b.shape[0]
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        acc = np.empty((m, n), dtype=np.int32)
        self.lib.q8o_gemm_accumulators(m, n, k, _ptr(a), k, _ptr(b), _ptr(bias), azp, bzp, _ptr(acc))
        return acc
