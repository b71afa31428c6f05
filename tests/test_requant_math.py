"""CPU test of the product's epilogue arithmetic (qnnpack_b200/csrc/requant_math.h, the same header
the CUDA kernels include) compiled for the host: the fused one-multiply-one-shift form must equal the
oracle's two-rounding specification on every input, including ties and int32 extremes."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import q8_oracle as O
from tests.test_oracle import KAT_ZERO_POINTS, kat_exact_divide, kat_rounding_away, kat_rounding_up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = """
#include "requant_math.h"
#include <string.h>
#include <stddef.h>
extern "C" void hc_requant(size_t n, const int32_t* in, float scale, unsigned char zp, unsigned char qmin,
                           unsigned char qmax, unsigned char* out, int force_slow) {
  uint32_t bits; memcpy(&bits, &scale, 4);
  Q8Requant p = q8_make_requant(bits, zp, qmin, qmax);
  if (force_slow == 1) p.fused = 0;
  for (size_t i = 0; i < n; i++) out[i] = (unsigned char) q8_requant(in[i], p);
}
// the "U" form with the accumulator bound nmax; returns 0 (and leaves `out` alone) when the bound makes it ineligible
extern "C" int hc_requant_u(size_t n, const int32_t* in, float scale, unsigned char zp, unsigned char qmin,
                            unsigned char qmax, unsigned char* out, long long nmax) {
  uint32_t bits; memcpy(&bits, &scale, 4);
  Q8Requant p = q8_make_requant(bits, zp, qmin, qmax);
  q8_requant_enable_u(p, nmax);
  if (!p.u_ok || q8_requant_mode(p) < 5) return 0;
  for (size_t i = 0; i < n; i++) out[i] = (unsigned char) q8_requant(in[i], p);
  return 1;
}
"""


@pytest.fixture(scope="module")
def hostcheck(tmp_path_factory):
    d = tmp_path_factory.mktemp("hc")
    src = d / "hc.cpp"
    src.write_text(SRC)
    so = d / "hc.so"
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "qnnpack_b200", "csrc"),
                           "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.hc_requant.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p, C.c_int]

    lib.hc_requant_u.argtypes = [C.c_size_t, C.c_void_p, C.c_float, C.c_uint8, C.c_uint8, C.c_uint8, C.c_void_p,
                                 C.c_longlong]
    lib.hc_requant_u.restype = C.c_int

    def run(x, scale, zp, qmin, qmax, slow=0, nmax=None):
        x = np.ascontiguousarray(x, dtype=np.int32)
        out = np.empty(x.shape, dtype=np.uint8)
        if nmax is not None:
            ok = lib.hc_requant_u(x.size, x.ctypes.data, float(np.float32(scale)), zp, qmin, qmax, out.ctypes.data, nmax)
            return out if ok else None
        lib.hc_requant(x.size, x.ctypes.data, float(np.float32(scale)), zp, qmin, qmax, out.ctypes.data, slow)
        return out
    return run


def _scales():
    s = [np.float32(m * 2.0 ** e) for e in range(-31, 0) for m in (1.0, 1.25, 1.5, 1.9999999)]
    s += [np.float32(float.fromhex("0x1.FFFFFEp-1")), np.float32(2.0 ** -32), np.float32(0.75)]
    return [v for v in s if v < 1.0 and v >= np.float32(2.0 ** -32)]


def test_fused_requant_equals_specification(hostcheck, oracle_c):
    rng = np.random.default_rng(11)
    edge = np.array([0, 1, -1, 2, -2, 2**31 - 1, -2**31, 2**30, -2**30, 2**30 - 1, -2**30 + 1], dtype=np.int64)
    for scale in _scales():
        inv = 1.0 / float(scale)
        ties = np.array([int(round((k + 0.5) * inv)) + d for k in range(-130, 130) for d in (-1, 0, 1)], dtype=np.float64)
        x = np.concatenate([edge, rng.integers(-2**31, 2**31, 2000), rng.integers(-2**20, 2**20, 2000),
                            rng.integers(-300, 300, 1000) * inv, ties]).clip(-2**31, 2**31 - 1).astype(np.int32)
        for zp, qmin, qmax in ((0, 0, 255), (127, 1, 254), (255, 0, 255), (128, 128, 255), (100, 0, 128)):
            want = oracle_c.requantize_q31(x, scale, zp, qmin, qmax)
            assert np.array_equal(hostcheck(x, scale, zp, qmin, qmax), want), (scale, zp, qmin, qmax)
            assert np.array_equal(hostcheck(x, scale, zp, qmin, qmax, slow=1), want), (scale, zp, qmin, qmax)


def test_u_form_requant_equals_specification(hostcheck, oracle_c):
    """The 4-instruction "U" form (requant_math.h) under its accumulator bound |n| <= nmax, for bounds from a small
    1x1 layer up to the full int32 range; it must either declare itself ineligible or be exact."""
    rng = np.random.default_rng(12)
    eligible = 0
    for scale in _scales():
        inv = 1.0 / float(scale)
        for nmax in (2**31 - 1, 2**30, 2**27, 16 * 65025 + 2**20, 9 * 65025):
            ties = np.array([int(round((k + 0.5) * inv)) + d for k in range(-130, 130) for d in (-1, 0, 1)], dtype=np.float64)
            x = np.concatenate([[0, 1, -1, nmax, -nmax, nmax - 1, 1 - nmax], rng.integers(-nmax, nmax + 1, 3000),
                                rng.integers(-300, 300, 1000) * inv, ties]).clip(-nmax, nmax).astype(np.int32)
            for zp, qmin, qmax in ((0, 0, 255), (127, 1, 254), (255, 0, 255), (128, 128, 255), (100, 0, 128)):
                got = hostcheck(x, scale, zp, qmin, qmax, nmax=nmax)
                if got is None:
                    continue
                eligible += 1
                assert np.array_equal(got, oracle_c.requantize_q31(x, scale, zp, qmin, qmax)), (scale, nmax, zp, qmin, qmax)
    assert eligible > 1000


@pytest.mark.parametrize("s", range(1, 32))
def test_fused_requant_known_answers(hostcheck, s):
    for zp in KAT_ZERO_POINTS:
        for kat in (kat_exact_divide, kat_rounding_up, kat_rounding_away):
            x, want = kat(s, zp)
            assert np.array_equal(hostcheck(x, np.float32(2.0 ** -s), zp, 0, 255), want), (kat.__name__, s, zp)
