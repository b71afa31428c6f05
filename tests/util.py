"""Helpers shared by the parity tests: run a case through any qnnpack.h implementation or an oracle."""
from __future__ import annotations

import hashlib
import os
import sys             

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import q8_oracle as O  # noqa: E402
from tests import cases as CS  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "q8_golden.npz")


def case_seed(case) -> int:
    if case.get("seed") is not None:
        return case["seed"]
    return int.from_bytes(hashlib.sha256(case["name"].encode()).digest()[:4], "little")


def conv_setup(case):
    """-> (x, kernel, bias, create-kwargs) with the tester's data-derived output quantisation."""
    x, k, b = CS.make_conv_data(case, case_seed(case))
    cin = case["groups"] * case["gic"]
    acc = O.conv_accumulators_np(x[..., :cin], k, b, pad=case["pad"], ksize=case["ksize"], stride=case["stride"],
                                 dilation=case["dilation"], groups=case["groups"], gic=case["gic"], goc=case["goc"],
                                 izp=case["izp"], kzp=case["kzp"])
    oscale, ozp = CS.derive_output_quant(acc)
    kw = dict(pad=case["pad"], ksize=case["ksize"], stride=case["stride"], dilation=case["dilation"],
              groups=case["groups"], gic=case["gic"], goc=case["goc"], izp=case["izp"], input_scale=1.0,
              kzp=case["kzp"], kernel_scale=1.0, ozp=ozp, output_scale=oscale, qmin=case["qmin"], qmax=case["qmax"])
    return x, k, b, kw


def run_conv(lib, case, x, k, b, kw):
    """lib: anything with .convolution(x, kernel, bias, out_stride=..., **kw) (oracle, reference or product)."""
    out_stride = case["groups"] * case["goc"] + case["out_extra"]
    return lib.convolution(x, k, b, out_stride=out_stride, **kw)


def fc_setup(case):
    x, k, b = CS.make_fc_data(case, case_seed(case))
    acc = (x[:, :case["k"]].astype(np.int64) - case["izp"]) @ (k.astype(np.int64) - case["kzp"]).T + b.astype(np.int64)
    oscale, ozp = CS.derive_output_quant(acc)
    kw = dict(izp=case["izp"], input_scale=1.0, kzp=case["kzp"], kernel_scale=1.0, ozp=ozp, output_scale=oscale,
              qmin=case["qmin"], qmax=case["qmax"])
    return x, k, b, kw


def run_fc(lib, case, x, k, b, kw):
    return lib.fully_connected(x, k, b, out_stride=case["n"] + case["out_extra"], **kw)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def assert_same_bytes(got: np.ndarray, want: np.ndarray, what: str):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if not np.array_equal(got, want):
        bad = np.argwhere(got != want)
        first = tuple(bad[0])
        raise AssertionError(f"{what}: {len(bad)} of {got.size} bytes differ; first at {first}: "
                             f"got {got[first]} want {want[first]}")
