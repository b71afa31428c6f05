"""qnnpack_b200 — B200-native implementation of QNNPACK's q8gemm / q8conv / q8dwconv hot path.

The product is the C-ABI shared library ``qnnpack_b200/lib/libqnnpack.so`` (sources in ``csrc/``,
public headers in ``/include``).  This package only builds it (``build``) and mirrors the qnnpack.h
operator interface for Python callers, tests and benchmarks (``api``).
"""
from .api import PRODUCT_LIB, QnnpackError, QnnpackLibrary, load, output_dim  # noqa: F401
