"""CPU tests of the drop-in boundary: libqnnpack.so builds, loads, exports every symbol that
include/qnnpack.h and include/qnnpack_cuda.h declare, and fails loudly without a B200."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from qnnpack_b200 import build
    return build.build()


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qnnp_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_api():
    from qnnpack_b200 import _capi
    assert _declared("qnnpack.h") == sorted(_capi.ALL_QNNPACK_H_SYMBOLS)


@pytest.mark.parametrize("header", ["qnnpack.h", "qnnpack_cuda.h"])
def test_library_exports_every_declared_symbol(lib_path, header):
    lib = C.CDLL(lib_path)
    missing = [s for s in _declared(header) if not hasattr(lib, s)]
    assert not missing, missing


def test_headers_compile_as_c_and_cxx(tmp_path):
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "qnnpack_cuda.h"\nint main(void) { return (int) qnnp_status_success; }\n')
    for cc, std in (("gcc", "-std=c99"), ("g++", "-std=c++11")):
        subprocess.check_call([cc, std, "-x", "c" if cc == "gcc" else "c++", "-fsyntax-only", "-Wall", "-Werror",
                               "-I", os.path.join(ROOT, "include"), str(src)])


def test_initialize_fails_loudly_without_a_b200(lib_path):
    import torch
    lib = C.CDLL(lib_path)
    st = lib.qnnp_initialize()
    if torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10:
        assert st == 0
    else:
        assert st == 4  # qnnp_status_unsupported_hardware — no CPU fallback exists
        op = C.c_void_p()
        # every create on an uninitialised library reports qnnp_status_uninitialized (src/convolution.c:69)
        lib.qnnp_create_fully_connected_nc_q8.restype = C.c_int
        st = lib.qnnp_create_fully_connected_nc_q8(C.c_size_t(8), C.c_size_t(8), 0, C.c_float(1.0), 0, C.c_float(1.0),
                                                   None, None, 0, C.c_float(2.0), 0, 255, 0, C.byref(op))
        assert st == 1
        assert lib.qnnp_delete_operator(None) == 2  # src/operator-delete.c:17


def test_product_python_api_raises_without_gpu(lib_path):
    import torch
    import qnnpack_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qnnpack_b200.QnnpackError):
        qnnpack_b200.load()
