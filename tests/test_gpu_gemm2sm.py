"""GPU parity of the CTA-pair GEMM kernel (q8_gemm2sm_kernel: cta_group::2 UMMAs, 128-byte-swizzled TMA operands,
weights streamed from a K-contiguous packing): fully-connected / 1x1 operators whose weights do not fit shared memory.
Byte-exact against the C oracle; shapes cover ragged M (not a multiple of 256 / 128), ragged N (not a multiple of 240),
K that is not a multiple of the 128-byte stage, many tiles per cluster, clamps and zero points."""
import numpy as np
import pytest

from tests import cases as CS, util as U

pytestmark = pytest.mark.gpu

CASES = [
    CS.fc_case("g2_m1000_k4096_n500", 1000, 4096, 500),
    CS.fc_case("g2_m256_k2048_n240", 256, 2048, 240),
    CS.fc_case("g2_m300_k1040_n250", 300, 1040, 250),          # K = 8 * 128 + 16
    CS.fc_case("g2_m4100_k1024_n1024", 4100, 1024, 1024),
    CS.fc_case("g2_m513_k1536_n700_zp", 513, 1536, 700, izp=3, kzp=250),
    CS.fc_case("g2_m640_k1024_n960_qmin", 640, 1024, 960, qmin=128),
    CS.fc_case("g2_m384_k2048_n512_kzp0", 384, 2048, 512, kzp=0),
]


@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("ctas", [0, 2])
@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_gemm2sm(gpu_lib, oracle_c, case, ctas, monkeypatch):
    if ctas:
        monkeypatch.setenv("QNNP_CUDA_MAX_CTAS", str(ctas))   # one cluster walks every tile
    x, k, b, kw = U.fc_setup(case)
    got = U.run_fc(gpu_lib, case, x, k, b, kw)
    U.assert_same_bytes(got, U.run_fc(oracle_c, case, x, k, b, kw), case["name"])


@pytest.mark.timeout(300, method="thread")
def test_gemm2sm_is_selected_and_matches_the_single_cta_kernel(gpu_lib, monkeypatch):
    """The same operator through both kernels (QNNP_CUDA_NO_GEMM2SM routes to the single-CTA kernel): same bytes."""
    case = CS.fc_case("g2_vs_1sm", 1500, 2048, 1000)
    x, k, b, kw = U.fc_setup(case)
    a = U.run_fc(gpu_lib, case, x, k, b, kw)
    monkeypatch.setenv("QNNP_CUDA_NO_GEMM2SM", "1")
    c = U.run_fc(gpu_lib, case, x, k, b, kw)
    U.assert_same_bytes(a, c, "pair kernel vs single-CTA kernel")
