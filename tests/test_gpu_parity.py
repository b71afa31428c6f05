"""GPU parity tests: the sm_100a kernels, called through the C ABI of libqnnpack.so, must produce
byte-identical outputs to (a) the golden vectors generated from the unmodified compiled reference and
(b) the oracle on the same seeded inputs — including the bytes BETWEEN output pixels, which must be
left untouched (0xA5 canary; the reference never writes them, SURVEY.md §7 trap 7).

Case grids restate test/q8gemm.cc, test/q8conv.cc, test/q8dwconv.cc, test/convolution.cc and
test/fully-connected.cc with fixed seeds (tests/cases.py)."""
import numpy as np
import pytest

from oracle import q8_oracle as O
from tests import cases as CS, util as U

pytestmark = pytest.mark.gpu


# ---- epilogue as a stand-alone kernel (test/requantization.cc Q31 rows) -----------------------------
def test_device_requant_known_answers(gpu_lib):
    from tests.test_oracle import kat_exact_divide, kat_rounding_away, kat_rounding_up
    for s in range(1, 32):
        for zp in range(0, 256, 5):
            for kat in (kat_exact_divide, kat_rounding_up, kat_rounding_away):
                x, want = kat(s, zp)
                assert np.array_equal(gpu_lib.requantize_q31(x, np.float32(2.0 ** -s), zp, 0, 255), want), (kat.__name__, s, zp)
    for zp in range(256):
        lo = gpu_lib.requantize_q31(np.full(16, -2**31, np.int32), np.float32(2.0 ** -32), zp, 0, 255)
        assert lo.min() == max(0, zp - 1)
    hi = gpu_lib.requantize_q31(np.full(16, 2**31 - 1, np.int32), np.float32(float.fromhex("0x1.FFFFFEp-1")), 255, 0, 255)
    assert (hi == 255).all()


def test_device_requant_random_matches_oracle(gpu_lib, oracle_c):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.integers(-2**31, 2**31, 1 << 18), rng.integers(-2**18, 2**18, 1 << 18)]).astype(np.int32)
    for scale, zp, qmin, qmax in ((0.75, 127, 1, 254), (2.0 ** -11 * 1.3, 3, 0, 255), (2.0 ** -31, 255, 0, 200),
                                  (0.5, 0, 0, 255), (2.0 ** -24 * 1.1, 9, 7, 99), (2.0 ** -25 * 1.7, 200, 0, 255)):
        assert np.array_equal(gpu_lib.requantize_q31(x, scale, zp, qmin, qmax), oracle_c.requantize_q31(x, scale, zp, qmin, qmax))


# ---- q8dwconv -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CS.DW_UKERNEL_CASES, ids=lambda c: c["name"])
def test_q8dwconv(gpu_lib, golden, case):
    x, k, b, kw = U.conv_setup(case)
    U.assert_same_bytes(U.run_conv(gpu_lib, case, x, k, b, kw), golden[f"conv/{case['name']}/y"], case["name"])


# depthwise shapes that take the tcgen05 path (channels % 16 == 0, dense pixels): geometry classes of its planner —
# 16-row tiles vs whole images stacked, one vs two parity planes, ragged tiles, every weight-operand mode, clamps


@pytest.mark.parametrize("case", CS.DW_TC_CASES, ids=lambda c: c["name"])
def test_q8dwconv_tensor_core_path(gpu_lib, oracle_c, case, monkeypatch):
    monkeypatch.setenv("QNNP_CUDA_DW_UMMA", "1")   # also where the router would prefer the CUDA-core kernel
    x, k, b, kw = U.conv_setup(case)
    before = gpu_lib.dw_umma_launch_count()
    got = U.run_conv(gpu_lib, case, x, k, b, kw)
    assert gpu_lib.dw_umma_launch_count() == before + 1, "expected the tcgen05 depthwise kernel"
    U.assert_same_bytes(got, U.run_conv(oracle_c, case, x, k, b, kw), case["name"])


# every form of the tcgen05 depthwise kernel on the same shapes: channel-pair form forced on / off, weight blocks resident or
# travelling with every item, 16-byte stores (the router's own choice is what test_q8dwconv_tensor_core_path runs)
DW_FORMS = [dict(QNNP_CUDA_DW_PAIR="1"), dict(QNNP_CUDA_DW_NO_PAIR="1"), dict(QNNP_CUDA_DW_PAIR="1", QNNP_CUDA_DW_B_STREAM="1"),
            dict(QNNP_CUDA_DW_NO_PAIR="1", QNNP_CUDA_DW_B_STREAM="1", QNNP_CUDA_DW_STORE16="1")]
DW_FORM_CASES = [c for c in CS.DW_TC_CASES if c["name"] in (
    "tc_c32_14x14", "tc_c32_s2_rows", "tc_c48_7x7_stack2", "tc_c160_112", "tc_c144_56_b2_negated", "tc_c48_s2_rows_odd_groups",
    "tc_c144_s2_28_b3", "tc_c192_28_b3", "tc_c384_14_b5", "tc_c32_out_stride", "tc_c64_kzp0_u8", "tc_c64_kzp128_s8")]


@pytest.mark.parametrize("form", DW_FORMS, ids=lambda f: "+".join(k[13:].lower() for k in f))
@pytest.mark.parametrize("case", DW_FORM_CASES, ids=lambda c: c["name"])
def test_q8dwconv_tensor_core_forms(gpu_lib, oracle_c, case, form, monkeypatch):
    monkeypatch.setenv("QNNP_CUDA_DW_UMMA", "1")
    for k, v in form.items():
        monkeypatch.setenv(k, v)
    x, k, b, kw = U.conv_setup(case)
    before = gpu_lib.dw_umma_launch_count()
    got = U.run_conv(gpu_lib, case, x, k, b, kw)
    assert gpu_lib.dw_umma_launch_count() == before + 1, "expected the tcgen05 depthwise kernel"
    U.assert_same_bytes(got, U.run_conv(oracle_c, case, x, k, b, kw), case["name"])


# 3x3 over 3 dense channels (the MobileNetV2 stem shape class): run loader fed from bulk-staged raw rows; the items
# of these cases straddle image boundaries and rows that are not multiples of 16 bytes


@pytest.mark.parametrize("case", CS.STEM_CASES, ids=lambda c: c["name"])
def test_stem_conv_raw_row_loader(gpu_lib, oracle_c, case):
    x, k, b, kw = U.conv_setup(case)
    U.assert_same_bytes(U.run_conv(gpu_lib, case, x, k, b, kw), U.run_conv(oracle_c, case, x, k, b, kw), case["name"])


# ---- q8gemm through the fully-connected operator ------------------------------------------------------
@pytest.mark.parametrize("case", CS.GEMM_UKERNEL_CASES, ids=lambda c: c["name"])
def test_q8gemm(gpu_lib, golden, case):
    x, k, b, kw = U.fc_setup(case)
    U.assert_same_bytes(U.run_fc(gpu_lib, case, x, k, b, kw), golden[f"fc/{case['name']}/y"], case["name"])


# ---- convolution operator (gemm / conv / dwconv / direct routes) ---------------------------------------
@pytest.mark.parametrize("case", CS.OPERATOR_CASES, ids=lambda c: c["name"])
def test_convolution_operator(gpu_lib, golden, case):
    x, k, b, kw = U.conv_setup(case)
    U.assert_same_bytes(U.run_conv(gpu_lib, case, x, k, b, kw), golden[f"conv/{case['name']}/y"], case["name"])


def test_kernel_routing(gpu_lib):
    """Which kernel family serves which shape (reference selection: src/convolution.c:180-189)."""
    def kind(**kw):
        kernel = np.zeros((kw["groups"], kw["goc"], kw["ksize"][0], kw["ksize"][1], kw["gic"]), np.uint8)
        st, op = gpu_lib.create_convolution(kernel, np.zeros(kw["groups"] * kw["goc"], np.int32), izp=0, input_scale=1.0,
                                            kzp=0, kernel_scale=1.0, ozp=0, output_scale=2.0, **kw)
        assert st == 0
        name = gpu_lib.kernel_name(op)
        gpu_lib.delete(op)
        return name
    assert kind(groups=1, gic=32, goc=16, ksize=(1, 1)) == "igemm-gemm"
    assert kind(groups=1, gic=32, goc=16, ksize=(1, 1), stride=(2, 2)) == "igemm-conv"
    assert kind(groups=1, gic=3, goc=32, ksize=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)) == "igemm-conv"
    assert kind(groups=32, gic=1, goc=1, ksize=(3, 3), pad=(1, 1, 1, 1)) == "dwconv3x3"
    assert kind(groups=32, gic=1, goc=1, ksize=(5, 5), pad=(2, 2, 2, 2)) == "direct"
    assert kind(groups=2, gic=8, goc=8, ksize=(1, 1)) == "direct"


# ---- API behaviour (error order: src/convolution.c:69-168; batch 0: :396-399, operator-run.c:642) ------
def test_status_codes(gpu_lib):
    k = np.zeros((1, 4, 1, 1, 4), np.uint8)
    b = np.zeros(4, np.int32)
    base = dict(gic=4, goc=4, izp=0, input_scale=1.0, kzp=0, kernel_scale=1.0, ozp=0, output_scale=2.0)
    assert gpu_lib.create_convolution(k, b, **{**base, "ksize": (0, 1)})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "stride": (1, 0)})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "dilation": (0, 1)})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "input_scale": 0.0})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "kernel_scale": float("inf")})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "output_scale": -1.0})[0] == 2
    assert gpu_lib.create_convolution(k, b, **{**base, "output_scale": 1.0})[0] == 3   # scale >= 1
    assert gpu_lib.create_fully_connected(k.reshape(4, 4), b, izp=0, input_scale=1.0, kzp=0, kernel_scale=4.0, ozp=0,
                                          output_scale=2.0)[0] == 3
    assert gpu_lib.delete(None) == 2
    st, op = gpu_lib.create_convolution(k, b, **base)
    assert st == 0
    x = np.zeros((1, 2, 2, 4), np.uint8)
    y = np.full((1, 2, 2, 4), 0xA5, np.uint8)
    assert gpu_lib.setup_convolution(op, 1, 0, 2, x, 4, y, 4) == 2       # zero height
    assert gpu_lib.setup_convolution(op, 0, 2, 2, x, 4, y, 4) == 0       # batch 0 is legal ...
    assert gpu_lib.run(op) == 0 and (y == 0xA5).all()                    # ... and run is a no-op
    assert gpu_lib.setup_convolution(op, 1, 2, 2, x, 4, y, 4) == 0       # re-setup is allowed
    assert gpu_lib.run(op) == 0 and (y == 0).all()
    assert gpu_lib.delete(op) == 0


def test_resetup_with_new_pointers_and_batch(gpu_lib, oracle_c):
    """setup borrows input/output; the same operator can be set up again (reference realloc's its tables)."""
    case = CS.conv_case("resetup", 2, 9, 9, 1, 32, 48)
    x, k, b, kw = U.conv_setup(case)
    st, op = gpu_lib.create_convolution(k, b, **kw)
    assert st == 0
    for n in (2, 1, 2):
        xi = np.ascontiguousarray(x[:n])
        out = np.full((n, 9, 9, 48), 0xA5, np.uint8)
        assert gpu_lib.setup_convolution(op, n, 9, 9, xi, 32, out, 48) == 0
        assert gpu_lib.run(op) == 0
        U.assert_same_bytes(out, oracle_c.convolution(xi, k, b, **kw), f"batch {n}")
    gpu_lib.delete(op)


# ---- MobileNetV2 layer shapes (bench/convolution.cc:453-537), batch 1, against reference digests ------
@pytest.mark.parametrize("entry", CS.MOBILENET_V2, ids=lambda e: e[0])
def test_mobilenet_v2_layer_batch1(gpu_lib, golden, entry):
    case = CS.mobilenet_case(entry, 1)
    x, k, b, kw = U.conv_setup(case)
    y = U.run_conv(gpu_lib, case, x, k, b, kw)
    assert U.digest(y) == str(golden[f"mnv2/{case['name']}/y_digest"]), case["name"]


# ---- persistent loops: many consecutive work items per CTA ----------------------------------------------------
# With 148 CTAs the shapes above give every CTA at most one or two items.  QNNP_CUDA_MAX_CTAS shrinks the grid so that
# each CTA walks a long item sequence: smem-ring wrap-around, TMEM accumulator-stage parity, staged bulk stores,
# raw-row ring, mixed-radix item stepping of the depthwise kernel — all compared byte for byte with the oracle.


@pytest.mark.timeout(120, method="thread")   # a wedged kernel must fail the test, not hang the box
@pytest.mark.parametrize("ctas", [1, 3])
@pytest.mark.parametrize("case", CS.PERSISTENT_CASES, ids=lambda c: c["name"])
def test_many_items_per_cta(gpu_lib, oracle_c, case, ctas, monkeypatch):
    monkeypatch.setenv("QNNP_CUDA_MAX_CTAS", str(ctas))
    monkeypatch.setenv("QNNP_CUDA_DW_UMMA", "1")
    x, k, b, kw = U.conv_setup(case)
    U.assert_same_bytes(U.run_conv(gpu_lib, case, x, k, b, kw), U.run_conv(oracle_c, case, x, k, b, kw), case["name"])
