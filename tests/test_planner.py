"""CPU test of the host logic that plans the tensor-core kernel (qnnp_cuda_debug_plan_igemm needs no GPU):
every plan must fit the B200's 227 KB of shared memory and 512 TMEM columns, keep >= 2 ring stages,
and folded mode (bias + zero-point correction as extra UMMAs) must keep the weights resident."""
import ctypes as C
import itertools

import pytest

NAMES = ("K nkc skc k_stages mt n_tiles n_tile n_mma has_corr b_res stages stage_B staging bias_B b_off bias_off a_off "
         "stage_off total bulk folded steps blk good").split()
SMEM_OPTIN = 232448


@pytest.fixture(scope="module")
def plan():
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    lib.qnnp_cuda_debug_plan_igemm.argtypes = [C.c_size_t, C.c_size_t, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int)]

    def f(k, n, groups=1, folded=0, steps=1):
        out = (C.c_int * 24)()
        ok = lib.qnnp_cuda_debug_plan_igemm(k, n, groups, folded, steps, out)
        return dict(zip(NAMES, out)) if ok else None
    return f


def check(d, k, n, folded):
    assert d["K"] == k and d["nkc"] % 2 == 0 and d["nkc"] * 16 >= k
    assert d["skc"] % 2 == 0 and 2 <= d["skc"] <= d["nkc"]
    assert d["k_stages"] == -(-d["nkc"] // d["skc"])
    assert d["n_tile"] % 16 == 0 and d["n_tiles"] * d["n_tile"] >= n
    assert d["n_mma"] == d["n_tile"] + (0 if folded else 16) and d["n_mma"] <= 256
    assert 1 <= d["mt"] <= 8 and d["mt"] * d["n_mma"] <= 256          # one TMEM accumulator stage = 256 columns
    assert d["stages"] >= 2 and d["stages"] <= 16
    assert d["total"] <= SMEM_OPTIN - 1024                             # leaves room for the static barrier block
    assert d["a_off"] % 128 == 0 and d["stage_B"] % 2048 == 0 or not d["b_res"]
    a_stage = d["mt"] * d["skc"] * 2048
    assert d["stage_B"] == a_stage + (0 if d["b_res"] else d["skc"] * d["n_mma"] * 16)
    if folded:
        assert d["b_res"] == 1 and d["folded"] == 1


def test_mobilenet_v2_plans(plan):
    from qnnpack_b200 import mobilenet_v2 as M
    for l in M.layers():
        if l.kind == "dw":
            continue
        for folded in (1, 0):
            d = plan(l.k_eff, l.cout, 1, folded, 1)
            if d is None:
                assert folded == 1, f"{l.name}: no plan at all"
                continue
            check(d, l.k_eff, l.cout, folded)


def test_plan_grid(plan):
    ks = [1, 3, 8, 9, 15, 16, 17, 27, 31, 32, 33, 64, 100, 144, 255, 256, 600, 1024, 1280, 4096, 11520]
    ns = [1, 4, 15, 16, 17, 24, 96, 144, 239, 240, 241, 256, 257, 1000, 1280, 4096]
    for k, n, folded in itertools.product(ks, ns, (0, 1)):
        d = plan(k, n, 1, folded, 2)
        if d is None:
            assert folded == 1 or k > 4096, (k, n)   # "ones" mode must always find a plan for sane sizes
            continue
        check(d, k, n, folded)


def test_grouped_plan(plan):
    for groups in (2, 3, 8):
        d = plan(40, 24, groups, 0, 0)
        check(d, 40, 24, 0)
