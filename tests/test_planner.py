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


# ---- panel epilogue (q8_igemm_sm100.cu, out_mode 2): staging writes replayed against the TMA swizzle definition ----
def _tma_swizzle(addr, width):  # noqa: E302
    """Shared-memory byte address -> physical address for a box whose inner extent equals the swizzle span: the 16-byte
    chunk bits [4, 4+b) are XORed with address bits [7, 7+b)  (SWIZZLE_128B b=3, 64B b=2, 32B b=1, none b=0)."""
    b = {128: 3, 64: 2, 32: 1, 16: 0}[width]
    return addr ^ (((addr >> 7) & ((1 << b) - 1)) << 4)


@pytest.mark.parametrize("folded", [0, 1])
@pytest.mark.parametrize("n_tile,mt", [(16, 8), (32, 8), (48, 5), (64, 4), (96, 2), (112, 2), (144, 1), (160, 1), (192, 1),
                                       (240, 1), (256, 1), (128, 2), (80, 3)])
def test_panel_epilogue_tables(n_tile, mt, folded):
    import numpy as np
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    out = (C.c_int * 50)()
    lib.qnnp_cuda_debug_panel_tables(n_tile, mt, folded, out)
    panels, box_rows = out[0], out[1]
    pan = [(out[2 + 4 * k], out[3 + 4 * k], out[4 + 4 * k], out[5 + 4 * k]) for k in range(panels)]
    assert sum(w for _, w, _, _ in pan) == n_tile and box_rows == (256 if mt % 2 == 0 else 128)
    W = 32 if folded else 16
    per_sub = (n_tile + W - 1) // W
    rows = mt * 128
    staging = np.full(rows * n_tile, -1, np.int64)           # value = row * 1000 + column
    phase_slots = {}
    for j in range(mt):
        for c in range(per_sub):
            x, y = out[18 + 2 * c], out[19 + 2 * c]
            pitch, lsh, mask = y & 0xFF, (y >> 8) & 0xFF, y >> 16
            width = min(W, n_tile - c * W)
            for row in range(128):
                jrow = j * 128 + row
                a0 = (x + jrow * pitch) ^ ((row << lsh) & mask)
                for h in range(width // 16):
                    a = a0 ^ (16 * h)
                    col = c * W + 16 * h
                    assert (staging[a:a + 16] == -1).all(), "two chunks land on the same bytes"
                    staging[a:a + 16] = jrow * 1000 + col + np.arange(16)
                    phase_slots.setdefault((j, c, h, row // 8), set()).add((a >> 4) & 7)
    assert (staging >= 0).all()
    # every 8-lane store phase (8 consecutive rows, same chunk) hits 8 different 16-byte bank groups
    assert all(len(v) == 8 for v in phase_slots.values())
    # the tensor stores read each panel through the hardware swizzle and must see row-major [row][column]
    for col0, width, off, cls in pan:
        assert off % 1024 == 0 and cls == {16: 0, 32: 1, 64: 2, 128: 3}[width]
        for r0 in range(0, rows, box_rows):
            for r in range(min(box_rows, rows - r0)):
                for cc in range(0, width, 16):
                    logical = off + (r0 + r) * width + cc          # dense box image relative to the box start ...
                    phys = off + r0 * width + _tma_swizzle(r * width + cc, width)  # ... as the TMA addresses it
                    assert (staging[phys:phys + 16] == (r0 + r) * 1000 + col0 + cc + np.arange(16)).all(), (col0, r0 + r, cc)
                    del logical


@pytest.mark.parametrize("folded", [0, 1])
@pytest.mark.parametrize("n_tile,mt", [(16, 8), (48, 5), (80, 3), (112, 2)])
def test_dense_epilogue_tables(n_tile, mt, folded):
    """Dense mode of the panel epilogue (narrow outputs with contiguous rows): every byte of the [row][N] image is written
    exactly once, where the 1-D bulk store expects it, and the store phases stay conflict-free (N / 16 odd)."""
    import numpy as np
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    out = (C.c_int * 50)()
    lib.qnnp_cuda_debug_panel_tables(n_tile, mt, folded | 2, out)
    assert out[0] == 0  # no panels
    W = 32 if folded else 16
    per_sub = (n_tile + W - 1) // W
    img = np.full(mt * 128 * n_tile, -1, np.int64)
    slots = {}
    for j in range(mt):
        for c in range(per_sub):
            x, y = out[18 + 2 * c], out[19 + 2 * c]
            pitch, lsh, mask = y & 0xFF, (y >> 8) & 0xFF, y >> 16
            assert pitch == n_tile and mask == 0
            width = min(W, n_tile - c * W)
            for row in range(128):
                jrow = j * 128 + row
                a0 = (x + jrow * pitch) ^ ((row << lsh) & mask)
                for h in range(width // 16):
                    a = a0 + 16 * h   # (kernel: a1 = a0 + 16 in dense mode)
                    assert (img[a:a + 16] == -1).all()
                    img[a:a + 16] = jrow * 1000 + c * W + 16 * h + np.arange(16)
                    slots.setdefault((j, c, h, row // 8), set()).add((a >> 4) & 7)
    want = (np.arange(mt * 128)[:, None] * 1000 + np.arange(n_tile)[None, :]).reshape(-1)
    assert (img == want).all()
    assert all(len(v) == 8 for v in slots.values())
