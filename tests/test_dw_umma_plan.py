"""CPU replay of the depthwise tensor-core kernel's addressing (qnnpack_b200/csrc/q8_dwconv_umma_sm100.cu).

The host planner (qnnp_cuda_debug_plan_dwconv, callable without a GPU) prescribes the TMA boxes, the byte
offsets of the two taps each UMMA reads (a_off / a_lbo), the row-group stride (sbo) and the work-item grid.
This test plays the device's part in NumPy: it fills shared memory the way the TMA unit would (zero fill
outside the image), gathers every UMMA operand row exactly where the descriptors point, applies the
border-class bias, and compares the accumulators with a direct evaluation of the reference formula
    acc[c] = bias[c] + sum over taps inside the image of (a_tap[c] - izp) * (w_tap[c] - kzp)
(reference src/q8dwconv/up8x9-sse2.c:14-482 with the padding semantics of src/indirection.c:81-150).
Hardware semantics (descriptor encoding, TMEM lanes) are covered by the GPU parity tests."""
import ctypes as C

import numpy as np
import pytest

NAMES = ("G mt xt yt nt nb Q whole planes box_rows box_px plane_tx plane_bytes a_bytes b_bytes cg_bytes stage_bytes "
         "num_stages smem_total x_org0 x_org1 a_off0 a_off1 a_off2 a_off3 a_off4 a_lbo0 a_lbo1 a_lbo2 a_lbo3 a_lbo4 sbo "
         "nb_cols b_signed acc_stride cblocks cgs total_items").split()
T0 = (0, 3, 6, 1, 7)     # first tap (ky*3+kx) of UMMA u
T1 = (2, 5, 8, 4, -1)    # second tap (-1: empty slot)
SMEM_OPTIN = 232448


@pytest.fixture(scope="module")
def plan():
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    lib.qnnp_cuda_debug_plan_dwconv.argtypes = [C.c_int] * 10 + [C.POINTER(C.c_int)]

    def f(c, n, h, w, s, pad=(1, 1), wmode=2):
        oh = (h + 2 * pad[0] - 3) // s + 1
        ow = (w + 2 * pad[1] - 3) // s + 1
        out = (C.c_int * 40)()
        ok = lib.qnnp_cuda_debug_plan_dwconv(c, n, h, w, oh, ow, s, pad[0], pad[1], wmode, out)
        return (dict(zip(NAMES, out)), oh, ow) if ok else (None, oh, ow)
    return f


def replay(d, x, wk, bias, izp, kzp, s, pad, oh, ow):
    """-> accumulators [N][OH][OW][C] computed by following the plan."""
    n, h, w, c = x.shape
    dk = wk.astype(np.int64) - kzp                      # [C][9]
    acc = np.full((n, oh, ow, c), np.iinfo(np.int64).min, dtype=np.int64)
    written = np.zeros((n, oh, ow, c), dtype=np.int32)
    for item in range(d["total_items"]):
        r, cb = divmod(item, d["cblocks"])
        q, xtile = divmod(r, d["xt"])
        nblk, ytile = divmod(q, d["yt"])
        n0, oy0, ox0 = nblk * d["nb"], ytile * 16, xtile * d["mt"] * 8
        mt_eff = min(d["mt"], (ow - ox0 + 7) // 8)
        g_eff = min(d["G"], d["cgs"] - cb * d["G"])
        y0 = (0 if d["whole"] else oy0 * s) - pad[0]
        for gi in range(g_eff):
            cg = cb * d["G"] + gi
            # ---- TMA: one box per plane, zero fill outside the tensor
            smem = np.zeros(d["a_bytes"] + 64 * 1024, dtype=np.uint8)   # slack: a wrong plan would read stale zeros
            for par in range(d["planes"]):
                xo = ox0 + (d["x_org0"], d["x_org1"])[par]
                box = np.zeros((d["nb"], d["box_rows"], d["box_px"], 16), dtype=np.uint8)
                for i in range(d["nb"]):
                    for ry in range(d["box_rows"]):
                        for rx in range(d["box_px"]):
                            iy = y0 + ry
                            ix = (xo + rx) if s == 1 else 2 * (xo + rx) + par
                            inb = (xo + rx) >= 0 and ((xo + rx) < w if s == 1 else (xo + rx) < w // 2)
                            if n0 + i < n and 0 <= iy < h and inb and 0 <= ix < w:
                                box[i, ry, rx] = x[n0 + i, iy, ix, cg * 16:cg * 16 + 16]
                smem[par * d["plane_bytes"]:par * d["plane_bytes"] + d["plane_tx"]] = box.reshape(-1)
            # ---- UMMAs + epilogue for each sub-tile
            for j in range(mt_eff):
                a = np.zeros((128, 16), dtype=np.int64)
                for u in range(5):
                    for ch, tap in ((0, T0[u]), (1, T1[u])):
                        if tap < 0:
                            continue
                        base = d["a_off%d" % u] + ch * d["a_lbo%d" % u] + j * 128
                        for m in range(128):
                            off = base + (m // 8) * d["sbo"] + (m % 8) * 16
                            a[m] += smem[off:off + 16].astype(np.int64) * dk[cg * 16:cg * 16 + 16, tap]
                for m in range(128):
                    g, px = divmod(m, 8)
                    img, oyl = divmod(g, d["Q"])
                    nn, oy, ox = n0 + img, oy0 + oyl, ox0 + 8 * j + px
                    if not (img < d["nb"] and nn < n and oy < oh and ox < ow):
                        continue
                    iy0, ix0 = oy * s - pad[0], ox * s - pad[1]
                    ssum = np.zeros(16, dtype=np.int64)
                    for ky in range(3):
                        for kx in range(3):
                            if 0 <= iy0 + ky < h and 0 <= ix0 + kx < w:
                                ssum += dk[cg * 16:cg * 16 + 16, ky * 3 + kx]
                    acc[nn, oy, ox, cg * 16:cg * 16 + 16] = a[m] + bias[cg * 16:cg * 16 + 16] - izp * ssum
                    written[nn, oy, ox, cg * 16:cg * 16 + 16] += 1
    assert (written == 1).all(), "every output must be produced exactly once"
    return acc


def direct(x, wk, bias, izp, kzp, s, pad, oh, ow):
    n, h, w, c = x.shape
    out = np.zeros((n, oh, ow, c), dtype=np.int64) + bias.astype(np.int64)
    xi = x.astype(np.int64) - izp
    dk = wk.astype(np.int64) - kzp
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(3):
                for kx in range(3):
                    iy, ix = oy * s - pad[0] + ky, ox * s - pad[1] + kx
                    if 0 <= iy < h and 0 <= ix < w:
                        out[:, oy, ox, :] += xi[:, iy, ix, :] * dk[:, ky * 3 + kx]
    return out


CASES = [
    # (C, N, H, W, stride, pad)          geometry classes of MobileNetV2 at reduced sizes + odd ones
    (16, 1, 20, 23, 1, (1, 1)),          # row tiles with a ragged last tile, ragged columns, one channel group
    (48, 2, 7, 7, 1, (1, 1)),            # whole-image mode, two images stacked per item, odd group count
    (32, 3, 14, 14, 1, (1, 1)),          # whole-image mode, Q = 16, batch not a multiple of nb
    (32, 1, 40, 36, 2, (1, 1)),          # stride 2: parity planes, row tiles
    (16, 3, 14, 14, 2, (1, 1)),          # stride 2, whole-image mode (Q = 8, nb = 2)
    (32, 2, 9, 12, 1, (0, 0)),           # no padding
    (16, 1, 18, 16, 2, (0, 0)),          # stride 2 without padding: planes swap roles
    (16, 2, 5, 70, 1, (1, 1)),           # wide rows: several x tiles
    (16, 3, 28, 28, 2, (1, 1)),          # 28 -> 14 rows, stride 2: Q = 15, a second stacked image would lose rows
    (16, 3, 18, 6, 2, (0, 0)),           # Q = 9, 8 valid rows: likewise
]


@pytest.mark.parametrize("c,n,h,w,s,pad", CASES)
def test_plan_addressing_reproduces_the_convolution(plan, c, n, h, w, s, pad):
    d, oh, ow = plan(c, n, h, w, s, pad)
    assert d is not None
    assert d["smem_total"] <= SMEM_OPTIN - 1024 and d["num_stages"] >= 2
    assert d["acc_stride"] <= 256 and d["mt"] * d["G"] * d["nb_cols"] == d["acc_stride"]
    assert (d["nb"] - 1) * d["Q"] + oh <= 16 if d["whole"] else d["Q"] == 16   # every valid row inside the 16 groups
    for u in range(5):  # descriptor fields are 14 bits of 16-byte units
        assert d["a_off%d" % u] % 16 == 0 and d["a_lbo%d" % u] % 16 == 0 and 0 <= d["a_lbo%d" % u] < (1 << 18)
    assert d["sbo"] % 16 == 0 and d["sbo"] < (1 << 18)
    rng = np.random.default_rng(c * 1000 + h * 10 + s)
    x = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    wk = rng.integers(0, 256, (c, 9), dtype=np.uint8)
    bias = rng.integers(-10000, 10000, c).astype(np.int64)
    izp, kzp = 131, 77
    got = replay(d, x, wk, bias, izp, kzp, s, pad, oh, ow)
    want = direct(x, wk, bias, izp, kzp, s, pad, oh, ow)
    assert np.array_equal(got, want)


def _random_geometries(count, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < count:
        s = int(rng.integers(1, 3))
        c = int(rng.choice([16, 32, 48, 80]))
        n = int(rng.integers(1, 4))
        h = int(rng.integers(3, 26))
        w = int(rng.integers(3, 30))
        if s == 2 and w % 2:
            w += 1
        pad = (int(rng.integers(0, 3)), int(rng.integers(0, 3)))
        if (h + 2 * pad[0] - 3) // s + 1 < 1 or (w + 2 * pad[1] - 3) // s + 1 < 1:
            continue
        out.append((c, n, h, w, s, pad))
    return out


@pytest.mark.parametrize("c,n,h,w,s,pad", _random_geometries(24, 2026))
def test_plan_addressing_on_random_geometries(plan, c, n, h, w, s, pad):
    """Same replay on random shapes: odd sizes, paddings 0..2 on either axis, both strides, 1-3 images."""
    d, oh, ow = plan(c, n, h, w, s, pad)
    assert d is not None, (c, n, h, w, s, pad)
    assert d["smem_total"] <= SMEM_OPTIN - 1024 and d["acc_stride"] <= 256 and d["mt"] * d["G"] <= 16
    rng = np.random.default_rng(h * 131 + w * 7 + s)
    x = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    wk = rng.integers(0, 256, (c, 9), dtype=np.uint8)
    bias = rng.integers(-100000, 100000, c).astype(np.int64)
    izp, kzp = int(rng.integers(0, 256)), int(rng.integers(0, 256))
    assert np.array_equal(replay(d, x, wk, bias, izp, kzp, s, pad, oh, ow), direct(x, wk, bias, izp, kzp, s, pad, oh, ow))


def test_mobilenet_depthwise_layers_are_eligible_and_fit(plan):
    from qnnpack_b200 import mobilenet_v2 as M
    for l in M.layers():
        if l.kind != "dw":
            continue
        d, oh, ow = plan(l.cin, 4096, l.h, l.h, l.stride)
        assert d is not None, l
        assert d["num_stages"] >= 3 and d["smem_total"] <= SMEM_OPTIN - 1024, (l, d)


def test_ineligible_shapes_fall_back(plan):
    assert plan(24, 1, 8, 8, 1)[0] is None      # channels not a multiple of 16
    assert plan(16, 1, 9, 9, 2)[0] is None      # stride 2 needs an even width


def replay_packed(d, x, wpack, bias_cls, signed_b, s, pad, oh, ow, acc_sign=1):
    """Like replay(), but with the library's own packed operands: every UMMA multiplies the two 16-byte K-chunks the
    descriptors address (the empty tap slot included: its LBO is 0 and its weights must be zero) with the B block
    [2 chunks][nb_cols rows][16 B] of (channel group, u); the epilogue adds the two operand halves and bias_cls."""
    n, h, w, c = x.shape
    nbc = d["nb_cols"]
    ub = 2 * nbc * 16
    acc = np.zeros((n, oh, ow, c), dtype=np.int64)
    for item in range(d["total_items"]):
        r, cb = divmod(item, d["cblocks"])
        q, xtile = divmod(r, d["xt"])
        nblk, ytile = divmod(q, d["yt"])
        n0, oy0, ox0 = nblk * d["nb"], ytile * 16, xtile * d["mt"] * 8
        mt_eff = min(d["mt"], (ow - ox0 + 7) // 8)
        g_eff = min(d["G"], d["cgs"] - cb * d["G"])
        y0 = (0 if d["whole"] else oy0 * s) - pad[0]
        for gi in range(g_eff):
            cg = cb * d["G"] + gi
            smem = np.zeros(d["a_bytes"] + 64 * 1024, dtype=np.uint8)
            for par in range(d["planes"]):
                xo = ox0 + (d["x_org0"], d["x_org1"])[par]
                box = np.zeros((d["nb"], d["box_rows"], d["box_px"], 16), dtype=np.uint8)
                for i in range(d["nb"]):
                    for ry in range(d["box_rows"]):
                        for rx in range(d["box_px"]):
                            iy = y0 + ry
                            ix = (xo + rx) if s == 1 else 2 * (xo + rx) + par
                            inb = (xo + rx) >= 0 and ((xo + rx) < w if s == 1 else (xo + rx) < w // 2)
                            if n0 + i < n and 0 <= iy < h and inb and 0 <= ix < w:
                                box[i, ry, rx] = x[n0 + i, iy, ix, cg * 16:cg * 16 + 16]
                smem[par * d["plane_bytes"]:par * d["plane_bytes"] + d["plane_tx"]] = box.reshape(-1)
            for j in range(mt_eff):
                a = np.zeros((128, nbc), dtype=np.int64)
                for u in range(5):
                    blk = wpack[(cg * 5 + u) * ub:(cg * 5 + u + 1) * ub].reshape(2, nbc, 16)
                    bmat = (blk.view(np.int8) if signed_b else blk).astype(np.int64)
                    for ch in range(2):
                        base = d["a_off%d" % u] + ch * d["a_lbo%d" % u] + j * 128
                        rows = np.stack([smem[base + (m // 8) * d["sbo"] + (m % 8) * 16:][:16] for m in range(128)])
                        a += rows.astype(np.int64) @ bmat[ch].T
                for m in range(128):
                    g, px = divmod(m, 8)
                    img, oyl = divmod(g, d["Q"])
                    nn, oy, ox = n0 + img, oy0 + oyl, ox0 + 8 * j + px
                    if not (img < d["nb"] and nn < n and oy < oh and ox < ow):
                        continue
                    iy0, ix0 = oy * s - pad[0], ox * s - pad[1]
                    rm = sum(1 << ky for ky in range(3) if 0 <= iy0 + ky < h)
                    cm = sum(1 << kx for kx in range(3) if 0 <= ix0 + kx < w)
                    v = a[m, :16] + (a[m, 16:32] if nbc == 32 else 0)
                    acc[nn, oy, ox, cg * 16:cg * 16 + 16] = acc_sign * v + bias_cls[rm * 8 + cm, cg * 16:cg * 16 + 16]
    return acc


@pytest.mark.parametrize("kzp,izp", [(127, 131), (128, 7), (0, 255), (255, 0), (1, 128)])
@pytest.mark.parametrize("c,n,h,w,s,pad", [(32, 2, 12, 13, 1, (1, 1)), (16, 3, 14, 14, 2, (1, 1)), (48, 2, 7, 7, 1, (1, 1)),
                                            (16, 1, 22, 10, 2, (0, 1))])
def test_kernel_replay_on_the_packed_operands(plan, c, n, h, w, s, pad, kzp, izp):
    """End to end on the host: the bytes pack_dw_umma_host produces (diagonal B blocks, operand split, border-class
    biases) + the planner's addressing reproduce the reference accumulators for every weight-operand mode."""
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    lib.qnnp_cuda_debug_pack_dwconv.argtypes = [C.c_size_t, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_void_p]
    rng = np.random.default_rng(c + h + kzp)
    wk = rng.integers(0, 256, (c, 9), dtype=np.uint8)
    if kzp not in (0, 128):
        wk[0, 0], wk[-1, 8] = 0, 255                    # make sure the 9-bit range of w - kzp is really used
    bias = rng.integers(-50000, 50000, c).astype(np.int32)
    wpack = np.zeros((c // 16) * 5 * 2 * 32 * 16, dtype=np.uint8)
    bias_cls = np.zeros((64, c), dtype=np.int32)
    wmode = lib.qnnp_cuda_debug_pack_dwconv(c, izp, kzp, wk.ctypes.data, bias.ctypes.data, 0, wpack.ctypes.data,
                                            bias_cls.ctypes.data)
    dmin, dmax = int(wk.min()) - kzp, int(wk.max()) - kzp
    # 0: w - kzp fits s8; 3: kzp - w fits s8 (negated operand, e.g. kzp = 127: d in [-127, 128]); 2: two operands
    assert wmode == (1 if kzp == 0 else 0 if (dmin >= -128 and dmax <= 127) else 3 if (-dmax >= -128 and -dmin <= 127) else 2)
    oh, ow = (h + 2 * pad[0] - 3) // s + 1, (w + 2 * pad[1] - 3) // s + 1
    d = plan(c, n, h, w, s, pad, wmode)[0]
    assert d is not None and d["nb_cols"] == (32 if wmode == 2 else 16) and d["b_signed"] == (0 if wmode == 1 else 1)
    x = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    got = replay_packed(d, x, wpack, bias_cls.astype(np.int64), wmode != 1, s, pad, oh, ow, acc_sign=-1 if wmode == 3 else 1)
    assert np.array_equal(got, direct(x, wk, bias.astype(np.int64), izp, kzp, s, pad, oh, ow))
    # the "U" requantisation offset rides on the same table
    bias_u = np.zeros((64, c), dtype=np.int32)
    lib.qnnp_cuda_debug_pack_dwconv(c, izp, kzp, wk.ctypes.data, bias.ctypes.data, 1, wpack.ctypes.data, bias_u.ctypes.data)
    assert np.array_equal(bias_u.view(np.uint32), bias_cls.view(np.uint32) ^ np.uint32(0x80000000))


NAMES32 = NAMES + ["pair"] + ["a_off9_%d" % t for t in range(9)]


def replay_packed_pair(d, x, wpack32, bias_cls, signed_b, s, pad, oh, ow, acc_sign):
    """Channel-pair form of the kernel (DwTcParams::pair), replayed on the library's own plan and packed operands.
    Shared memory is modelled byte for byte: the TMA writes a box of 32-byte pixels with the 32-byte swizzle (16-byte chunk
    index XOR bit 7 of the byte's shared-memory address; planes start on 256-byte boundaries), every UMMA row m reads the
    32 bytes of the pixel its SWIZZLE_32B descriptor addresses — start = tap offset + sub-tile, 32 bytes per row, SBO per
    8-row group — through the same XOR, and multiplies them with the tap's [2 K-chunks][32 rows][16 B] block."""
    n, h, w, c = x.shape
    assert d["pair"] == 1 and d["nb_cols"] == 16 and d["G"] % 2 == 0
    acc = np.zeros((n, oh, ow, c), dtype=np.int64)
    written = np.zeros((n, oh, ow, c), dtype=np.int32)
    pairs_per_item = d["G"] // 2

    def sw(addr):  # physical address of logical byte `addr` of a swizzled plane region
        return addr ^ (((addr >> 7) & 1) << 4)

    for item in range(d["total_items"]):
        r, cb = divmod(item, d["cblocks"])
        q, xtile = divmod(r, d["xt"])
        nblk, ytile = divmod(q, d["yt"])
        n0, oy0, ox0 = nblk * d["nb"], ytile * 16, xtile * d["mt"] * 8
        mt_eff = min(d["mt"], (ow - ox0 + 7) // 8)
        g_eff = min(d["G"], d["cgs"] - cb * d["G"])
        y0 = (0 if d["whole"] else oy0 * s) - pad[0]
        for gp in range((g_eff + 1) // 2):
            pr = cb * pairs_per_item + gp                       # global channel pair
            smem = np.zeros(d["a_bytes"] + 64 * 1024, dtype=np.uint8)
            for par in range(d["planes"]):
                xo = ox0 + (d["x_org0"], d["x_org1"])[par]
                base = par * d["plane_bytes"]
                assert base % 256 == 0
                for i in range(d["nb"]):
                    for ry in range(d["box_rows"]):
                        for rx in range(d["box_px"]):
                            iy = y0 + ry
                            ix = (xo + rx) if s == 1 else 2 * (xo + rx) + par
                            inb = (xo + rx) >= 0 and ((xo + rx) < w if s == 1 else (xo + rx) < w // 2)
                            px = np.zeros(32, dtype=np.uint8)
                            if n0 + i < n and 0 <= iy < h and inb and 0 <= ix < w:
                                ch = x[n0 + i, iy, ix, pr * 32:pr * 32 + 32]    # channels beyond C: zero fill
                                px[:len(ch)] = ch
                            row = base + ((i * d["box_rows"] + ry) * d["box_px"] + rx) * 32
                            for k in range(32):
                                smem[sw(row + k)] = px[k]
            for j in range(mt_eff):
                a = np.zeros((128, 32), dtype=np.int64)
                for t in range(9):
                    blk = wpack32[(pr * 9 + t) * 1024:(pr * 9 + t + 1) * 1024].reshape(2, 32, 16)
                    bmat = (blk.view(np.int8) if signed_b else blk).astype(np.int64)       # [chunk][n][k % 16]
                    bfull = np.concatenate([bmat[0], bmat[1]], axis=1)                      # [n][k]
                    start = d["a_off9_%d" % t] + j * 256
                    rows = np.zeros((128, 32), dtype=np.int64)
                    for m in range(128):
                        ra = start + (m // 8) * d["sbo"] + (m % 8) * 32
                        rows[m] = [smem[sw(ra + k)] for k in range(32)]
                    a += rows @ bfull.T
                for m in range(128):
                    g, px_ = divmod(m, 8)
                    img, oyl = divmod(g, d["Q"])
                    nn, oy, ox = n0 + img, oy0 + oyl, ox0 + 8 * j + px_
                    if not (img < d["nb"] and nn < n and oy < oh and ox < ow):
                        continue
                    iy0, ix0 = oy * s - pad[0], ox * s - pad[1]
                    rm = sum(1 << ky for ky in range(3) if 0 <= iy0 + ky < h)
                    cm = sum(1 << kx for kx in range(3) if 0 <= ix0 + kx < w)
                    for half in range(2):
                        gi = 2 * gp + half
                        if gi >= g_eff:
                            continue
                        c0 = (cb * d["G"] + gi) * 16
                        acc[nn, oy, ox, c0:c0 + 16] = acc_sign * a[m, 16 * half:16 * half + 16] + bias_cls[rm * 8 + cm, c0:c0 + 16]
                        written[nn, oy, ox, c0:c0 + 16] += 1
    assert (written == 1).all()
    return acc


@pytest.mark.parametrize("kzp,izp", [(127, 131), (128, 7), (0, 255)])
@pytest.mark.parametrize("c,n,h,w,s,pad", [(32, 2, 12, 13, 1, (1, 1)), (96, 1, 20, 22, 2, (1, 1)), (48, 3, 7, 7, 1, (1, 1)),
                                            (32, 1, 22, 10, 2, (0, 1)), (144, 1, 9, 9, 1, (1, 1))])
def test_pair_form_replay_on_the_packed_operands(c, n, h, w, s, pad, kzp, izp):
    """The channel-pair form (32-byte pixels, SWIZZLE_32B tiles, one K = 32 UMMA per tap): the library's plan and packed
    operands reproduce the reference accumulators, every output exactly once — odd group counts (C = 48, 144) included."""
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    lib.qnnp_cuda_debug_plan_dwconv32.argtypes = [C.c_int] * 10 + [C.POINTER(C.c_int)]
    lib.qnnp_cuda_debug_pack_dwconv32.argtypes = [C.c_size_t, C.c_uint8, C.c_void_p, C.c_void_p]
    lib.qnnp_cuda_debug_pack_dwconv.argtypes = [C.c_size_t, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                                C.c_void_p]
    rng = np.random.default_rng(c * 3 + h + kzp)
    wk = rng.integers(0, 256, (c, 9), dtype=np.uint8)
    if kzp == 127:
        wk[0, 0], wk[-1, 8] = 0, 255
    bias = rng.integers(-50000, 50000, c).astype(np.int32)
    wpack32 = np.zeros(((c + 31) // 32) * 9 * 1024, dtype=np.uint8)
    wmode = lib.qnnp_cuda_debug_pack_dwconv32(c, kzp, wk.ctypes.data, wpack32.ctypes.data)
    assert wmode == {127: 3, 128: 0, 0: 1}[kzp]
    bias_cls = np.zeros((64, c), dtype=np.int32)
    wpack16 = np.zeros((c // 16) * 5 * 2 * 32 * 16, dtype=np.uint8)
    assert lib.qnnp_cuda_debug_pack_dwconv(c, izp, kzp, wk.ctypes.data, bias.ctypes.data, 0, wpack16.ctypes.data,
                                           bias_cls.ctypes.data) == wmode
    oh, ow = (h + 2 * pad[0] - 3) // s + 1, (w + 2 * pad[1] - 3) // s + 1
    out = (C.c_int * 48)()
    assert lib.qnnp_cuda_debug_plan_dwconv32(c, n, h, w, oh, ow, s, pad[0], pad[1], wmode, out)
    d = dict(zip(NAMES32, out))
    assert d["smem_total"] <= SMEM_OPTIN and d["stage_bytes"] == (d["G"] // 2) * d["cg_bytes"] and d["cg_bytes"] % 256 == 0
    x = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    got = replay_packed_pair(d, x, wpack32, bias_cls.astype(np.int64), wmode != 1, s, pad, oh, ow, -1 if wmode == 3 else 1)
    assert np.array_equal(got, direct(x, wk, bias.astype(np.int64), izp, kzp, s, pad, oh, ow))


def test_item_stepping_is_equivalent_to_decoding():
    """Model of the kernel's item walk (q8_dwconv_umma_sm100.cu: first_pos / advance_pos with the host's step digits):
    a CTA's k-th item, reached by adding the grid size in (cb, xtile, ytile, nblk) digits with carries, must be the item
    the flat index first + k * grid decodes to."""
    rng = np.random.default_rng(5)
    for _ in range(300):
        cblocks, xt, yt, nt = (int(v) for v in rng.integers(1, 9, 4))
        nt = int(rng.integers(1, 200))
        total = cblocks * xt * yt * nt
        grid = int(rng.integers(1, min(total, 148) + 1))
        r = grid
        step_cb, r = r % cblocks, r // cblocks
        step_x, r = r % xt, r // xt
        step_y, r = r % yt, r // yt
        step_n = r
        for first in {0, grid - 1, int(rng.integers(0, grid))}:
            q, r = divmod(first, cblocks)
            cb = r
            q, x = divmod(q, xt)
            nb, y = divmod(q, yt)
            item = first
            while item < total:
                want_q, want_cb = divmod(item, cblocks)
                want_q, want_x = divmod(want_q, xt)
                want_nb, want_y = divmod(want_q, yt)
                assert (cb, x, y, nb) == (want_cb, want_x, want_y, want_nb)
                cb += step_cb
                carry = cb >= cblocks
                cb -= cblocks if carry else 0
                x += step_x + carry
                carry = x >= xt
                x -= xt if carry else 0
                y += step_y + carry
                carry = y >= yt
                y -= yt if carry else 0
                nb += step_n + carry
                item += grid
