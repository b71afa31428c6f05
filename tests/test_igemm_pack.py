"""CPU replay of the tensor-core kernel's operand algebra on the bytes the host really packs
(qnnp_cuda_debug_pack_igemm = pack_igemm_host of qnnpack_b200/csrc/qnnpack_api.cu, no GPU needed).

DESIGN.md §2 claims that the accumulator leaving TMEM is the reference accumulator
    acc[n] = bias[n] + sum_k (a[k] - izp) * (w[n][k] - kzp)                       (src/qnnpack/pack.h:24-43 folded into bias')
in both regroupings:
  folded: u8 x s8 UMMAs with B1 = w XOR 0x80 (raw u8 x u8 when kzp == 0), a constant (128 - kzp) operand whose "tail"
          copy is zero in the K padding, and bias' as signed base-255 digits against the constant A row [255 x31, 1];
  ones  : raw u8 x u8 UMMA + a row of ones (row sum of A), epilogue adds bias' - kzp * rowsum.
This test plays the UMMAs in NumPy on the packed blob — canonical K-major layout [chunk][row][16 B], K = 32 per step,
signedness per instruction descriptor — with GARBAGE in A's K padding (the kernel relies on zero weights there), and
compares with the formula, over random shapes, zero points and bias magnitudes (1..4 bias steps and the fallback)."""
import ctypes as C

import numpy as np
import pytest

META = "folded nkc n_tiles n_tile n_mma blk_chunks bias_steps b_signed has_b2 k_tail_pad has_corr".split()


@pytest.fixture(scope="module")
def pack():
    from qnnpack_b200 import build
    lib = C.CDLL(build.build())
    lib.qnnp_cuda_debug_pack_igemm.argtypes = [C.c_size_t, C.c_size_t, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p,
                                               C.POINTER(C.c_int), C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p,
                                               C.POINTER(C.c_size_t)]

    def f(kernel, bias, izp, kzp):
        n, k = kernel.shape
        meta = (C.c_int * 16)()
        blob = np.zeros(8 << 20, dtype=np.uint8)
        fb = np.zeros(1 << 16, dtype=np.int32)
        nb, nf = C.c_size_t(blob.size), C.c_size_t(fb.size)
        kernel = np.ascontiguousarray(kernel)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        ok = lib.qnnp_cuda_debug_pack_igemm(k, n, izp, kzp, kernel.ctypes.data, bias.ctypes.data, meta, blob.ctypes.data,
                                            C.byref(nb), fb.ctypes.data, C.byref(nf))
        assert ok == 1
        return dict(zip(META, meta)), blob[:nb.value].copy(), fb[:nf.value].copy()
    return f


def replay(meta, blob, fbias, a_rows, n_out, kzp):
    """a_rows: [M][nkc*16] uint8 (padding bytes arbitrary) -> accumulators [M][n_out] as the epilogue sees them."""
    nkc, n_tile, n_mma, blk = meta["nkc"], meta["n_tile"], meta["n_mma"], meta["blk_chunks"]
    a = a_rows.astype(np.int64)
    out = np.zeros((a.shape[0], n_out), dtype=np.int64)
    for nt in range(meta["n_tiles"]):
        b = blob[nt * blk * n_mma * 16:(nt + 1) * blk * n_mma * 16].reshape(blk, n_mma, 16)   # [chunk][row][byte]
        u8 = b.astype(np.int64)
        s8 = b.view(np.int8).astype(np.int64)

        def kmat(chunk0, chunks, signed):  # -> [rows][chunks*16] operand as the UMMA reads it
            src = s8 if signed else u8
            return src[chunk0:chunk0 + chunks].transpose(1, 0, 2).reshape(n_mma, chunks * 16)

        if meta["folded"]:
            acc = np.zeros((a.shape[0], n_mma), dtype=np.int64)
            a_const = np.array([255] * 31 + [1], dtype=np.int64)
            for t in range(meta["bias_steps"]):                       # accumulator := bias' (u8 x s8)
                acc += (a_const @ kmat(nkc + 4 + 2 * t, 2, True).T)[None, :]
            acc += a @ kmat(0, nkc, bool(meta["b_signed"])).T         # main operand
            if meta["has_b2"]:
                for c in range(0, nkc, 2):                            # (128 - kzp) * sum_k a, per K = 32 step
                    tail = meta["k_tail_pad"] and c + 2 == nkc
                    acc += a[:, c * 16:(c + 2) * 16] @ kmat(nkc + (2 if tail else 0), 2, True).T
            vals = acc[:, :n_tile]
        else:
            raw = a @ kmat(0, nkc, False).T                           # u8 x u8, row n_tile = ones
            rowsum = raw[:, n_tile]
            vals = raw[:, :n_tile] + fbias[nt * n_tile:(nt + 1) * n_tile].astype(np.int64)[None, :] - kzp * rowsum[:, None]
        lo, hi = nt * n_tile, min((nt + 1) * n_tile, n_out)
        out[:, lo:hi] = vals[:, :hi - lo]
    return out


def _cases():
    rng = np.random.default_rng(77)
    out = []
    for i in range(60):
        k = int(rng.choice([1, 3, 15, 16, 17, 24, 27, 31, 32, 33, 64, 96, 100, 144, 320]))
        n = int(rng.choice([1, 8, 16, 24, 96, 144, 200, 241, 300, 500]))
        kzp = int(rng.choice([0, 1, 77, 127, 128, 129, 255]))
        izp = int(rng.integers(0, 256))
        bmag = int(rng.choice([100, 10_000, 900_000, 40_000_000, 2_000_000_000]))
        mode = ["auto", "ones", "folded"][i % 3]
        out.append((k, n, izp, kzp, bmag, mode))
    return out


@pytest.mark.parametrize("k,n,izp,kzp,bmag,mode", _cases())
def test_packed_operands_reproduce_the_reference_accumulator(pack, monkeypatch, k, n, izp, kzp, bmag, mode):
    if mode == "auto":
        monkeypatch.delenv("QNNP_CUDA_IGEMM_MODE", raising=False)
    else:
        monkeypatch.setenv("QNNP_CUDA_IGEMM_MODE", mode)
    rng = np.random.default_rng(k * 1000 + n)
    w = rng.integers(0, 256, (n, k), dtype=np.uint8)
    bias = rng.integers(-bmag, bmag + 1, n).astype(np.int32)
    meta, blob, fbias = pack(w, bias, izp, kzp)
    assert meta["nkc"] % 2 == 0 and meta["nkc"] * 16 >= k and meta["n_tiles"] * meta["n_tile"] >= n
    if mode == "ones":
        assert not meta["folded"]
    m = 37
    a = rng.integers(0, 256, (m, meta["nkc"] * 16), dtype=np.uint8)      # padding bytes are garbage on purpose
    a[0, :k] = 255                                                       # extremes
    a[1, :k] = 0
    got = replay(meta, blob, fbias, a, n, kzp)
    want = bias.astype(np.int64)[None, :] + (a[:, :k].astype(np.int64) - izp) @ (w.astype(np.int64) - kzp).T
    assert np.array_equal(got, want), meta
    assert np.abs(got).max() < 2 ** 31                                    # and it fits the int32 accumulator


def test_bias_steps_cover_the_stated_range_and_fall_back(pack, monkeypatch):
    monkeypatch.delenv("QNNP_CUDA_IGEMM_MODE", raising=False)
    w = np.full((96, 16), 127, dtype=np.uint8)                            # N > K: folded is the default
    seen = set()
    for b in (0, 900_000, 1_500_000, 2_500_000, 3_900_000, 5_000_000, 100_000_000):   # one step covers |bias'| <= 1.0 M
        meta, _, _ = pack(w, np.full(96, b, dtype=np.int32), 127, 127)
        seen.add((meta["folded"], meta["bias_steps"]))
    assert seen == {(1, 1), (1, 2), (1, 3), (1, 4), (0, 0)}   # 1..4 bias steps on the tensor core, then "ones" mode
