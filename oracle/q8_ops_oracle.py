"""TEST INFRASTRUCTURE — NumPy restatement of the operators beside the convolution path, pinned to the compiled
reference by tests/test_ops_oracle.py (CPU) and used as a second checker for the GPU kernels.  Never imported by the
product.  Formulas: reference src/qnnpack/requantization.h:200-265 (avgpool params), :327-414 (add params),
:482-498 qnnp_avgpool_quantize, :500-522 qnnp_add_quantize; src/indirection.c:134-260 (deconvolution / max-pooling taps)."""
from __future__ import annotations

import numpy as np


def _f32(x):
    return np.float32(x)


def _bits(f):
    return int(np.array([f], np.float32).view(np.uint32)[0])


def add_params(a_zp, a_scale, b_zp, b_scale, y_zp, y_scale, y_min, y_max):
    a_os, b_os = _f32(a_scale) / _f32(y_scale), _f32(b_scale) / _f32(y_scale)
    mx = max(a_os, b_os)
    exp = (_bits(mx) >> 23) - 127
    shift = 21 - exp
    mult = np.array([(21 - exp + 127) << 23], np.uint32).view(np.float32)[0]
    am = int(np.rint(_f32(a_os * mult)))  # lrintf: round half to even, like np.rint
    bm = int(np.rint(_f32(b_os * mult)))
    zpp = (-(am * a_zp + bm * b_zp)) & 0xFFFFFFFF
    return dict(am=am, bm=bm, shift=shift, zpp=zpp, y_zp=y_zp, y_min=y_min, y_max=y_max)


def add(a, b, p):
    acc = (p["zpp"] + a.astype(np.int64) * p["am"] + b.astype(np.int64) * p["bm"]) & 0xFFFFFFFF
    acc = np.where(acc >= 2 ** 31, acc - 2 ** 32, acc)  # int32
    mask = (1 << p["shift"]) - 1
    rem = (acc & mask) - (acc < 0)
    acc = (acc >> p["shift"]) + (rem > (mask >> 1))
    return np.clip(acc + p["y_zp"], p["y_min"], p["y_max"]).astype(np.uint8)


def avg_quant(scale, ozp, omin, omax):
    bits = _bits(_f32(scale))
    return dict(mult=(bits & 0x007FFFFF) | 0x00800000, shift=127 + 23 - (bits >> 23), ozp=ozp, lo=omin - ozp, hi=omax - ozp)


def avg_quantize(n, q):
    n = n.astype(np.int64)
    n = np.where(n >= 2 ** 31, n - 2 ** 32, n)
    adj = n * q["mult"] - (n < 0)
    y = (adj + (1 << (q["shift"] - 1))) >> q["shift"]
    return (np.clip(y, q["lo"], q["hi"]) + q["ozp"]).astype(np.uint8)


def global_average_pooling(x, izp, in_scale, ozp, out_scale, omin, omax):
    """x: [batch, width, channels]"""
    width = x.shape[1]
    q = avg_quant(_f32(in_scale) / (_f32(out_scale) * _f32(width)), ozp, omin, omax)
    return avg_quantize(x.astype(np.int64).sum(1) - width * izp, q)


def pool2d(x, kind, pad, pool, stride, dil, izp=0, in_scale=1.0, ozp=0, out_scale=1.0, omin=0, omax=255):
    """x: [n, h, w, c]; pad = (top, right, bottom, left)."""
    n, h, w, c = x.shape
    d = dil if kind == "max" else (1, 1)
    oh = (pad[0] + h + pad[2] - ((pool[0] - 1) * d[0] + 1)) // stride[0] + 1
    ow = (pad[3] + w + pad[1] - ((pool[1] - 1) * d[1] + 1)) // stride[1] + 1
    out = np.zeros((n, oh, ow, c), np.int64)
    for oy in range(oh):
        for ox in range(ow):
            acc = np.zeros((n, c), np.int64)
            for ky in range(pool[0]):
                for kx in range(pool[1]):
                    iy, ix = oy * stride[0] + ky * d[0] - pad[0], ox * stride[1] + kx * d[1] - pad[3]
                    if kind == "max":  # clamped to the edge (src/indirection.c:218-224)
                        iy, ix = min(max(iy, 0), h - 1), min(max(ix, 0), w - 1)
                        acc = np.maximum(acc, x[:, iy, ix, :])
                    elif 0 <= iy < h and 0 <= ix < w:  # padded taps read izp: (izp - izp) = 0
                        acc += x[:, iy, ix, :].astype(np.int64) - izp
            out[:, oy, ox, :] = acc
    if kind == "max":
        return np.clip(out, omin, omax).astype(np.uint8)
    q = avg_quant(_f32(in_scale) / (_f32(out_scale) * _f32(pool[0] * pool[1])), ozp, omin, omax)
    return avg_quantize(out, q)


def deconv_accumulators(x, k, b, pad, adj, ks, stride, dil, groups, gic, goc, izp, kzp):
    """x: [n, h, w, groups*gic], k: [groups, gic, kh, kw, goc] (test/deconvolution-operator-tester.h:411)
    -> int64 [n, oh, ow, groups*goc]; tap mapping src/indirection.c:134-190."""
    n, h, w, _ = x.shape
    oh = stride[0] * (h - 1) + adj[0] + (ks[0] - 1) * dil[0] + 1 - (pad[0] + pad[2])
    ow = stride[1] * (w - 1) + adj[1] + (ks[1] - 1) * dil[1] + 1 - (pad[1] + pad[3])
    acc = np.zeros((n, oh, ow, groups * goc), np.int64) + b.astype(np.int64)
    xs = x.astype(np.int64) - izp
    kk = k.astype(np.int64) - kzp
    for oy in range(oh):
        for ox in range(ow):
            for ky in range(ks[0]):
                yy = oy + pad[0] - ky * dil[0]
                if yy < 0 or yy % stride[0] or yy // stride[0] >= h:
                    continue
                for kx in range(ks[1]):
                    xx = ox + pad[3] - kx * dil[1]
                    if xx < 0 or xx % stride[1] or xx // stride[1] >= w:
                        continue
                    px = xs[:, yy // stride[0], xx // stride[1], :].reshape(n, groups, gic)
                    acc[:, oy, ox, :] += np.einsum("ngc,gco->ngo", px, kk[:, :, ky, kx, :]).reshape(n, groups * goc)
    return acc
