"""Shared case definitions for the parity tests.

Each case is a dict of qnnp_create_convolution2d_nhwc_q8 arguments plus input geometry.  The grids
restate, with fixed seeds, the reference's own test matrices (which draw from std::random_device):
  * OPERATOR_CASES  <- test/convolution.cc (1x1, grouped, 1x3/3x1/3x3 with each padding side, strides,
                       dilation, batch, input/output pixel strides, depthwise 3x3/5x5 variants)
  * DW_UKERNEL_CASES <- test/q8dwconv.cc Q8DWCONV_UP8x9 (channels ==8, %8, >8, subsampling 2,
                       input/output stride 17, qmin/qmax, zero-point-only variants)
  * GEMM_UKERNEL_CASES <- test/q8gemm.cc Q8GEMM_4x4c2 (k==8, k>8, k%8==0, strided a/c, qmin128,
                       qmax128, azp0/bzp0/nozp, sub-tiles) mapped onto the fully-connected operator
  * MOBILENET_V2     <- bench/convolution.cc:453-537
"""
from __future__ import annotations

import numpy as np


def conv_case(name, n, h, w, groups, gic, goc, ks=(1, 1), stride=(1, 1), dil=(1, 1), pad=(0, 0, 0, 0),
              izp=127, kzp=127, qmin=0, qmax=255, in_extra=0, out_extra=0, seed=None):
    return dict(name=name, n=n, h=h, w=w, groups=groups, gic=gic, goc=goc, ksize=ks, stride=stride, dilation=dil,
                pad=pad, izp=izp, kzp=kzp, qmin=qmin, qmax=qmax, in_extra=in_extra, out_extra=out_extra, seed=seed)


def _operator_cases():
    c = []
    a = c.append
    # test/convolution.cc: 1x1 family
    a(conv_case("1x1", 1, 27, 29, 1, 23, 19))
    a(conv_case("1x1_qmin", 1, 27, 29, 1, 23, 19, qmin=128))
    a(conv_case("1x1_qmax", 1, 27, 29, 1, 23, 19, qmax=128))
    a(conv_case("1x1_in_stride", 1, 27, 29, 1, 23, 19, in_extra=5))
    a(conv_case("1x1_out_stride", 1, 27, 29, 1, 23, 19, out_extra=7))
    a(conv_case("1x1_batch", 3, 13, 14, 1, 23, 19))
    a(conv_case("1x1_batch_strides", 3, 13, 14, 1, 23, 19, in_extra=5, out_extra=7))
    a(conv_case("1x1_k16_aligned", 2, 9, 9, 1, 16, 96))
    a(conv_case("1x1_k24", 2, 9, 9, 1, 24, 144))
    a(conv_case("1x1_k32_n16", 2, 9, 9, 1, 32, 16))
    a(conv_case("1x1_k144_n24", 2, 9, 9, 1, 144, 24))
    a(conv_case("1x1_wide_n", 1, 5, 5, 1, 64, 384))
    a(conv_case("1x1_big_k", 1, 5, 5, 1, 960, 320))
    a(conv_case("1x1_n1000", 1, 3, 3, 1, 1280, 1000))
    a(conv_case("xzp_like_k256", 1, 11, 9, 1, 256, 24, izp=93, kzp=201))
    a(conv_case("grouped_1x1", 1, 24, 25, 2, 17, 19))
    a(conv_case("grouped_1x1_strides", 2, 9, 8, 3, 5, 7, in_extra=3, out_extra=2))
    # 1x3 / 3x1 / 3x3
    a(conv_case("1x3", 1, 20, 19, 1, 17, 15, ks=(1, 3), pad=(0, 1, 0, 1)))
    a(conv_case("3x1", 1, 19, 20, 1, 17, 15, ks=(3, 1), pad=(1, 0, 1, 0)))
    a(conv_case("3x3", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_no_pad", 1, 13, 12, 1, 15, 17, ks=(3, 3)))
    a(conv_case("3x3_pad_left", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(0, 0, 0, 1)))
    a(conv_case("3x3_pad_right", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(0, 1, 0, 0)))
    a(conv_case("3x3_pad_top", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(1, 0, 0, 0)))
    a(conv_case("3x3_pad_bottom", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(0, 0, 1, 0)))
    a(conv_case("3x3_in_stride", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(1, 1, 1, 1), in_extra=5))
    a(conv_case("3x3_out_stride", 1, 13, 12, 1, 15, 17, ks=(3, 3), pad=(1, 1, 1, 1), out_extra=6))
    a(conv_case("3x3_batch", 3, 10, 9, 1, 15, 17, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_s2", 1, 13, 14, 1, 27, 19, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_s1x2", 1, 13, 14, 1, 27, 19, ks=(3, 3), stride=(1, 2), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_s2x1", 1, 13, 14, 1, 27, 19, ks=(3, 3), stride=(2, 1), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_d2", 1, 14, 13, 1, 27, 19, ks=(3, 3), dil=(2, 2), pad=(2, 2, 2, 2)))
    a(conv_case("3x3_d2x1", 1, 14, 13, 1, 27, 19, ks=(3, 3), dil=(2, 1), pad=(2, 1, 2, 1)))
    a(conv_case("3x3_aligned16", 2, 9, 9, 1, 16, 32, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("3x3_aligned32_s2", 1, 12, 12, 1, 32, 48, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1), izp=5, kzp=250))
    a(conv_case("stem_3x3_s2_c3", 2, 32, 32, 1, 3, 32, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1), izp=93, kzp=201))
    a(conv_case("grouped_3x3", 1, 10, 11, 2, 14, 13, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("5x5_big_k", 1, 9, 9, 1, 40, 24, ks=(5, 5), pad=(2, 2, 2, 2)))
    # depthwise (test/convolution.cc depthwise_*)
    a(conv_case("dw3x3", 1, 15, 14, 27, 1, 1, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("dw3x3_no_pad", 1, 15, 14, 27, 1, 1, ks=(3, 3)))
    a(conv_case("dw3x3_s2", 1, 15, 14, 27, 1, 1, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)))
    a(conv_case("dw3x3_s1x2", 1, 15, 14, 27, 1, 1, ks=(3, 3), stride=(1, 2), pad=(1, 1, 1, 1)))
    a(conv_case("dw3x3_s2x1", 1, 15, 14, 27, 1, 1, ks=(3, 3), stride=(2, 1), pad=(1, 1, 1, 1)))
    a(conv_case("dw3x3_d2", 1, 15, 14, 27, 1, 1, ks=(3, 3), dil=(2, 2), pad=(2, 2, 2, 2)))
    a(conv_case("dw3x3_d2x1", 1, 15, 14, 27, 1, 1, ks=(3, 3), dil=(2, 1), pad=(2, 1, 2, 1)))
    a(conv_case("dw3x3_batch", 3, 15, 14, 27, 1, 1, ks=(3, 3), pad=(1, 1, 1, 1)))
    a(conv_case("dw5x5", 1, 15, 14, 27, 1, 1, ks=(5, 5), pad=(2, 2, 2, 2)))
    a(conv_case("dw5x5_s2", 1, 15, 14, 27, 1, 1, ks=(5, 5), stride=(2, 2), pad=(2, 2, 2, 2)))
    a(conv_case("dw_multiplier", 1, 9, 9, 8, 1, 3, ks=(3, 3), pad=(1, 1, 1, 1)))
    return c


def _dw_ukernel_cases():
    c = []
    a = c.append
    base = dict(ks=(3, 3), pad=(1, 1, 1, 1))
    a(conv_case("dw_c8", 1, 5, 7, 8, 1, 1, **base))
    a(conv_case("dw_c8_s2", 1, 5, 9, 8, 1, 1, stride=(1, 2), **base))
    a(conv_case("dw_c8_in_stride17", 1, 5, 7, 8, 1, 1, in_extra=9, **base))
    a(conv_case("dw_c8_out_stride19", 1, 5, 7, 8, 1, 1, out_extra=11, **base))
    a(conv_case("dw_c8_qmin128", 1, 5, 7, 8, 1, 1, qmin=128, **base))
    a(conv_case("dw_c8_qmax128", 1, 5, 7, 8, 1, 1, qmax=128, **base))
    a(conv_case("dw_c8_izp_only", 1, 5, 7, 8, 1, 1, izp=255, kzp=0, **base))
    a(conv_case("dw_c8_kzp_only", 1, 5, 7, 8, 1, 1, izp=0, kzp=255, **base))
    a(conv_case("dw_c8_nozp", 1, 5, 7, 8, 1, 1, izp=0, kzp=0, **base))
    for ch in (16, 24, 64, 128):
        a(conv_case(f"dw_c{ch}", 1, 5, 6, ch, 1, 1, **base))
    for ch in (9, 10, 11, 13, 15, 33):
        a(conv_case(f"dw_c{ch}", 1, 5, 6, ch, 1, 1, **base))
    a(conv_case("dw_c12_strides", 2, 6, 5, 12, 1, 1, in_extra=4, out_extra=8, **base))
    a(conv_case("dw_c32_s2", 2, 12, 12, 32, 1, 1, stride=(2, 2), **base))
    a(conv_case("dw_c96_s2_odd", 1, 11, 13, 96, 1, 1, stride=(2, 2), **base))
    a(conv_case("dw_c144_w1", 1, 7, 1, 144, 1, 1, **base))
    return c


def fc_case(name, m, k, n, izp=127, kzp=127, qmin=0, qmax=255, in_extra=0, out_extra=0, seed=None):
    return dict(name=name, m=m, k=k, n=n, izp=izp, kzp=kzp, qmin=qmin, qmax=qmax, in_extra=in_extra,
                out_extra=out_extra, seed=seed)


def _gemm_ukernel_cases():
    c = []
    a = c.append
    # test/q8gemm.cc Q8GEMM_4x4c2__SSE2 grid (mr=4, nr=4) -> operator sizes
    a(fc_case("k_eq_8", 4, 8, 4))
    a(fc_case("k_eq_8_strided_a", 4, 8, 4, in_extra=29))
    a(fc_case("k_eq_8_strided_c", 4, 8, 4, out_extra=13))
    a(fc_case("k_eq_8_qmin128", 4, 8, 4, qmin=128))
    a(fc_case("k_eq_8_qmax128", 4, 8, 4, qmax=128))
    a(fc_case("k_eq_8_azp0", 4, 8, 4, izp=0))
    a(fc_case("k_eq_8_bzp0", 4, 8, 4, kzp=0))
    a(fc_case("k_eq_8_nozp", 4, 8, 4, izp=0, kzp=0))
    for k in range(9, 16):
        a(fc_case(f"k_gt_8_{k}", 4, k, 4))
        a(fc_case(f"k_gt_8_{k}_strided_a", 4, k, 4, in_extra=29))
        a(fc_case(f"k_gt_8_{k}_azp0", 4, k, 4, izp=0))
        a(fc_case(f"k_gt_8_{k}_bzp0", 4, k, 4, kzp=0))
    for k in (9, 12, 15):
        for m in range(1, 5):
            for n in range(1, 5):
                a(fc_case(f"k_gt_8_{k}_subtile_{m}x{n}", m, k, n))
    for k in range(16, 129, 24):
        a(fc_case(f"k_div_8_{k}", 4, k, 4))
        a(fc_case(f"k_div_8_{k}_strided_a", 4, k, 4, in_extra=35))
        a(fc_case(f"k_div_8_{k}_strided_c", 4, k, 4, out_extra=13))
    # operator-level (test/fully-connected.cc): unit batch / small batch, strides, qmin/qmax
    a(fc_case("fc_unit_batch", 1, 23, 19))
    a(fc_case("fc_small_batch", 12, 23, 19))
    a(fc_case("fc_small_batch_strides", 12, 23, 19, in_extra=5, out_extra=7))
    a(fc_case("fc_small_batch_qmin", 12, 23, 19, qmin=128))
    a(fc_case("fc_small_batch_qmax", 12, 23, 19, qmax=128))
    # config[0] of BASELINE.json and tile-boundary sizes of the tensor-core kernel
    a(fc_case("m64_n64_k64", 64, 64, 64))
    a(fc_case("m127", 127, 32, 16))
    a(fc_case("m128", 128, 32, 16))
    a(fc_case("m129", 129, 32, 16))
    a(fc_case("m300_k144_n24", 300, 144, 24))
    a(fc_case("m257_n240", 257, 64, 240))
    a(fc_case("m200_n241", 200, 64, 241))
    a(fc_case("m140_n1000_k1280", 140, 1280, 1000))
    a(fc_case("m512_k2048_n512", 512, 2048, 512, izp=3, kzp=250))
    return c


OPERATOR_CASES = _operator_cases()
DW_UKERNEL_CASES = _dw_ukernel_cases()
GEMM_UKERNEL_CASES = _gemm_ukernel_cases()

# bench/convolution.cc:453-537 — (H, W, KH, KW, stride, groups, gic, goc); padding = k/2 (:44-47)
MOBILENET_V2 = [
    ("stem", 224, 224, 3, 3, 2, 1, 3, 32),
    ("b1_dw", 112, 112, 3, 3, 1, 32, 1, 1), ("b1_pw", 112, 112, 1, 1, 1, 1, 32, 16),
    ("b2_exp", 112, 112, 1, 1, 1, 1, 16, 96), ("b2_dw", 112, 112, 3, 3, 2, 96, 1, 1), ("b2_pw", 56, 56, 1, 1, 1, 1, 96, 24),
    ("b3_exp", 56, 56, 1, 1, 1, 1, 24, 144), ("b3_dw", 56, 56, 3, 3, 1, 144, 1, 1), ("b3_pw", 56, 56, 1, 1, 1, 1, 144, 24),
    ("b4_dw", 56, 56, 3, 3, 2, 144, 1, 1), ("b4_pw", 28, 28, 1, 1, 1, 1, 144, 32),
    ("b5_exp", 28, 28, 1, 1, 1, 1, 32, 192), ("b5_dw", 28, 28, 3, 3, 1, 192, 1, 1), ("b5_pw", 28, 28, 1, 1, 1, 1, 192, 32),
    ("b7_dw", 28, 28, 3, 3, 2, 192, 1, 1), ("b7_pw", 14, 14, 1, 1, 1, 1, 192, 64),
    ("b8_exp", 14, 14, 1, 1, 1, 1, 64, 384), ("b8_dw", 14, 14, 3, 3, 1, 384, 1, 1), ("b8_pw", 14, 14, 1, 1, 1, 1, 384, 64),
    ("b11_pw", 14, 14, 1, 1, 1, 1, 384, 96),
    ("b12_exp", 14, 14, 1, 1, 1, 1, 96, 576), ("b12_dw", 14, 14, 3, 3, 1, 576, 1, 1), ("b12_pw", 14, 14, 1, 1, 1, 1, 576, 96),
    ("b14_dw", 14, 14, 3, 3, 2, 576, 1, 1), ("b14_pw", 7, 7, 1, 1, 1, 1, 576, 160),
    ("b15_exp", 7, 7, 1, 1, 1, 1, 160, 960), ("b15_dw", 7, 7, 3, 3, 1, 960, 1, 1), ("b15_pw", 7, 7, 1, 1, 1, 1, 960, 160),
    ("b17_pw", 7, 7, 1, 1, 1, 1, 960, 320), ("last", 7, 7, 1, 1, 1, 1, 320, 1280),
]


def mobilenet_case(entry, batch, seed=None):
    name, h, w, kh, kw, s, g, gic, goc = entry
    return conv_case(f"mnv2_{name}", batch, h, w, g, gic, goc, ks=(kh, kw), stride=(s, s), pad=(kh // 2, kw // 2, kh // 2, kw // 2),
                     seed=seed)


def make_conv_data(case, seed):
    """uint8 inputs/weights uniform [0,255], bias uniform [-10000,10000] (test/gemm-microkernel-tester.h:182-185)."""
    rng = np.random.default_rng(seed)
    cin = case["groups"] * case["gic"]
    x = rng.integers(0, 256, (case["n"], case["h"], case["w"], cin + case["in_extra"]), dtype=np.uint8)
    k = rng.integers(0, 256, (case["groups"], case["goc"], case["ksize"][0], case["ksize"][1], case["gic"]), dtype=np.uint8)
    b = rng.integers(-10000, 10001, (case["groups"] * case["goc"],), dtype=np.int32)
    return x, k, b


def make_fc_data(case, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(0, 256, (case["m"], case["k"] + case["in_extra"]), dtype=np.uint8)
    k = rng.integers(0, 256, (case["n"], case["k"]), dtype=np.uint8)
    b = rng.integers(-10000, 10001, (case["n"],), dtype=np.int32)
    return x, k, b


def derive_output_quant(acc):
    """Output scale / zero point derived from the accumulator range so that all 256 codes occur
    (test/convolution-operator-tester.h:407-413, test/gemm-microkernel-tester.h:236-241).  The
    requantization scale is 1/output_scale with unit input and kernel scales."""
    amin, amax = int(acc.min()), int(acc.max())
    oscale = float(np.float32(max((amax - amin) / 255.0, 1.0 + 2.0 ** -20)))
    ozp = int(np.clip(round(127.5 - (amax + amin) / 2.0 / oscale), 0, 255))
    return oscale, ozp


# ---- round-1 additions: shapes that select specific device paths.  The GPU tests compare these with the oracle; the CPU
# suite pins the oracle to the compiled reference on the same cases (tests/test_oracle.py). -------------------------------
# depthwise shapes that take the tcgen05 path (channels % 16 == 0, dense pixels): geometry classes of its planner —
# 16-row tiles vs whole images stacked, one vs two parity planes, ragged tiles, every weight-operand mode, clamps
DW_TC = dict(ks=(3, 3), pad=(1, 1, 1, 1))
DW_TC_CASES = [
    conv_case("tc_c16_rows", 1, 20, 23, 16, 1, 1, **DW_TC),
    conv_case("tc_c48_7x7_stack2", 3, 7, 7, 48, 1, 1, **DW_TC),
    conv_case("tc_c32_14x14", 3, 14, 14, 32, 1, 1, **DW_TC),
    conv_case("tc_c32_s2_rows", 1, 40, 36, 32, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c16_s2_whole", 3, 14, 14, 16, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c32_nopad", 2, 9, 12, 32, 1, 1, ks=(3, 3)),
    conv_case("tc_c16_s2_nopad", 1, 18, 16, 16, 1, 1, ks=(3, 3), stride=(2, 2)),
    conv_case("tc_c16_wide", 2, 5, 70, 16, 1, 1, **DW_TC),
    conv_case("tc_c64_kzp128_s8", 2, 17, 17, 64, 1, 1, kzp=128, **DW_TC),      # w - kzp fits s8: one operand
    conv_case("tc_c64_kzp0_u8", 2, 17, 17, 64, 1, 1, kzp=0, izp=3, **DW_TC),    # u8 weights
    conv_case("tc_c32_kzp255", 1, 12, 12, 32, 1, 1, kzp=255, izp=255, **DW_TC),
    conv_case("tc_c32_clamp", 1, 12, 12, 32, 1, 1, qmin=40, qmax=200, **DW_TC),
    conv_case("tc_c160_112", 1, 112, 112, 160, 1, 1, **DW_TC),
    conv_case("tc_c32_out_stride", 1, 10, 10, 32, 1, 1, out_extra=16, **DW_TC),
    conv_case("tc_c32_in_stride", 1, 10, 10, 32, 1, 1, in_extra=16, **DW_TC),
    # whole-image mode where a second stacked image would not fit its rows into the 16 row groups (found by the CPU replay)
    conv_case("tc_c32_s2_28_b3", 3, 28, 28, 32, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c16_s2_18x6_nopad_b3", 3, 18, 6, 16, 1, 1, ks=(3, 3), stride=(2, 2)),
    conv_case("tc_c32_12x12_b5", 5, 12, 12, 32, 1, 1, **DW_TC),
    # round 2: kzp = 127 (the default above) now takes the single NEGATED operand (kzp - w fits s8); these keep the
    # two-operand form (32 accumulator columns per unit) covered on every geometry class, and the plain single operand at 128
    conv_case("tc_c32_rows_kzp60_two_operands", 1, 20, 23, 32, 1, 1, kzp=60, **DW_TC),
    conv_case("tc_c48_7x7_stack2_kzp200_two_operands", 3, 7, 7, 48, 1, 1, kzp=200, **DW_TC),
    conv_case("tc_c32_s2_rows_kzp90_two_operands", 1, 40, 36, 32, 1, 1, stride=(2, 2), kzp=90, **DW_TC),
    conv_case("tc_c64_56_kzp1_two_operands", 2, 56, 56, 64, 1, 1, kzp=1, **DW_TC),
    conv_case("tc_c32_s2_rows_kzp128", 1, 40, 36, 32, 1, 1, stride=(2, 2), kzp=128, **DW_TC),
    conv_case("tc_c144_56_b2_negated", 2, 56, 56, 144, 1, 1, **DW_TC),
    # channel-pair form (32-channel TMA boxes): odd group counts (last pair half empty), stride 2 with both parity planes,
    # stacked whole images, many pairs per item
    conv_case("tc_c48_s2_rows_odd_groups", 2, 40, 36, 48, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c144_s2_28_b3", 3, 28, 28, 144, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c96_s2_112", 1, 112, 112, 96, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c192_28_b3", 3, 28, 28, 192, 1, 1, **DW_TC),
    conv_case("tc_c384_14_b5", 5, 14, 14, 384, 1, 1, **DW_TC),
    conv_case("tc_c576_s2_14_b4", 4, 14, 14, 576, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("tc_c960_7_b6_streamed_weights", 6, 7, 7, 960, 1, 1, **DW_TC),
]

STEM_CASES = [
    conv_case("stem_72_b5", 5, 72, 72, 1, 3, 32, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),
    conv_case("stem_70x74_b4", 4, 70, 74, 1, 3, 24, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),
    conv_case("stem_s1_40_b4", 4, 40, 40, 1, 3, 16, ks=(3, 3), pad=(1, 1, 1, 1)),
    conv_case("stem_nopad_66_b4", 4, 66, 66, 1, 3, 32, ks=(3, 3), stride=(2, 2)),
    # images too small for the raw-row ring (an item would span more than two): per-thread global gather instead
    conv_case("stem_small_20_b3", 3, 20, 20, 1, 3, 32, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),
    conv_case("stem_small_s1_9x11_b5", 5, 9, 11, 1, 3, 16, ks=(3, 3), pad=(1, 1, 1, 1)),
]

PERSISTENT_CASES = [
    conv_case("pers_1x1_expand", 2, 40, 40, 1, 16, 96),                      # folded, TMA, bulk stores, mt = 2
    conv_case("pers_1x1_project", 2, 40, 40, 1, 96, 24),                     # ones mode, several k-chunks
    conv_case("pers_1x1_n144", 1, 36, 36, 1, 24, 144),                       # 16-column remainder units
    conv_case("pers_1x1_wide", 1, 14, 14, 1, 320, 1280),                     # streamed weights, 5 n-tiles
    conv_case("pers_3x3_conv", 1, 30, 30, 1, 16, 32, ks=(3, 3), pad=(1, 1, 1, 1)),   # cp.async conv loader
    conv_case("pers_stem", 3, 72, 72, 1, 3, 32, ks=(3, 3), stride=(2, 2), pad=(1, 1, 1, 1)),  # raw-row ring
    conv_case("pers_dw_s1", 5, 40, 40, 64, 1, 1, **DW_TC),                   # tcgen05 depthwise, row tiles
    conv_case("pers_dw_s2", 5, 40, 40, 64, 1, 1, stride=(2, 2), **DW_TC),
    conv_case("pers_dw_stacked", 9, 7, 7, 96, 1, 1, **DW_TC),                # stacked images, odd batch
    conv_case("pers_dw_c20", 3, 30, 30, 20, 1, 1, **DW_TC),                  # streaming dp4a kernel
]

