"""Generates tests/golden/q8_golden.npz from the UNMODIFIED compiled reference (oracle/_ref).

Run where /root/reference exists:   make -C oracle ref && python tests/golden/make_golden.py
For every case of tests/cases.py the reference's own create -> setup -> run produces the expected
uint8 output (with 0xA5 canaries in the pixel-stride gaps).  Small outputs are stored verbatim,
MobileNetV2 batch-1 layer outputs as SHA-256 digests.  Inputs are regenerated from the case seed;
their digest is stored too, so a drifting RNG fails loudly instead of silently.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref as R  # noqa: E402
from tests import cases as CS, util as U  # noqa: E402


def main():
    rf = R.QnnpackHost()
    out = {}
    for case in CS.OPERATOR_CASES + CS.DW_UKERNEL_CASES:
        x, k, b, kw = U.conv_setup(case)
        y = U.run_conv(rf, case, x, k, b, kw)
        out[f"conv/{case['name']}/y"] = y
        out[f"conv/{case['name']}/in_digest"] = np.array(U.digest(x) + U.digest(k) + U.digest(b))
    for case in CS.GEMM_UKERNEL_CASES:
        x, k, b, kw = U.fc_setup(case)
        y = U.run_fc(rf, case, x, k, b, kw)
        out[f"fc/{case['name']}/y"] = y
        out[f"fc/{case['name']}/in_digest"] = np.array(U.digest(x) + U.digest(k) + U.digest(b))
    for entry in CS.MOBILENET_V2:
        case = CS.mobilenet_case(entry, 1)
        x, k, b, kw = U.conv_setup(case)
        y = U.run_conv(rf, case, x, k, b, kw)
        out[f"mnv2/{case['name']}/y_digest"] = np.array(U.digest(y))
        out[f"mnv2/{case['name']}/in_digest"] = np.array(U.digest(x) + U.digest(k) + U.digest(b))
    np.savez_compressed(U.GOLDEN, **out)
    print("wrote", U.GOLDEN, os.path.getsize(U.GOLDEN), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
