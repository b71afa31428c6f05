#!/bin/bash
# CTA-pair GEMM vs the single-CTA kernel with 32-byte activation slabs on the deep-K projections
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2y_$tag.json 2> $O/r2y_$tag.err; echo "$tag exit $?"; }
run default A=1
run nog2 QNNP_CUDA_NO_GEMM2SM=1
run g2k256 QNNP_CUDA_GEMM2SM_MIN_K=256
python - <<'PY'
import json
tags = ("default", "nog2", "g2k256")
r = {t: json.load(open("gpurun_out/r2y_%s.json" % t)) for t in tags}
print({t: r[t]["ms_per_step"] for t in tags})
for i, l in enumerate(r["default"]["layers"]):
    ms = [r[t]["layers"][i]["ms"] for t in tags]
    if max(ms) - min(ms) > 0.004:
        print("%-12s" % l["layer"] + "".join("%9.3f" % m for m in ms))
PY
