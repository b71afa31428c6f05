#!/bin/bash
# depthwise tile-shape sweep with the single (negated) weight operand
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2r_$tag.json 2> $O/r2r_$tag.err; echo "$tag exit $?"; }
run auto A=1
run mt4 QNNP_CUDA_DW_MT=4
run mt4g2 QNNP_CUDA_DW_MT=4 QNNP_CUDA_DW_G=2
run mt4g3 QNNP_CUDA_DW_MT=4 QNNP_CUDA_DW_G=3
run mt2 QNNP_CUDA_DW_MT=2
run mt8 QNNP_CUDA_DW_MT=8
python - <<'PY'
import json
tags = ("auto", "mt4", "mt4g2", "mt4g3", "mt2", "mt8")
r = {t: json.load(open("gpurun_out/r2r_%s.json" % t)) for t in tags}
print("%-10s" % "layer" + "".join("%9s" % t for t in tags))
for i, l in enumerate(r["auto"]["layers"]):
    if l["kind"] == "dw":
        print("%-10s" % l["layer"] + "".join("%9.3f" % r[t]["layers"][i]["ms"] for t in tags))
PY
