#!/bin/bash
# depthwise tile-shape sweep on the final forms (looks at 56x56x144 in particular)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2x_$tag.json 2> $O/r2x_$tag.err; echo "$tag exit $?"; }
run auto A=1
run mt2g6 QNNP_CUDA_DW_MT=2 QNNP_CUDA_DW_G=6
run mt2g8 QNNP_CUDA_DW_MT=2
run mt3 QNNP_CUDA_DW_MT=3
run mt7 QNNP_CUDA_DW_MT=8 QNNP_CUDA_DW_G=2
python - <<'PY'
import json
tags = ("auto", "mt2g6", "mt2g8", "mt3", "mt7")
r = {t: json.load(open("gpurun_out/r2x_%s.json" % t)) for t in tags}
print("%-10s" % "layer" + "".join("%9s" % t for t in tags))
for i, l in enumerate(r["auto"]["layers"]):
    if l["kind"] == "dw":
        print("%-10s" % l["layer"] + "".join("%9.3f" % r[t]["layers"][i]["ms"] for t in tags))
PY
