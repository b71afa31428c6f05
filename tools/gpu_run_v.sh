#!/bin/bash
# depthwise A/B in one session: resident weights on/off, pair form forced on for all strides
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2v_$tag.json 2> $O/r2v_$tag.err; echo "$tag exit $?"; }
run default A=1
run bstream QNNP_CUDA_DW_B_STREAM=1
run pairall QNNP_CUDA_DW_PAIR=1
run default2 A=1
python - <<'PY'
import json
tags = ("default", "bstream", "pairall", "default2")
r = {t: json.load(open("gpurun_out/r2v_%s.json" % t)) for t in tags}
print("%-10s" % "layer" + "".join("%10s" % t for t in tags))
for i, l in enumerate(r["default"]["layers"]):
    if l["kind"] == "dw":
        print("%-10s" % l["layer"] + "".join("%10.3f" % r[t]["layers"][i]["ms"] for t in tags))
PY
