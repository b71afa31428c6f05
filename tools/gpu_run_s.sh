#!/bin/bash
# stride-2 depthwise: tensor-core kernel (single negated operand) vs streaming kernel
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/r2s_$tag.json 2> $O/r2s_$tag.err; echo "$tag exit $?"; }
run base A=1
run s2umma QNNP_CUDA_DW_S2_UMMA=1
python - <<'PY'
import json
tags = ("base", "s2umma")
r = {t: json.load(open("gpurun_out/r2s_%s.json" % t)) for t in tags}
print({t: (r[t]["ms_per_step"], (r[t].get("parity_check") or {}).get("mismatches")) for t in tags})
for i, l in enumerate(r["base"]["layers"]):
    if l["kind"] == "dw":
        print("%-10s" % l["layer"] + "".join("%9.3f" % r[t]["layers"][i]["ms"] for t in tags))
PY
