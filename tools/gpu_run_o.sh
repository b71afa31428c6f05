#!/bin/bash
# stem A/B: interior fast path on/off, sub-tiles per item
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2o_$tag.json 2> $O/r2o_$tag.err; echo "$tag exit $?"; }
run base A=1
run nofast QNNP_LIB_PATH=$PWD/qnnpack_b200/lib/libqnnpack_nofast.so
run mt4 QNNP_CUDA_MAX_SUBTILES=4
run base2 A=1
python - <<'PY'
import json
for t in ("base", "nofast", "mt4", "base2"):
    b = json.load(open("gpurun_out/r2o_%s.json" % t))
    print(t, b["ms_per_step"], [(l["layer"], round(l["ms"], 3)) for l in b["layers"] if l["kind"] != "dw"][:5])
PY
