#!/bin/bash
# stem sensitivity: sub-tiles per item, folded vs ones mode
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2n_$tag.json 2> $O/r2n_$tag.err; echo "$tag exit $?"; }
run base A=1
run mt4 QNNP_CUDA_MAX_SUBTILES=4
run mt2 QNNP_CUDA_MAX_SUBTILES=2
run ones QNNP_CUDA_IGEMM_MODE=ones
python - <<'PY'
import json
for t in ("base", "mt4", "mt2", "ones"):
    b = json.load(open("gpurun_out/r2n_%s.json" % t))
    print(t, b["ms_per_step"], [(l["layer"], round(l["ms"], 3)) for l in b["layers"] if l["kind"] != "dw"][:5])
PY
