#!/bin/bash
# experiment: depthwise tensor-core kernel without its output stores (timing only): how much of its time is the write path?
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2u_$tag.json 2> $O/r2u_$tag.err; echo "$tag exit $?"; }
run base A=1
run nostore QNNP_LIB_PATH=$PWD/qnnpack_b200/lib/libqnnpack_nostore.so
python - <<'PY'
import json
tags = ("base", "nostore")
r = {t: json.load(open("gpurun_out/r2u_%s.json" % t)) for t in tags}
for i, l in enumerate(r["base"]["layers"]):
    if l["kind"] == "dw":
        print("%-10s" % l["layer"] + "".join("%9.3f" % r[t]["layers"][i]["ms"] for t in tags))
PY
