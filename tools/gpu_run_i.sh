#!/bin/bash
# dw descriptor ring A/B: dw parity tests + short bench (layers table)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -x -q -m gpu > $O/r2i_pytest.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -4 $O/r2i_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $O/r2i_bench.json 2> $O/r2i_bench.err; echo "bench exit $?"; tail -3 $O/r2i_bench.err
python - <<'PY'
import json
b = json.load(open("gpurun_out/r2i_bench.json"))
print("ms/step", b["ms_per_step"], "parity", (b.get("parity_check") or {}).get("mismatches"), "e2e", b["e2e"]["value"])
for lb in b["layers"]:
    if lb["kind"] == "dw":
        print("   %-12s %7.3f  (%.0f GB/s)" % (lb["layer"], lb["ms"], lb["gbs"]))
print({k: (round(v["ms_per_step"], 2), round(v["frac_of_hbm_peak"], 3)) for k, v in b["per_kernel"].items()})
PY
