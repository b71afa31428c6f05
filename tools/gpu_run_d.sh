#!/bin/bash
# Round-2 GPU session D: full GPU suite + bench (with side measurements and the real-network run) on the current code.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
T=${1:-r2d}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/${T}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench exit $?"; tail -3 $O/${T}_bench.err
python - $T <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json"))
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "parity", (d.get("parity_check") or {}).get("mismatches"), "e2e", d.get("e2e") and round(d["e2e"]["value"]))
print({k: (round(v["ms_per_step"], 2), round(v["frac_of_hbm_peak"], 3)) for k, v in d["per_kernel"].items()})
ex = d.get("extras") or {}
print("tensor_bound", ex.get("tensor_bound_gemm"))
print("full_network", json.dumps(ex.get("full_network"))[:1200])
print("latency", ex.get("small_batch_latency"))
print("err", ex.get("error"))
for l in d["layers"]:
    print("   %-14s %-5s %7.3f ms %7.0f GB/s" % (l["layer"], l["kind"], l["ms"], l["gbs"]))
PY
