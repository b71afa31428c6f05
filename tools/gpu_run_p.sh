#!/bin/bash
# full validation of the round-2 code: smoke(), every GPU test, the default bench (with extras) and the reference arm
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2p_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r2p_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2p_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2p_pytest.log
timeout 900 python bench.py > $O/r2p_bench.json 2> $O/r2p_bench.err; echo "bench exit $?"; tail -3 $O/r2p_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2p_bench_ref.json 2> $O/r2p_bench_ref.err; echo "ref exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2p_bench.json"))
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "parity", (d.get("parity_check") or {}).get("mismatches"), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
print("roofline", d.get("roofline")); print("clocks", d.get("clocks")); print("cpu_baseline", json.dumps(d.get("cpu_baseline"))[:600])
print(json.dumps(d.get("extras"), indent=1)[:3500])
print({k: (round(v["ms_per_step"], 2), round(v["frac_of_hbm_peak"], 3)) for k, v in d["per_kernel"].items()})
r = json.load(open("gpurun_out/r2p_bench_ref.json")); print("reference arm", {k: r.get(k) for k in ("impl", "value", "unit", "ms_per_step", "cpu_baseline")})
PY
