#!/bin/bash
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -x -q -m gpu > $O/r2h_pytest.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -4 $O/r2h_pytest.log
QNNP_CUDA_DW_S2_UMMA=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2h_bench_s2umma.json 2> $O/r2h_bench_s2umma.err; echo "bench s2umma exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2h_bench.json 2> $O/r2h_bench.err; echo "bench exit $?"; tail -3 $O/r2h_bench.err
python - <<'PY'
import json
a = json.load(open("gpurun_out/r2h_bench_s2umma.json")); b = json.load(open("gpurun_out/r2h_bench.json"))
print("s2umma ms/step", a["ms_per_step"], "default", b["ms_per_step"], "parity", (b.get("parity_check") or {}).get("mismatches"), "e2e", b["e2e"]["value"])
for la, lb in zip(a["layers"], b["layers"]):
    if la["kind"] == "dw":
        print("   %-12s s2umma %7.3f  default %7.3f  (%.0f GB/s)" % (la["layer"], la["ms"], lb["ms"], lb["gbs"]))
print({k: (round(v["ms_per_step"], 2), round(v["frac_of_hbm_peak"], 3)) for k, v in b["per_kernel"].items()})
ex = b.get("extras") or {}
print("tensor_bound", json.dumps(ex.get("tensor_bound_gemm"))[:900]); print("int8_peak", ex.get("int8_peak"), ex.get("int8_peak_sustained"))
PY
