#!/bin/bash
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm2sm.py tests/test_gpu_ops.py -x -q > $O/r2g_pytest.log 2>&1; rc=$?; echo "pytest exit $rc"; tail -5 $O/r2g_pytest.log
for shape in "65536 4096 4096" "16384 8192 8192" "32768 2048 8192" "65536 1024 1024"; do
  timeout 120 python tools/gemm_big.py $shape 2>&1 | tail -1
done | tee $O/r2g_gemm.jsonl
QNNP_CUDA_GEMM2SM_MIN_K=512 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2g_bench_mink512.json 2> $O/r2g_bench_mink512.err; echo "bench mink512 exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2g_bench.json 2> $O/r2g_bench.err; echo "bench exit $?"; tail -3 $O/r2g_bench.err
python - <<'PY'
import json
a = json.load(open("gpurun_out/r2g_bench_mink512.json")); b = json.load(open("gpurun_out/r2g_bench.json"))
print("mink512 ms/step", a["ms_per_step"], "default", b["ms_per_step"], "parity", (b.get("parity_check") or {}).get("mismatches"))
for la, lb in zip(a["layers"], b["layers"]):
    if abs(la["ms"] - lb["ms"]) > 0.004:
        print("   %-12s mink512 %7.3f  default %7.3f" % (la["layer"], la["ms"], lb["ms"]))
ex = b.get("extras") or {}
print("tensor_bound", ex.get("tensor_bound_gemm")); print("int8_peak", ex.get("int8_peak"))
fn = ex.get("full_network") or {}
print("full_network", fn.get("ms_per_step"), fn.get("images_per_s"), {k: round(v["frac_of_hbm_peak"], 3) for k, v in (fn.get("by_kind") or {}).items()}, (fn.get("parity_check") or {}).get("mismatches"))
PY
