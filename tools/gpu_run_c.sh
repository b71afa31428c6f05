#!/bin/bash
# Round-2 GPU session C: full GPU test suite on the current code, bench with side measurements (int8 peak, tensor-bound
# GEMM, small-batch latency, host-pointer path), ncu of the depthwise kernel after the contiguous-chunk schedule,
# compute-sanitizer memcheck / racecheck on the persistent-loop, tensor-core-depthwise and stem cases.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $O/r2c_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_umma -c 2 -o $O/r2c_dw_umma_first2 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2c_ncu_dw.log 2>&1; echo "ncu exit $?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "many_items or tensor_core or stem" > $O/r2c_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 $O/r2c_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "many_items and (pers_1x1_expand or pers_dw_s1 or pers_stem)" > $O/r2c_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 $O/r2c_sanitizer_racecheck.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "parity", (d.get("parity_check") or {}).get("mismatches"), "e2e", d.get("e2e") and round(d["e2e"]["value"]))
print(json.dumps(d.get("extras"), indent=1)[:3000])
for l in d["layers"]:
    print("   %-14s %-5s %7.3f ms %7.0f GB/s" % (l["layer"], l["kind"], l["ms"], l["gbs"]))
PY
