#!/bin/bash
# final round-2 evidence: ncu of the first tensor-core GEMM and depthwise launches, launch list, sanitizer on the new paths
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:q8_igemm -c 3 -o $O/r2w_igemm_first3 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2w_ncu_ig.log 2>&1; echo "ncu igemm exit $?"
timeout 800 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_umma -c 3 -o $O/r2w_dw_umma_first3 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2w_ncu_dw.log 2>&1; echo "ncu dw exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2w_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2w_launches.log 2>&1; echo "launch list exit $?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "tc_ or stem or pers_dw or pers_1x1" > $O/r2w_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 $O/r2w_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "tc_c32_s2_rows or tc_c144 or tc_c32_14x14 or pers_dw_s1" > $O/r2w_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -3 $O/r2w_sanitizer_racecheck.log
