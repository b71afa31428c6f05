#!/bin/bash
# ncu of the first three tensor-core GEMM launches (stem, b1_project, b2_expand) on the final round-2 code
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:q8_igemm -c 3 -o $O/r2m_igemm_first3 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2m_ncu.log 2>&1; echo "ncu exit $?"
