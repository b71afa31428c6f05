#!/bin/bash
# ncu of the tensor-core depthwise kernel after the descriptor ring (b1_dw, b3_dw)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_umma -c 2 -o $O/r2j_dw_umma_first2 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2j_ncu_dw.log 2>&1; echo "ncu exit $?"
