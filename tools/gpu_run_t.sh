#!/bin/bash
# ncu of the tensor-core depthwise kernel with the single negated operand (b1_dw, b2_dw stride 2, b3_dw)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_umma -c 3 -o $O/r2t_dw_umma_first3 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2t_ncu_dw.log 2>&1; echo "ncu exit $?"
