#!/bin/bash
# ncu of the 24 -> 144 expansion (K = 24: cp.async loader path)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q8_igemm --launch-skip 4 -c 1 -o gpurun_out/r2z_b3_expand \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > gpurun_out/r2z_ncu.log 2>&1; echo "ncu exit $?"
