#!/bin/bash
# Round-2 GPU session F: CTA-pair GEMM kernel — parity first (outer timeout: a wedged kernel must not hold the box),
# then throughput and one ncu capture.
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_gemm2sm.py -x -q > $O/r2f_pytest.log 2>&1; rc=$?; echo "pytest gemm2sm exit $rc"; tail -15 $O/r2f_pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
for shape in "65536 4096 4096" "16384 8192 8192" "65536 1024 1024" "14400 1024 1024"; do
  timeout 120 python tools/gemm_big.py $shape 2>&1 | tail -1
done | tee $O/r2f_gemm.jsonl
QNNP_CUDA_NO_GEMM2SM=1 timeout 120 python tools/gemm_big.py 65536 4096 4096 2>&1 | tail -1 | tee $O/r2f_gemm_1sm.json
timeout 300 ncu --set full --clock-control none -k regex:gemm2sm -s 1 -c 1 -o $O/r2f_gemm2sm python tools/gemm_big.py 65536 4096 4096 1 > $O/r2f_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $O/r2f_gemm2sm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for i,c in enumerate(h):
    if any(k in c for k in ('gpu__time_duration.sum','lts__throughput.avg.pct','sm__pipe_tensor_cycles_active.avg.pct','lts__t_sector_hit_rate.pct','gpu__dram_throughput.avg.pct','smsp__issue_active.avg.pct','l1tex__throughput.avg.pct','lts__t_bytes.sum.per_second','sm__throughput.avg.pct','lts__t_sectors_srcunit_tex_op_read.sum')):
        print(c, [r[i] for r in rows[2:]])
"
