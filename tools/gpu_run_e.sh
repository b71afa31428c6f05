#!/bin/bash
# ncu capture of the large (tensor-bound) q8gemm on the current igemm kernel: is it L2-bound?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
cat > /tmp/gemm_big.py <<'PY'
import numpy as np, torch, qnnpack_b200
lib = qnnpack_b200.load()
dev = torch.device("cuda", 0)
M, N, K = 65536, 4096, 4096
rng = np.random.default_rng(0)
w = rng.integers(0, 256, (N, K), dtype=np.uint8); b = rng.integers(-1000, 1000, (N,), dtype=np.int32)
st, op = lib.create_fully_connected(w, b, izp=127, input_scale=1.0, kzp=127, kernel_scale=float(np.float32(1.0 / (128.0 * K ** 0.5))), ozp=127, output_scale=1.0)
assert st == 0
x = torch.randint(0, 256, (M * K,), dtype=torch.uint8, device=dev); y = torch.empty(M * N, dtype=torch.uint8, device=dev)
assert lib.setup_fully_connected(op, M, x.data_ptr(), K, y.data_ptr(), N) == 0
torch.cuda.synchronize()
for _ in range(3):
    assert lib.run(op) == 0
PY
timeout 600 ncu --set full --clock-control none -k regex:q8_igemm -s 1 -c 1 -o $O/r2e_gemm_big python /tmp/gemm_big.py > $O/r2e_ncu.log 2>&1; echo "ncu exit $?"
ncu -i $O/r2e_gemm_big.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for i,c in enumerate(h):
    if any(k in c for k in ('gpu__time_duration.sum','lts__throughput.avg.pct','lts__t_bytes.sum.per_second','l1tex__m_xbar2l1tex_read_bytes.sum.per_second','sm__pipe_tensor_cycles_active.avg.pct','lts__t_sector_hit_rate.pct','dram__bytes_read.sum ','gpu__dram_throughput.avg.pct','lts__t_sectors_srcunit_tex_op_read.sum ','smsp__issue_active.avg.pct','l1tex__throughput.avg.pct','lts__d_sectors_fill','lts__t_sectors.sum.per_second','lts__t_bytes.sum ','sm__throughput.avg.pct')):
        print(c, [r[i] for r in rows[2:]])
"
