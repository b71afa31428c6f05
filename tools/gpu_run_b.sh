#!/bin/bash
# Round-2 GPU session B: new-operator parity, depthwise A/B (round-1 kernel vs current) + ncu of the tcgen05 depthwise
# kernel, large (tensor-bound) q8gemm through the FC operator.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q > $O/r2b_pytest_ops.log 2>&1; echo "pytest ops exit $?"; tail -3 $O/r2b_pytest_ops.log
QNNP_LIB_PATH=$PWD/qnnpack_b200/lib/libqnnpack_olddw.so timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > $O/r2b_bench_olddw.json 2> $O/r2b_bench_olddw.err; echo "bench olddw exit $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > $O/r2b_bench_cur.json 2> $O/r2b_bench_cur.err; echo "bench cur exit $?"
timeout 300 python - > $O/r2b_gemm.json 2> $O/r2b_gemm.err <<'PY'
import json, numpy as np, torch, qnnpack_b200
lib = qnnpack_b200.load()
dev = torch.device("cuda", 0)
out = []
for (M, N, K) in ((65536, 4096, 4096), (14400, 1024, 1024), (65536, 1024, 1024), (16384, 8192, 8192)):
    rng = np.random.default_rng(0)
    w = rng.integers(0, 256, (N, K), dtype=np.uint8)
    b = rng.integers(-1000, 1000, (N,), dtype=np.int32)
    st, op = lib.create_fully_connected(w, b, izp=127, input_scale=1.0, kzp=127, kernel_scale=float(np.float32(1.0 / (128.0 * K ** 0.5))),
                                        ozp=127, output_scale=1.0)
    assert st == 0
    x = torch.randint(0, 256, (M * K,), dtype=torch.uint8, device=dev)
    y = torch.empty(M * N, dtype=torch.uint8, device=dev)
    assert lib.setup_fully_connected(op, M, x.data_ptr(), K, y.data_ptr(), N) == 0
    torch.cuda.synchronize()
    for _ in range(2):
        assert lib.run(op) == 0
    s = torch.cuda.Stream(device=dev)
    lib.set_stream(s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record(s)
    for _ in range(reps):
        assert lib.run_async(op) == 0
    e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    lib.set_stream(0)
    out.append({"m": M, "n": N, "k": K, "ms": ms, "tops": 2.0 * M * N * K / ms / 1e9})
    lib.delete(op)
print(json.dumps(out))
PY
echo "gemm exit $?"; cat $O/r2b_gemm.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv3x3_umma -c 2 -o $O/r2b_dw_umma_first2 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check > $O/r2b_ncu_dw.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import json
a = json.load(open("gpurun_out/r2b_bench_olddw.json")); b = json.load(open("gpurun_out/r2b_bench_cur.json"))
print("olddw ms/step", a["ms_per_step"], "cur", b["ms_per_step"])
for la, lb in zip(a["layers"], b["layers"]):
    if la["kind"] == "dw":
        print("   %-10s old %7.3f  cur %7.3f" % (la["layer"], la["ms"], lb["ms"]))
PY
