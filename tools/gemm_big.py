"""Large q8gemm through the fully-connected operator (device pointers): prints TOPS; used under ncu by the GPU sessions."""
import json
import sys

import numpy as np
import torch

import qnnpack_b200

lib = qnnpack_b200.load()
dev = torch.device("cuda", 0)
M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (65536, 4096, 4096)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
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
e0.record(s)
for _ in range(reps):
    assert lib.run_async(op) == 0
e1.record(s)
s.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"m": M, "n": N, "k": K, "ms": ms, "tops": 2.0 * M * N * K / ms / 1e9}))
