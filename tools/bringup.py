"""Kernel bring-up on a real B200 (run under gpurun): each step in its own subprocess with a timeout so
that a hung kernel cannot take the remaining steps with it.  Writes everything to gpurun_out/."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

STEP_IGEMM_DUMP = r'''
import numpy as np, torch, sys, os
sys.path.insert(0, %(root)r)
import qnnpack_b200
from oracle import q8_oracle as O
lib = qnnpack_b200.load()
co = O.COracle()
M, K, N = %(M)d, %(K)d, %(N)d
rng = np.random.default_rng(0)
x = rng.integers(0, 256, (M, K), dtype=np.uint8)
w = rng.integers(0, 256, (N, K), dtype=np.uint8)
b = rng.integers(-1000, 1000, (N,), dtype=np.int32)
import ctypes as C
kw = dict(izp=%(izp)d, input_scale=1.0, kzp=%(kzp)d, kernel_scale=1.0, ozp=128, output_scale=%(oscale)f)
st, op = lib.create_fully_connected(w, b, **kw)
assert st == 0
folded = lib.lib.qnnp_cuda_debug_operator_is_folded(op)
lib.delete(op)
plan = (C.c_int * 24)()
lib.lib.qnnp_cuda_debug_plan_igemm.argtypes = [C.c_size_t, C.c_size_t, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int)]
assert lib.lib.qnnp_cuda_debug_plan_igemm(K, N, 1, folded, 1, plan) == 1
mt, n_tiles, n_tile, n_mma = plan[4], plan[5], plan[6], plan[7]
print("plan: folded", folded, "mt", mt, "skc", plan[2], "k_stages", plan[3], "n_tiles", n_tiles, "n_tile", n_tile, "n_mma", n_mma,
      "resident", plan[9], "stages", plan[10])
m_tiles = -(-M // 128)
m_super = -(-m_tiles // mt)
items = m_super * n_tiles
dump = torch.full((items * mt * 128 * n_mma,), -777, dtype=torch.int32, device="cuda")
lib.lib.qnnp_cuda_debug_set_accumulator_dump(dump.data_ptr())
y = lib.fully_connected(x, w, b, **kw)
lib.lib.qnnp_cuda_debug_set_accumulator_dump(None)
torch.cuda.synchronize()
acc = dump.cpu().numpy().reshape(items, mt, 128, n_mma)
if folded:   # accumulator == the reference accumulator: bias + sum (a - izp)(w - kzp)
    want_raw = (x.astype(np.int64) - %(izp)d) @ (w.astype(np.int64) - %(kzp)d).T + b.astype(np.int64)
else:        # raw u8 x u8 products; bias and zero points are applied in the epilogue
    want_raw = x.astype(np.int64) @ w.astype(np.int64).T
want_sum = x.astype(np.int64).sum(1)
ok_raw = True
for it in range(items):
    st_, nt = it // n_tiles, it %% n_tiles
    for j in range(mt):
        r0 = (st_ * mt + j) * 128
        rows = min(128, M - r0)
        if rows <= 0:
            continue
        cols = min(n_tile, N - nt * n_tile)
        got = acc[it, j, :rows, :cols]
        ref = want_raw[r0:r0+rows, nt*n_tile:nt*n_tile+cols]
        if not np.array_equal(got, ref):
            ok_raw = False
            bad = np.argwhere(got != ref)
            print("item", it, "sub", j, "raw acc mismatches", len(bad), "of", got.size, "first", bad[:4].tolist())
            print(" got ", got[:3, :8].tolist()); print(" want", ref[:3, :8].tolist())
            print(" diff", (got[:3, :8] - ref[:3, :8]).tolist())
        if not folded:
            gs = acc[it, j, :rows, n_tile]
            if not np.array_equal(gs, want_sum[r0:r0+rows]):
                ok_raw = False
                print("item", it, "sub", j, "rowsum mismatch: got", gs[:6].tolist(), "want", want_sum[r0:r0+6].tolist())
want = co.fully_connected(x, w, b, **kw)
print("RAW_OK", ok_raw, "OUT_OK", bool(np.array_equal(y, want)), "mismatching bytes", int((y != want).sum()), "of", y.size)
np.savez_compressed(os.path.join(%(out)r, "igemm_dump_%(M)d_%(K)d_%(N)d.npz"), acc=acc, x=x, w=w, y=y, want=want)
'''

STEPS = [
    ("env", "import torch; print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0)); "
            "import subprocess; print(subprocess.run(['nvidia-smi'], capture_output=True, text=True).stdout[:1500])", 45),
    ("igemm 128x32x16 nozp", dict(M=128, K=32, N=16, izp=0, kzp=0, oscale=40000.0), 45),
    ("igemm 128x32x16 kzp128", dict(M=128, K=32, N=16, izp=0, kzp=128, oscale=40000.0), 45),
    ("igemm 200x24x144", dict(M=200, K=24, N=144, izp=9, kzp=255, oscale=40000.0), 45),
    ("igemm 128x32x16 zp", dict(M=128, K=32, N=16, izp=7, kzp=5, oscale=40000.0), 45),
    ("igemm 300x144x24", dict(M=300, K=144, N=24, izp=7, kzp=5, oscale=90000.0), 45),
    ("igemm 1000x64x384", dict(M=1000, K=64, N=384, izp=127, kzp=127, oscale=400.0), 45),
    ("igemm 700x1280x1000", dict(M=700, K=1280, N=1000, izp=127, kzp=127, oscale=2000.0), 45),
    ("igemm 5000x32x16", dict(M=5000, K=32, N=16, izp=3, kzp=200, oscale=40000.0), 45),
    ("igemm 3000x192x32", dict(M=3000, K=192, N=32, izp=127, kzp=127, oscale=900.0), 45),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    results = {}
    for name, spec, tmo in STEPS:
        if spec is None:
            cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-x",
                   "-k", "requant or q8dwconv or dw3x3 or dw5x5 or grouped or routing", "-p", "no:cacheprovider"]
        elif isinstance(spec, str):
            cmd = [sys.executable, "-c", spec]
        else:
            cmd = [sys.executable, "-c", STEP_IGEMM_DUMP % dict(root=ROOT, out=OUT, **spec)]
        print(f"===== {name}", flush=True)
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=tmo, cwd=ROOT)
            tail = (p.stdout + p.stderr)[-3000:]
            print(tail, flush=True)
            results[name] = dict(rc=p.returncode)
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT (hung kernel?)", (e.stdout or b"")[-1500:], (e.stderr or b"")[-1500:], flush=True)
            results[name] = dict(rc="timeout")
    json.dump(results, open(os.path.join(OUT, "bringup.json"), "w"), indent=1)
    print(json.dumps(results))


if __name__ == "__main__":
    main()
