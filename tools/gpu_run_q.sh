#!/bin/bash
# two-GPU check of the bench contract (both arms launched the way the driver launches them)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2q_bench2.json 2> $O/r2q_bench2.err; echo "bench2 exit $?"; tail -3 $O/r2q_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2q_ref2.json 2> $O/r2q_ref2.err; echo "ref2 exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("metric", "value", "n_gpus", "ms_per_step", "scaling", "gpu_launches")}, "parity", (d.get("parity_check") or {}).get("mismatches"), "e2e", d["e2e"]["value"])
r = json.loads(open("gpurun_out/r2q_ref2.json").read().strip().splitlines()[-1]); print({k: r.get(k) for k in ("impl", "value", "n_gpus")})
PY
