#!/bin/bash
# experiment: CTA-pair GEMM for shallower K (wide-N 7x7 / 14x14 expansions)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
for mk in 512 128 64; do
QNNP_CUDA_GEMM2SM_MIN_K=$mk timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/r2l_k$mk.json 2> $O/r2l_k$mk.err; echo "mink $mk exit $?"
done
python - <<'PY'
import json
r = {k: json.load(open("gpurun_out/r2l_k%d.json" % k)) for k in (512, 128, 64)}
print({k: (v["ms_per_step"], (v.get("parity_check") or {}).get("mismatches")) for k, v in r.items()})
for ls in zip(*[r[k]["layers"] for k in (512, 128, 64)]):
    if ls[0]["kind"] != "dw":
        print("   %-12s " % ls[0]["layer"] + "  ".join("%7.3f" % l["ms"] for l in ls))
PY
