#!/bin/bash
# experiment: igemm folded epilogue without the per-value XOR (accumulators re-armed with 2^31 by tcgen05.st) — timing only
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2k_base.json 2> $O/r2k_base.err; echo "base exit $?"
QNNP_LIB_PATH=$PWD/qnnpack_b200/lib/libqnnpack_preload.so timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-extras > $O/r2k_preload.json 2> $O/r2k_preload.err; echo "preload exit $?"; tail -2 $O/r2k_preload.err
python - <<'PY'
import json
a = json.load(open("gpurun_out/r2k_base.json")); b = json.load(open("gpurun_out/r2k_preload.json"))
print("base ms/step", a["ms_per_step"], "preload", b["ms_per_step"])
for la, lb in zip(a["layers"], b["layers"]):
    if la["kind"] != "dw":
        print("   %-12s base %7.3f  preload %7.3f  %+.1f%%" % (la["layer"], la["ms"], lb["ms"], 100 * (lb["ms"] / la["ms"] - 1)))
PY
