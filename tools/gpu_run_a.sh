#!/bin/bash
# Round-2 GPU session A: parity (all GPU tests incl. the new chain / >4 GiB / odd-base tests), bench with the parity gate,
# A/B against the round-1 epilogue, int8 tensor peak, ncu capture of the first three igemm launches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest.log
tail -5 $O/r2a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err; echo "bench exit $?"
QNNP_CUDA_NO_PANEL_STORE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check > $O/r2a_bench_nopanel.json 2> $O/r2a_bench_nopanel.err; echo "bench nopanel exit $?"
timeout 120 python - > $O/r2a_peak.json 2> $O/r2a_peak.err <<'PY'
import json, qnnpack_b200
lib = qnnpack_b200.load()
out = []
for iters in (500, 4000, 20000):
    tops, ms = lib.measure_int8_peak(iters, 5)
    out.append({"iters": iters, "tops": tops, "ms_per_launch": ms})
print(json.dumps(out))
PY
echo "peak exit $?"; cat $O/r2a_peak.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q8_igemm -c 3 -o $O/r2a_igemm_first3 \
  python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity-check > $O/r2a_ncu_igemm.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/r2a_bench.json", "gpurun_out/r2a_bench_nopanel.json"):
    try:
        d = json.load(open(f))
        print(f, "ms/step", round(d["ms_per_step"], 3), "value", round(d["value"]), "parity", (d.get("parity_check") or {}).get("mismatches"),
              "e2e", d.get("e2e") and round(d["e2e"]["value"]))
        for l in d["layers"]:
            print("   %-14s %-5s %7.3f ms %7.0f GB/s" % (l["layer"], l["kind"], l["ms"], l["gbs"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
