"""Compare per-layer times of bench JSON files: python tools/cmp.py a.json b.json ..."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qnnpack_b200 import mobilenet_v2 as M
Ls = M.layers()
runs = [(os.path.basename(f).replace('.json', ''), json.load(open(f))) for f in sys.argv[1:]]
peak = 6484.3
print('ms/step', {n: round(r['ms_per_step'], 2) for n, r in runs})
seen = set()
for i, L in enumerate(Ls):
    key = (L.kind, L.h, L.cin, L.cout, L.stride)
    if key in seen:
        continue
    seen.add(key)
    ideal = L.algorithmic_bytes(runs[0][1]['config']['batch_per_gpu']) / 1e6 / peak
    print(f"{L.name:14s} {L.kind:4s} {L.h:3d} {L.cin:4d}->{L.cout:4d} s{L.stride} ideal={ideal:.3f} " +
          " ".join(f"{n}={r['layers'][i]['ms']:.3f}({ideal / r['layers'][i]['ms']:.2f})" for n, r in runs))
for n, r in runs:
    print(n, {k: (round(v['ms_per_step'], 2), round(v['frac_of_hbm_peak'], 3)) for k, v in r['per_kernel'].items()})
