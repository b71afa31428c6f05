"""Top stalled SASS instructions of each kernel in an `ncu --page source --csv` dump.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_source_top.py src.csv [N]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    f = lambda x: int(float(x)) if x else 0
    for k in kernels:
        hdr, data = k["hdr"], k["data"]
        isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot, totex = sum(f(r[isamp]) for r in data), sum(f(r[iex]) for r in data)
        print(f"== {k['name'][:90]}\n   samples {tot}  warp-instructions {totex}  sass lines {len(data)}")
        agg = {}
        for r in data:
            for j in stall:
                agg[hdr[j][6:]] = agg.get(hdr[j][6:], 0) + f(r[j])
        print("   stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:8])
        top = sorted(range(len(data)), key=lambda i: -f(data[i][isamp]))[:n]
        for i in sorted(top):
            r = data[i]
            st = sorted(((hdr[j][6:], f(r[j])) for j in stall), key=lambda x: -x[1])[:3]
            print(f"   {i:5d} {f(r[isamp]):7d} {f(r[iex]):9d}  {r[isrc][:64]:64s} {st}")


if __name__ == "__main__":
    main()
