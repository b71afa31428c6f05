"""Summarise an Nsight Compute report into the handful of numbers DESIGN.md / bench.py quote.
usage: python tools/ncu_summary.py report.ncu-rep [> profiles/xyz.summary.txt]   (needs `ncu` on PATH, no GPU)"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n## {name}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        vals = dict(zip(hdr, r))
        unit = dict(zip(hdr, units))
        for k in WANT:
            if k in vals and vals[k] != "":
                print(f"  {k:75s} {vals[k]:>16s} {unit[k]}")
        try:
            rd = float(vals["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit["dram__bytes_read.sum"]]
            wr = float(vals["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit["dram__bytes_write.sum"]]
            t = float(vals["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}[unit["gpu__time_duration.sum"]]
            print(f"  {'=> dram traffic (read+write)':75s} {(rd + wr) / 1e9:16.3f} GB   -> {(rd + wr) / t / 1e9:.0f} GB/s under the profiler")
        except Exception:
            pass
        st = sorted(((h[len(STALL):].replace("_per_issue_active.ratio", ""), float(v)) for h, v in vals.items()
                     if h.startswith(STALL) and v not in ("",)), key=lambda x: -x[1])
        print("  top stall reasons (warps per issue-active cycle): " + ", ".join(f"{a}={b:.2f}" for a, b in st[:6]))


if __name__ == "__main__":
    main()
