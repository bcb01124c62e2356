#!/usr/bin/env python3
"""profiles/<tag>_kernels.json from the raw-metric page of an `ncu --set full` capture of tools/profile_kernels.py:
one row per (kernel, grid, block) with the median over its captured launches.

  ncu -i prof_all.ncu-rep --page raw --csv > gpurun_out/prof_<tag>_all_raw.csv
  python tools/kernels_summary.py gpurun_out/prof_<tag>_all_raw.csv profiles/<tag>_kernels.json"""
import collections
import csv
import json
import statistics
import sys

SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "s": 1.0, "ns": 1e-9,
         "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]

    def num(row, key):
        if key not in hdr:
            return None
        i = hdr.index(key)
        try:
            return float(row[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        except ValueError:
            return None

    groups = collections.defaultdict(list)
    for row in rows[2:]:
        if len(row) < len(hdr):
            continue
        name = row[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        groups[(name, row[hdr.index("Grid Size")], row[hdr.index("Block Size")])].append(row)
    out = []
    for (name, grid, block), rs in sorted(groups.items()):
        def med(key, f=1.0, nd=2):
            v = [num(r, key) for r in rs]
            v = [x for x in v if x is not None]
            return round(statistics.median(v) * f, nd) if v else None
        cyc = med("sm__cycles_elapsed.max", nd=0)
        flop = None
        if cyc:
            per = lambda k: (med(k, nd=6) or 0.0) * cyc   # noqa: E731
            flop = 2 * per("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") + \
                per("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed") + \
                per("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed")
        dur = med("gpu__time_duration.sum", 1e6)
        out.append({"kernel": name, "grid": grid, "block": block, "launches": len(rs), "duration_us": dur,
                    "regs": med("launch__registers_per_thread", nd=0),
                    "issue_active_pct": med("sm__issue_active.avg.pct_of_peak_sustained_elapsed", nd=1),
                    "fp64_pipe_pct": med("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", nd=1),
                    "warp_inst": med("smsp__inst_executed.sum", nd=0),
                    "executed_fp64_gflops": round(flop / (dur * 1e-6) / 1e9, 1) if (flop and dur) else None,
                    "l2_pct": med("lts__throughput.avg.pct_of_peak_sustained_elapsed", nd=1),
                    "dram_read_bytes": med("dram__bytes_read.sum", nd=0), "dram_write_bytes": med("dram__bytes_write.sum", nd=0)})
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    for r in out:
        print(f"{r['kernel']:28s} {r['grid']:>16s} {r['block']:>14s} n={r['launches']:3d} {r['duration_us']:9.2f} us regs {r['regs']}")


if __name__ == "__main__":
    main()
