#!/usr/bin/env python3
"""Turns gpurun_out/launches_rNN.csv + prof_rNN_kmove.ncu-rep into the committed summaries under profiles/."""
import collections, csv, json, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = {}
# ---- launch list
lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
    agg[name].append((v, row["Grid Size"]))
tot = sum(v for x in agg.values() for v, _ in x)
rows = []
for k, x in sorted(agg.items(), key=lambda kv: -sum(v for v, _ in kv[1])):
    vs = sorted(v for v, _ in x)
    rows.append(dict(kernel=k, launches=len(vs), total_us=round(sum(vs), 1), share=round(sum(vs) / tot, 4),
                     min_us=round(vs[0], 2), median_us=round(vs[len(vs) // 2], 2), max_us=round(vs[-1], 2)))
out["launch_list"] = rows
by_grid = collections.defaultdict(list)
for k, x in agg.items():
    if k.startswith("k_move"):
        for v, g in x:
            by_grid[g].append(v)
out["k_move_by_grid"] = {g: dict(n=len(v), mean_us=round(sum(v) / len(v), 2)) for g, v in sorted(by_grid.items())}
# ---- full profile
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}_kmove.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2:]
ix = {h: i for i, h in enumerate(hdr)}
want = ["Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]
prof = []
for r in data:
    prof.append({w: (r[ix[w]] + " " + units[ix[w]]).strip() for w in want if w in ix})
out["k_move_full"] = prof
json.dump(out, open(f"profiles/{tag}_ncu_summary.json", "w"), indent=1)
print(json.dumps(out["launch_list"], indent=1))
print(json.dumps(out["k_move_by_grid"], indent=1))
for p in prof:
    print(p)
