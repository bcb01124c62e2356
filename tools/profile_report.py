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

# ---- every kernel of the path (tools/profile_kernels.py under ncu --set full): last (warm) instance per (kernel, grid)
import os
allrep = f"gpurun_out/prof_{tag}_all_raw.csv"   # written on the GPU box by tools/round_evidence.sh (ncu --page raw --csv)
if os.path.exists(allrep):
    raw = open(allrep).read()
    rr = list(csv.reader(raw.splitlines()))
    hdr, units, data = rr[0], rr[1], rr[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        try:
            return float(r[ix[k]].replace(",", ""))
        except Exception:
            return 0.0

    def dur_us(r):
        v, u = num(r, "gpu__time_duration.sum"), units[ix["gpu__time_duration.sum"]]
        return v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v

    def bytes_of(r, k):
        u = units[ix[k]] if k in ix else "byte"
        return num(r, k) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)

    # last full instance per (kernel, grid): launches queued behind a stopped MC batch return at once (no-ops) and
    # would otherwise stand in for the real kernel of that grid
    most = {}
    for r in data:
        key = (r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]])
        most[key] = max(most.get(key, 0.0), num(r, "smsp__inst_executed.sum"))
    last = {}
    for r in data:
        key = (r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]])
        if num(r, "smsp__inst_executed.sum") >= 0.5 * most[key]:
            last[key] = r
    rows = []
    for (kname, grid), r in last.items():
        us = dur_us(r)
        # --set full carries the DFMA count (x2 = flops); DADD / DMUL are not in the set: a lower bound
        # (the derived metric is flops per elapsed cycle)
        fl = num(r, "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2") * num(r, "sm__cycles_elapsed.max")
        rows.append(dict(kernel=kname, grid=grid, duration_us=round(us, 2), regs=int(num(r, "launch__registers_per_thread")),
                         issue_active_pct=round(num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                         fp64_pipe_pct=round(num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 1),
                         warps_active_pct=round(num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                         warp_inst=int(num(r, "smsp__inst_executed.sum")),
                         executed_fp64_gflops=round(fl / (us * 1e-6) / 1e9, 1) if us > 0 else 0.0,
                         l2_gbs=round(32.0 * num(r, "lts__t_sectors.sum") / (us * 1e-6) / 1e9, 1) if us > 0 else 0.0,
                         l2_pct=round(num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), 1),
                         dram_read_bytes=int(bytes_of(r, "dram__bytes_read.sum")), dram_write_bytes=int(bytes_of(r, "dram__bytes_write.sum"))))
    rows.sort(key=lambda x: (x["kernel"], x["grid"]))
    json.dump(rows, open(f"profiles/{tag}_kernels.json", "w"), indent=1)
    print("| kernel | grid | µs | regs | issue % | FP64 pipe % | executed DFMA GFLOP/s | L2 GB/s (% of peak) | DRAM rd/wr bytes |")
    print("|---|---|---|---|---|---|---|---|---|")
    for x in rows:
        print(f"| `{x['kernel']}` | {x['grid']} | {x['duration_us']} | {x['regs']} | {x['issue_active_pct']} | {x['fp64_pipe_pct']} | "
              f"{x['executed_fp64_gflops']} | {x['l2_gbs']} ({x['l2_pct']}) | {x['dram_read_bytes']} / {x['dram_write_bytes']} |")
