#!/usr/bin/env python3
"""Instruction budget of one profiled launch: SASS instructions grouped by how often they executed.
Instructions of one loop body share an execution count, so the histogram reads as "loop X ran N times with
M instructions per trip" — what a kernel that is bound by issue slots (not by a pipe or by memory) needs to know.
usage: tools/ncu_instr_budget.py <report.ncu-rep> <launch index> [min share %]"""
import collections, csv, re, subprocess, sys
rep, launch = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hi[0]]
ix = {n: j for j, n in enumerate(h)}
body = [r for r in rows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(rows))] if len(r) > ix["Instructions Executed"]]


def gi(r):
    try:
        return int(r[ix["Instructions Executed"]])
    except ValueError:
        return 0


tot = sum(gi(r) for r in body)
groups = collections.defaultdict(list)
for r in body:
    groups[gi(r)].append(r)
print(f"launch {launch}: {tot} warp instructions over {len(body)} SASS instructions")
print("| trips | SASS instructions | warp instructions | share | dominant opcodes |")
print("|---|---|---|---|---|")
for trips, rs in sorted(groups.items(), key=lambda kv: -kv[0] * len(kv[1])):
    share = 100.0 * trips * len(rs) / max(tot, 1)
    if share < min_share:
        continue
    ops = collections.Counter()
    for r in rs:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        ops[m.group(2).split(".")[0] if m else "?"] += 1
    print(f"| {trips} | {len(rs)} | {trips * len(rs)} | {share:.1f} % | " + ", ".join(f"{o} {c}" for o, c in ops.most_common(6)) + " |")
