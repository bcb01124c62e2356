#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics per launch + stall/opcode breakdown + hottest source lines."""
import collections, csv, re, subprocess, sys
rep = sys.argv[1]
launch = sys.argv[2] if len(sys.argv) > 2 else "0"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for w in want:
    if w in idx:
        print(f"{w:72s} {units[idx[w]]:>10s} " + ' '.join(f"{r[idx[w]][:28]:>14s}" for r in data))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", launch, "--launch-count", "1",
                      "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
h2 = None
for i, r in enumerate(srows):
    if r and r[0] == 'Address':
        h2 = r; body = srows[i + 1:]; break
if h2 is None:
    for i, r in enumerate(srows):
        if '# Samples' in r:
            h2 = r; body = srows[i + 1:]; break
ix = {h: i for i, h in enumerate(h2)}
def gi(r, k):
    try: return int(r[ix[k]] or 0)
    except Exception: return 0
body = [r for r in body if len(r) > ix['# Samples']]
tot = sum(gi(r, '# Samples') for r in body) or 1
print('sass instructions', len(body), 'samples', tot)
stalls = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(gi(r, h) for r in body) for h in stalls}
print('stalls:', ', '.join(f"{h[6:]}={v/tot:.2f}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
ops = collections.Counter(); samp = collections.Counter()
srcname = 'Source'
for r in body:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix[srcname]])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += gi(r, 'Instructions Executed'); samp[op] += gi(r, '# Samples')
te = sum(ops.values()) or 1
print('warp instr executed', te)
print('ops:', ', '.join(f"{o}={c/te:.3f}" for o, c in ops.most_common(22)))
top = sorted(body, key=lambda r: -gi(r, '# Samples'))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]
for r in top:
    print(f"{gi(r,'# Samples'):6d} {gi(r,'Instructions Executed'):9d}  {r[ix[srcname]][:110]}")
