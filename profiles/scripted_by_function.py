#!/usr/bin/env python3
"""Per-function / per-line breakdown of a kernel profile over engine.cuh + scripted.cuh.
usage: scripted_by_function.py <ncu sass csv> <nvdisasm -g output> <raw csv> [mangled kernel name]"""
import csv, re, collections, sys, os
sass, dis, raw = sys.argv[1:4]
kern = sys.argv[4] if len(sys.argv) > 4 else '_Z6k_step10StepParams'
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'microrts_b200', 'csrc')
lines = []; cur = None; infn = False
for l in open(dis):
    if l.startswith('//---') and '.text.' in l:
        infn = kern in l; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): lines.append(cur)
rows = list(csv.reader(open(sass)))
hdr = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0]
for r, ln in zip(data, lines):
    a = agg[ln]; n = int(r[ix['Instructions Executed']]); sm = int(r[ix['# Samples']]); a[0] += n; a[1] += sm; a[2] += 1; tot[0] += n; tot[1] += sm
srcs = {f: open(os.path.join(root, f)).read().splitlines() for f in ('scripted.cuh', 'engine.cuh')}
fn = {}
for f, src in srcs.items():
    curf = None
    for i, l in enumerate(src, 1):
        m = re.match(r'(?:template\s*<[^>]*>\s*)?(?:DEV|DEVN)\s+[\w:<>\*&\s]+?\s+\**&?(\w+)\s*\(', l)
        if m: curf = m.group(1)
        fn[(f, i)] = curf
f = collections.defaultdict(lambda: [0, 0])
for (fl, ln), a in [(k, v) for k, v in agg.items() if k]:
    key = fn.get((fl, ln), fl); f[key][0] += a[0]; f[key][1] += a[1]
print("total warp instructions", tot[0])
for k, a in sorted(f.items(), key=lambda kv: -kv[1][1])[:24]: print("%-28s inst%%=%.2f smp%%=%.2f" % (k, 100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1]))
print()
for (fl, ln), a in sorted([(k, v) for k, v in agg.items() if k], key=lambda kv: -kv[1][1])[:24]:
    print(fl, ln, "inst%%=%.2f smp%%=%.2f n=%d" % (100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1], a[2]), (srcs[fl][ln - 1].strip()[:100] if fl in srcs else ''))
rows = list(csv.reader(open(raw)))
h = rows[0]; u = rows[1]; d = rows[2]
for i, n in enumerate(h):
    if n in ('gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size', 'smsp__thread_inst_executed_per_inst_executed.ratio') or ('issue_stalled' in n and 'per_issue_active' in n and float(d[i]) > 0.3): print(n, u[i], d[i])
