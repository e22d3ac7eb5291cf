#!/usr/bin/env python3
"""Join an `ncu --page source --print-source sass --csv` dump of k_step with nvdisasm line info and aggregate executed
instructions / stall samples per source line and per enclosing function.
usage: sass_by_line.py <ncu_sass.csv> <nvdisasm -g output> [engine.cuh] [kernel symbol substring, default k_step_fast]"""
import csv, re, sys, collections
sass_csv, dis, src = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None
kern = sys.argv[4] if len(sys.argv) > 4 else 'k_step_fast'
# instruction -> (file,line) from nvdisasm, for function k_step
lines = []
cur = None; infn = False
for l in open(dis):
    if l.startswith('//---') and '.text.' in l:
        infn = (kern + '10StepParams') in l or l.rstrip().endswith(kern)
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])  # inst, samples, no_inst, long_sb, wait+short, static count
tot = [0, 0]
for r, ln in zip(data, lines):
    a = agg[ln]
    inst = int(r[ix['Instructions Executed']]); smp = int(r[ix['# Samples']])
    a[0] += inst; a[1] += smp; a[2] += int(r[ix['stall_no_inst']]); a[3] += int(r[ix['stall_long_sb']])
    a[4] += int(r[ix['stall_wait']]) + int(r[ix['stall_short_sb']]) + int(r[ix['stall_mio']]); a[5] += 1
    tot[0] += inst; tot[1] += smp
# function ranges from the source file
fn_of = {}
if src:
    cur_fn = None
    for i, l in enumerate(open(src), 1):
        m = re.match(r'(?:DEVN?|static|template.*)?\s*(?:DEV|DEVN)\s+[\w:<>\*&\s]+?\s+\**&?(\w+)\s*\(', l)
        if m: cur_fn = m.group(1)
        fn_of[i] = cur_fn
fagg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
for (f, ln), a in agg.items():
    key = fn_of.get(ln, f) if f == 'engine.cuh' else f
    for k in range(6): fagg[key][k] += a[k]
print("total inst %d samples %d static %d" % (tot[0], tot[1], len(lines)))
print("%-28s %8s %7s %7s %7s %7s %7s" % ("function", "inst%", "smp%", "noinst%", "longsb%", "wait%", "static"))
for k, a in sorted(fagg.items(), key=lambda kv: -kv[1][1]):
    print("%-28s %8.2f %7.2f %7.2f %7.2f %7.2f %7d" % (k, 100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1], 100.0 * a[2] / tot[1], 100.0 * a[3] / tot[1], 100.0 * a[4] / tot[1], a[5]))
print()
print("top lines by samples")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%s:%d fn=%s inst%%=%.2f smp%%=%.2f noinst=%d longsb=%d wait=%d static=%d" % (f, ln, fn_of.get(ln), 100.0 * a[0] / tot[0], 100.0 * a[1] / tot[1], a[2], a[3], a[4], a[5]))
