#!/usr/bin/env python3
"""One `ncu --set full` capture -> the figures bench.py quotes beside its roofline (profiles/traffic.json entry).
usage: ncu_summary.py <raw.csv> <key> <label/source text> <game-cycles of the captured launch> <games> [traffic.json]
Merges {key: {dram_bytes_per_launch, warp_inst_per_game_cycle, lanes_active, issue_active_pct, duration_ms, ...}} into traffic.json."""
import csv, json, sys
raw, key, source, cycles, games = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), int(sys.argv[5])
path = sys.argv[6] if len(sys.argv) > 6 else "profiles/traffic.json"
rows = list(csv.reader(open(raw)))
names, units, vals = rows[0], rows[1], rows[2]
m = {}
for n, u, v in zip(names, units, vals):
    try:
        m[n] = (float(v.replace(",", "")), u)
    except ValueError:
        pass
def val(name, scale_units=True):
    if name not in m:
        return None
    v, u = m[name]
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(u, 1.0)
    return v * mult if scale_units else v
e = dict(source=source, games=games, game_cycles_in_launch=cycles,
         dram_bytes_per_launch=(val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
         dram_bytes_read=val("dram__bytes_read.sum"), dram_bytes_write=val("dram__bytes_write.sum"),
         duration_ms=val("gpu__time_duration.sum"),
         warp_inst=val("smsp__inst_executed.sum", False),
         lanes_active=val("smsp__thread_inst_executed_per_inst_executed.ratio", False),
         issue_active_pct=val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
         warps_active_per_sm=val("sm__warps_active.avg.per_cycle_active", False),
         registers_per_thread=val("launch__registers_per_thread", False),
         stall_no_instruction=val("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", False),
         stall_long_scoreboard=val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", False))
if e["warp_inst"] and cycles > 0:
    e["warp_inst_per_game_cycle"] = e["warp_inst"] / cycles
try:
    t = json.load(open(path))
except Exception:
    t = {}
t[key] = e
json.dump(t, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps({key: e}))
