#!/usr/bin/env python3
"""Static SASS statistics per kernel of a built library: instruction count and memory-instruction mix.
usage: sass_stats.py <lib.so>   (needs cuobjdump)"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, stats = None, {}
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); stats[cur] = {}; continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
        op = m.group(1).split('.')[0]
        stats[cur][op] = stats[cur].get(op, 0) + 1
for k, v in stats.items():
    print(k, "instructions=%d (%.1f KB)" % (sum(v.values()), sum(v.values()) * 16 / 1024.0),
          {o: v.get(o, 0) for o in ['LDS', 'STS', 'LD', 'ST', 'LDL', 'STL', 'LDG', 'STG', 'CALL', 'BRA', 'SHFL', 'VOTE', 'REDUX', 'MATCH']})
