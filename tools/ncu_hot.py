#!/usr/bin/env python
"""Top stall sites of one launch in an .ncu-rep: python tools/ncu_hot.py REP [launch_index] [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; idx = sys.argv[2] if len(sys.argv) > 2 else "0"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", idx,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:100])
h = rows[1]; c = {n: i for i, n in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[c['# Samples']] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][c['# Samples']] or 0))[:top]
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
for i in sorted(order):
    r = body[i]; n = int(r[c['# Samples']] or 0)
    st = sorted(((int(r[c[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100*n/tot:5.1f}%  exec={r[c['Instructions Executed']]:>8s}  {r[c['Source']][:80]:80s} {st}")
