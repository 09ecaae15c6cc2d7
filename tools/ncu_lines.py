#!/usr/bin/env python
"""Per-opcode and per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (no GPU needed)."""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; lines = []; sass = []; cur = None
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "":
        try: lines.append((int(r[0]), r[1], int(r[7]), int(r[6])))
        except ValueError: pass
    elif r[2] not in ("", "..."):
        try: sass.append((r[3], int(r[7]), int(r[6]), r))
        except ValueError: pass
ti = sum(s[1] for s in sass); ts = sum(s[2] for s in sass)
print(f"total warp-instructions {ti}  samples {ts}")
ci = Counter(); cs = Counter()
for src, n, s, _ in sass:
    t = src.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    ci[op] += n; cs[op] += s
for op, c in ci.most_common(18):
    print(f"  {op:8s} inst {c:11d} {100*c/ti:5.1f}%   samples {100*cs[op]/max(ts,1):5.1f}%")
ix = {h: i for i, h in enumerate(hdr)}
for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h:
        s = sum(int(r[3][ix[h]]) for r in sass if r[3][ix[h]].isdigit())
        if s * 50 > ts: print(f"  {h:24s} {100*s/ts:5.1f}%")
agg = Counter(); aggs = Counter(); txt = {}
for ln, src, n, s in lines:
    agg[ln] += n; aggs[ln] += s; txt[ln] = src
for ln, n in agg.most_common(top):
    print(f"  L{ln:4d} inst {n:10d} {100*n/ti:5.1f}% samp {100*aggs[ln]/max(ts,1):5.1f}% | {txt[ln].strip()[:110]}")
