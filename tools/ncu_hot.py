#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an .ncu-rep, with their dominant stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; cur = None; out = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "": cur = r[0]; continue
    if r[2] in ("", "..."): continue
    out.append((cur, r))
ix = {h: i for i, h in enumerate(hdr)}
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
seen = set(); res = []
for cur, r in out:
    if r[2] in seen: continue
    seen.add(r[2])
    reasons = {h[6:]: int(r[ix[h]]) for h in st if r[ix[h]].isdigit() and int(r[ix[h]]) > 0}
    res.append((int(r[6]), int(r[2], 16), cur, r[3].strip()[:64], reasons))
tot = sum(x[0] for x in res)
print("total samples", tot)
for samp, addr, cur, src, reasons in sorted(res, key=lambda x: -x[0])[:top]:
    t3 = sorted(reasons.items(), key=lambda x: -x[1])[:3]
    print(f"{100*samp/tot:5.1f}% L{cur:>4} {src:64s} {t3}")
