#!/usr/bin/env python
"""Executed warp-instructions, stall samples and shared-memory wavefronts per CUDA source line:
   ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python tools/ncu_lines.py [min-percent]"""
import csv, sys
rows = list(csv.reader(sys.stdin))
minpct = float(sys.argv[1]) if len(sys.argv) > 1 else 0.4
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or not r[0].isdigit() or len(r) < 21:
        continue
    try:
        ex = int(float(r[ix["Instructions Executed"]]))
        sm = int(float(r[ix["# Samples"]]))
        wf = float(r[ix["L1 Wavefronts Shared"]] or 0)
    except ValueError:
        # a source line with embedded quotes: take the numeric columns from the right
        try:
            off = len(r) - len(hdr)
            ex = int(float(r[ix["Instructions Executed"] + off]))
            sm = int(float(r[ix["# Samples"] + off]))
            wf = float(r[ix["L1 Wavefronts Shared"] + off] or 0)
        except Exception:
            continue
    out.append((cur_file, int(r[0]), ex, sm, wf, r[1].strip()))
te, ts, tw = sum(o[2] for o in out), sum(o[3] for o in out), sum(o[4] for o in out)
print(f"total executed {te}, samples {ts}, smem wavefronts {tw:.0f}")
for fn, ln, ex, sm, wf, src in out:
    if ex * 100.0 / max(te, 1) >= minpct or sm * 100.0 / max(ts, 1) >= minpct:
        print(f"{fn[:16]:<16}{ln:>5} ex={ex:>10} {100*ex/te:5.1f}% smp={sm:>6} {100*sm/max(ts,1):5.1f}% wf={wf:>10.0f} | {src[:100]}")
