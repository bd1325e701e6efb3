#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples per SASS opcode,
and the hottest instruction addresses. Usage: ncu -i X.ncu-rep --page source --csv | python tools/ncu_sass_summary.py"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops = defaultdict(lambda: [0, 0, 0])
tot_exec = tot_samp = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    sass = r[ix["Source"]].strip()
    ex = int(float(r[ix["Instructions Executed"]] or 0))
    sm = int(float(r[ix["# Samples"]] or 0))
    tok = sass.split()
    op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
    op = op.rstrip(";")
    base = op.split(".")[0]
    if base in ("LDL", "STL", "LDS", "STS", "LDG", "STG"):
        base = op if base in ("LDL", "STL") else base
    ops[base][0] += ex
    ops[base][1] += sm
    ops[base][2] += 1
    tot_exec += ex
    tot_samp += sm
    lines.append((sm, ex, r[ix["Address"]], sass, r))
print(f"total executed warp-instr {tot_exec}, samples {tot_samp}, static instr {len(lines)}")
print(f"{'opcode':<18}{'executed':>14}{'%exec':>8}{'samples':>10}{'%samp':>8}{'static':>8}")
for op, (ex, sm, n) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{op:<18}{ex:>14}{100 * ex / max(tot_exec, 1):>8.1f}{sm:>10}{100 * sm / max(tot_samp, 1):>8.1f}{n:>8}")
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
print("\nhottest instructions by stall samples:")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for sm, ex, addr, sass, r in sorted(lines, key=lambda t: -t[0])[:top]:
    st = sorted(((int(float(r[ix[c]] or 0)), c) for c in stall_cols), reverse=True)[:2]
    print(f"{sm:>7} {ex:>10} {addr[-6:]} {sass[:70]:<70} {st}")
