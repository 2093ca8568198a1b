#!/usr/bin/env python
"""scripts/ncu_source_mix.py <source.csv> -- opcode mix and stall samples from `ncu --page source --csv` (SASS view)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samples = collections.Counter(); stall = collections.Counter()
tot_i = tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    sass = r[ix["Source"]].strip()
    toks = sass.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("SHFL", "LDG", "STG", "DMMA", "MUFU")) and "." in op else "")
    try:
        n = int(float(r[ix["Instructions Executed"]] or 0)); s = int(float(r[ix["# Samples"]] or 0))
    except ValueError:
        continue          # repeated header (several launches in one report)
    ops[op] += n; samples[op] += s; tot_i += n; tot_s += s
    for c in stall_cols:
        v = r[ix[c]]
        if v: stall[c] += int(float(v))
print(f"total instructions {tot_i:,}  samples {tot_s:,}")
print(f"{'opcode':14s} {'inst':>14s} {'%inst':>7s} {'%samples':>9s}")
for op, n in ops.most_common(28):
    print(f"{op:14s} {n:14,d} {100*n/tot_i:7.2f} {100*samples[op]/max(tot_s,1):9.2f}")
print("stall reasons (all samples):")
ts = sum(stall.values())
for c, v in stall.most_common(8):
    print(f"  {c:28s} {100*v/max(ts,1):6.2f} %")
