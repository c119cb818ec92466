#!/usr/bin/env python3
"""Aggregate `ncu --page source --csv --print-source sass` by instruction-address buckets.
usage: sass_profile.py src.csv [bucket_instrs=64]
Prints per bucket: first instruction index, executed warp instructions, samples, and the main stall reasons."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[col["Instructions Executed"]] or 0) for r in body)
tot_s = sum(int(r[col["# Samples"]] or 0) for r in body)
print("instructions %d, executed warp instructions %d, samples %d" % (len(body), tot_inst, tot_s))
agg = {s: 0 for s in stalls}
for r in body:
    for s in stalls:
        agg[s] += int(r[col[s]] or 0)
print("stalls overall:", ", ".join("%s %.1f%%" % (s[6:], 100.0 * v / tot_s) for s, v in sorted(agg.items(), key=lambda x: -x[1]) if v))
for b0 in range(0, len(body), B):
    blk = body[b0:b0 + B]
    ie = sum(int(r[col["Instructions Executed"]] or 0) for r in blk)
    sm = sum(int(r[col["# Samples"]] or 0) for r in blk)
    if sm < 0.004 * tot_s and ie < 0.004 * tot_inst:
        continue
    st = {s: sum(int(r[col[s]] or 0) for r in blk) for s in stalls}
    top = ", ".join("%s %d" % (s[6:], v) for s, v in sorted(st.items(), key=lambda x: -x[1])[:4] if v)
    ops = {}
    for r in blk:
        op = r[col["Source"]].split()[0] if r[col["Source"]].split() else "?"
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + 1
    sig = " ".join("%s%d" % (k, v) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
    print("%5d  inst %5.1f%%  samp %5.1f%%  | %s | %s" % (b0, 100.0 * ie / tot_inst, 100.0 * sm / tot_s, top, sig))
