#!/usr/bin/env python3
"""Where the warps of one kernel wait: per-SASS-instruction stall samples (ncu source page) with the
dominant stall reasons.  usage: python profiles/ncu_stalls.py rep.ncu-rep <kernel substring> [top N]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur, hdr, rows, done = None, None, [], False
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        if rows: break          # first matching kernel only
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and pat in cur and hdr and len(r) >= len(hdr) - 2:
        rows.append(r)
if not rows: sys.exit("no kernel matches")
iS, iA, iE = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iA] or 0) for r in rows)
agg = {}
for r in rows:
    for i, h in reasons:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("# %s: %d instructions, %d samples" % (pat, len(rows), tot))
print("# by reason: " + "  ".join("%s %.1f%%" % (h[6:], 100.0 * v / tot) for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
order = sorted(range(len(rows)), key=lambda k: -int(rows[k][iA] or 0))[:top]
for k in sorted(order):
    r = rows[k]
    why = sorted(((int(r[i] or 0), h[6:]) for i, h in reasons), reverse=True)[:3]
    print("%5d %5.2f%% x%-9s %-60s %s" % (k, 100.0 * int(r[iA] or 0) / tot, r[iE], r[iS].strip()[:60],
                                          " ".join("%s:%d" % (h, v) for v, h in why if v)))
