#!/usr/bin/env python3
"""Per-instruction executed counts of one kernel from an .ncu-rep (source page, SASS view).
usage: python profiles/ncu_sass.py rep.ncu-rep <kernel substring> [units]  -> lines 'count/unit  stall%  sass'"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
unit = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Kernel Name":
        cur = r[1]; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and pat in cur and hdr and len(r) >= len(hdr) - 2:
        rows.append(r)
if not rows:
    sys.exit("no kernel matches")
iS, iE, iT, iSmp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in rows); smp = sum(int(r[iSmp]) for r in rows)
print("# %s : %d SASS instructions, %d warp-instr executed (%.1f per unit), %d samples" % (pat, len(rows), tot, tot / unit, smp))
for r in rows:
    e = int(r[iE])
    print("%9.2f %5.1f%% %5.1f  %s" % (e / unit, 100.0 * int(r[iSmp]) / max(smp, 1), int(r[iT]) / max(e, 1), r[iS].strip()))
