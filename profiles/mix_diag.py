#!/usr/bin/env python3
"""Debugging aid: is the mixed-h path (PSPH_NBR=4) deterministic run to run, and where do the default (PSPH_NBR unset:
the device picks) and the forced path differ?   usage: mix_diag.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import test_gpu_gather_grp as T

cfg, rows = T.mixed_h_rows(cluster=0)
runs = []
for nbr in ("4", "4", None, None, "0"):
    dt, bad, out = T.run_stages(cfg, rows, None, nbr=nbr)
    runs.append((nbr, out))
    print("run", nbr, "dt", dt, "bad", bad)
def cmp(a, b, tag):
    for f in T.FIELDS:
        x, y = a[f], b[f]
        d = np.abs(x - y); d = d if d.ndim == 1 else d.max(axis=1)
        nz = np.flatnonzero(d > 0)
        print(tag, f, "rows differing", len(nz), "max rel", float(d.max() / max(np.abs(x).max(), 1e-300)),
              "bc", dict(zip(*np.unique(a["bc_type"][nz], return_counts=True))) if len(nz) else {})
cmp(runs[0][1], runs[1][1], "forced4 vs forced4:")
cmp(runs[2][1], runs[3][1], "default vs default:")
cmp(runs[0][1], runs[2][1], "forced4 vs default:")
cmp(runs[0][1], runs[4][1], "forced4 vs fused:")
