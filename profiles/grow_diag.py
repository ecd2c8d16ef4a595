#!/usr/bin/env python3
"""Debugging aid: step 2 of tests/golden/plume3d_grow.npz (uniform h, after a domain growth) on each pair path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs, relerr, sparse_field, D2P
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE

z = golden("plume3d_grow")
ocfg = op.config_for("plume3d_grow", z)
step = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pre = z["pre%d" % step]
for env in ({}, {"PSPH_NBR": "0"}, {"PSPH_GATHER": "row"}, {"PSPH_NBR": "1"}, {"PSPH_SCAN": "1"}):
    for k in ("PSPH_NBR", "PSPH_GATHER", "PSPH_SCAN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cfg = to_product_cfg(ocfg)
    s = psph.Solver(cfg)
    attr, _ = cell_attrs(ocfg, z, "b%d" % step)
    s.set_cells(attr)
    s.upload(dump_to_particles(pre))
    dt = float(z["dt"][step - 1])
    if step == 1:
        assert s.timestep() == dt
    s.set_dt(dt, step, 1)
    s.search_neighs_consth()
    s.mom_engr_update()
    out = s.download(ORDER_REFERENCE)
    want = sparse_field(z, "s%d_mom_vel" % step, pre["vel"])
    d = np.abs(out["vel"] - want).max(axis=1) / np.abs(want).max()
    bad = np.flatnonzero(d > 1e-12)
    print(env, "vel relerr", d.max(), "rows off", len(bad))
    for r in bad[:5]:
        print("   row", r, "x", pre["coord"][r], "h", pre["smlen"][r], "want", want[r], "got", out["vel"][r])
    s.close()
