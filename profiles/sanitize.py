"""compute-sanitizer target: one step of the default pair path on a small synthetic cube (block tiles, lane-per-row
search, weight cache, density presums, near lists) and on the plume fixture (mixed regions, fallbacks, apply_bcond).
usage (GPU box): compute-sanitizer --tool memcheck python profiles/sanitize.py"""
import sys
sys.path[:0] = [".", "tests", "oracle"]
import numpy as np
import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs
from test_gpu_bcond import images_of
from test_gpu_gather_grp import near_rows
import plume_sph_b200 as psph
from plume_sph_b200 import cases

cfg, rows, spare = near_rows(n=16)
pos = rows["coords"]
rng = np.random.default_rng(4)
for p in spare[:10]:
    pos[p] = pos[spare[0]] + rng.uniform(-1, 1, 3) * 0.2     # near-list overflow
s = psph.Solver(cfg)
s.upload(rows)
for _ in range(2):
    print("cube step", s.step(0.0, 0))
out = s.download()
s.close()
z = golden("plume3d_small")
ocfg = op.config_for("plume3d_small", z)
s = psph.Solver(to_product_cfg(ocfg))
attr, _ = cell_attrs(ocfg, z, "b2")
s.set_cells(attr)
s.upload(dump_to_particles(z["bc_pre"]))
s.set_images(images_of(ocfg, z))
s.apply_bcond()
for _ in range(2):
    print("plume step", s.step(0.0, 1))
    s.apply_bcond()
sph = s.download(dtype=psph.PARTICLE_SPH_DTYPE)
s.close()
print("done", len(out), len(sph))
