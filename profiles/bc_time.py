import sys, numpy as np
sys.path[:0] = [".", "tests", "oracle"]
import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs
from test_gpu_bcond import images_of
import plume_sph_b200 as psph
for name in ("plume3d_small", "jpue3d"):
    z = golden(name); ocfg = op.config_for(name, z)
    s = psph.Solver(to_product_cfg(ocfg))
    attr, _ = cell_attrs(ocfg, z, "b2"); s.set_cells(attr)
    s.upload(dump_to_particles(z["bc_pre"])); s.set_images(images_of(ocfg, z))
    s.apply_bcond(); s.sync(); s.set_timing(True)
    for _ in range(10): s.apply_bcond()
    s.sync()
    print(name, "rows", len(z["bc_pre"]), "images", len(z["bc_img_row"]), "apply_bcond ms/call", s.stage_ms()["apply_bcond"] / 10)
    s.close()
