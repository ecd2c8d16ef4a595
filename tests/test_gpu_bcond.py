"""-m gpu: psph_apply_bcond (CUDA, through the C ABI) against the unmodified reference's apply_bcond at the end of
time step 1 (main.cc:468): state before the call, the reference's image table (BndImage) and the wall ghosts' state
after the call all come from the golden fixtures (tests/golden/make_golden.py, `bc_*` arrays).

Bars: wall ghosts with an image take the interpolated state within RTOL_STAGE of the field's max magnitude (the
device sums the partners of an image in lane order, the reference in plist order); every other row — wall ghosts
without an image included — is left bit-identical.
"""
import numpy as np
import pytest

import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs, image_cells, relerr, RTOL_STAGE
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE

pytestmark = pytest.mark.gpu

FIELDS = (("rho", "rho"), ("vel", "vel"), ("engr", "engr"), ("mass_frac", "mass_frac"), ("pressure", "pressure"),
          ("sound_speed", "sound_speed"), ("temperature", "temperature"), ("phase_density", "phase_density"),
          ("gamma", "gamma"), ("specific_heat_p", "specific_heat_p"))


def images_of(ocfg, z):
    img = np.zeros(len(z["bc_img_row"]), psph.IMAGE_DTYPE)
    img["ghost_key"] = z["bc_pre"]["key"][z["bc_img_row"]]
    img["cell"] = image_cells(ocfg, z)
    img["coord"] = z["bc_img_coord"]
    return img


@pytest.mark.parametrize("name", ["plume3d_small", "jpue3d"])
def test_apply_bcond_matches_reference(name):
    z = golden(name)
    ocfg = op.config_for(name, z)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, z, "b2")
        s.set_cells(attr)
        pre = z["bc_pre"]
        s.upload(dump_to_particles(pre))
        img = images_of(ocfg, z)
        rng = np.random.default_rng(5)
        s.set_images(img[rng.permutation(len(img))])       # table order is the driver's business
        assert s.apply_bcond() == 0
        assert s.errors() == 0
        out = s.download(ORDER_REFERENCE)
        assert np.array_equal(out["key"], pre["key"])
        w = z["bc_wall_rows"]
        for f, pf in FIELDS:
            e = relerr(out[pf][w], z["bc_post_" + f])
            assert e <= RTOL_STAGE, (f, e)
        # rows without an image: untouched, bit for bit
        touched = np.zeros(len(pre), bool)
        touched[z["bc_img_row"]] = True
        for f, pf in FIELDS + (("coord", "coords"), ("smlen", "smlen")):
            assert np.array_equal(out[pf][~touched], pre[f][~touched]), f
        # the call is idempotent on an unchanged fluid state (images read only real and pressure-ghost rows)
        assert s.apply_bcond() == 0
        again = s.download(ORDER_REFERENCE)
        for f, pf in FIELDS:
            assert np.array_equal(again[pf], out[pf]), f
    finally:
        s.close()


def test_no_images_is_a_no_op():
    z = golden("plume3d_small")
    ocfg = op.config_for("plume3d_small", z)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, z, "b2")
        s.set_cells(attr)
        s.upload(dump_to_particles(z["bc_pre"]))
        s.set_images(np.zeros(0, psph.IMAGE_DTYPE))
        assert s.apply_bcond() == 0
        out = s.download(ORDER_REFERENCE)
        assert np.array_equal(out["rho"], z["bc_pre"]["rho"])
    finally:
        s.close()


def test_an_image_without_its_ghost_is_reported():
    """the reference stops with "ghost particle missing" (bcond.cc:59-66, return 3); here every such image counts into the
    invariant-violation counter, and the other images are applied as usual"""
    z = golden("plume3d_small")
    ocfg = op.config_for("plume3d_small", z)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, z, "b2")
        s.set_cells(attr)
        s.upload(dump_to_particles(z["bc_pre"]))
        img = images_of(ocfg, z)
        img["ghost_key"][:3, 2] += 12345                    # three images point at ghosts that do not exist
        s.set_images(img)
        assert s.apply_bcond() == 0
        assert s.errors() == 3
        assert s.errors() == 0                              # (read and reset)
    finally:
        s.close()
