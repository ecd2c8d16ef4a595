"""CPU (-m "not gpu"): host-side logic of the drop-in layer — TimeProps (properties.h:20-150),
simulation.data parsing (dataread.cc:109-235), synthetic case generator."""
import os
import numpy as np
import pytest

from plume_sph_b200.driver import TimeProps, stof
from plume_sph_b200 import cases


def test_timeprops_clips_to_output_and_end_times():
    tp = TimeProps(max_time=0.25, max_steps=100, timeoutput=0.1)
    dts = []
    while not tp.ifend():
        dts.append(tp.incrtime(0.03))
        tp.ifoutput()
    t = np.cumsum(dts)
    assert any(abs(t - 0.1) < 1e-15) and any(abs(t - 0.2) < 1e-15) and abs(t[-1] - 0.25) < 1e-15
    assert max(dts) <= 0.03 and min(dts) > 0
    assert tp.ioutput == 3  # 0.1, 0.2 and the end time


def test_timeprops_erupt_window_and_step_limit():
    tp = TimeProps(max_time=10, max_steps=3, timeoutput=100, stat_erupt=0.05, end_erupt=0.1)
    assert not tp.iferupt()
    tp.incrtime(0.06)
    assert tp.iferupt()
    tp.incrtime(0.06)
    assert not tp.iferupt()
    tp.incrtime(0.01); tp.incrtime(0.01)
    assert tp.ifend()  # step 4 > max_steps 3 (properties.h:92-95)


def test_stof_rounds_to_float():
    assert stof("0.17") == float(np.float32(0.17)) != 0.17
    assert stof("100") == 100.0


def test_cube_case_shape_and_eos():
    cfg, rows = cases.cube(8, seed=5)
    assert len(rows) == 24 ** 3
    real = rows["bc_type"] == 100
    assert real.sum() == 512 and (rows["bc_type"][~real] == 1).all()
    assert (rows["phase_num"][real] == 2).sum() > 0
    # pressure equilibrium between air and erupted material, EOS-consistent secondaries
    p = rows["pressure"]
    assert np.allclose(p, p[0], rtol=1e-12)
    assert len(np.unique(rows["key"], axis=0)) == len(rows)
    # every particle lies inside the mesh with at least one spare bucket around it
    lo = np.array([cfg.mindom[k] for k in range(3)])
    hi = lo + np.array([cfg.ncell[k] for k in range(3)]) * cfg.delta
    assert (rows["coords"] > lo + cfg.delta - 1e-9).all() and (rows["coords"] < hi - cfg.delta + 1e-9).all()


def test_run_loop_calls_the_stages_in_the_reference_order():
    """driver.run_loop forwards the stage calls in the order of main.cc:217-486 (3-D): eruption injection after the
    re-bin (:414-417), the wall boundary condition at the end of the step (:468)"""
    from plume_sph_b200 import driver

    class Cfg:
        adaptive_sml, dim, sml_update_int, use_gsph, have_turbulence_lans = 1, 3, 2, 0, 1

    class Rec:
        cfg = Cfg()

        def __init__(self):
            self.calls = []

        def __getattr__(self, name):
            def f(*a):
                self.calls.append(name)
                return 0.25 if name == "timestep" else 0
            return f

    tp = driver.TimeProps(max_time=10.0, max_steps=100, timeoutput=100.0)
    s = Rec()
    assert driver.run_loop(s, tp, max_steps=2, erupt=True) == 2
    step1 = ["timestep", "set_dt", "search_neighs_consth", "mom_engr_update", "smooth_density", "smooth_velocity",
             "update_pos", "update_particles_in_bucket", "add_new_erupt", "apply_bcond"]
    step2 = step1[:3] + ["adaptive_sml"] + step1[3:]          # SML_UPDATE_INT = 2: every second step
    assert s.calls == step1 + step2
    s2 = Rec()
    driver.run_loop(s2, driver.TimeProps(max_time=10.0, max_steps=100, timeoutput=100.0, stat_erupt=5.0), max_steps=1, erupt=True)
    assert "add_new_erupt" not in s2.calls                     # outside the eruption window (properties.h iferupt)


@pytest.mark.parametrize("name,Ll,Lu,smlen,imin,imax", [
    ("plume3d_small", (-20800., -20800., 1666.), (20800., 20800., 24000.), 100.0, (-350., -350., 1666.), (350., 350., 2400.)),
    ("jpue3d", (-576., -576., 0.), (576., 576., 6000.), 8.0, (-60., -60., 0.), (60., 60., 100.))])
def test_background_mesh_matches_the_reference_buckets(name, Ll, Lu, smlen, imin, imax):
    """mesh.background_mesh (the lattice of preprocess.cc + Read_Grid, computed instead of read from funky*.h5) against
    the buckets of the unmodified reference at the start of step 1 (tests/golden/*.npz, `b0_*`): the same lattice, the
    same set of non-brief buckets, and bit-identical type, has_involved, bucket_index and bnd for every one of them"""
    import oracle_py as op
    from common import golden, cell_attrs
    from plume_sph_b200 import mesh
    z = golden(name)
    cfg = op.config_for(name, z)
    attr, lin = cell_attrs(cfg, z, "b0")
    mindom, n, d, mine = mesh.background_mesh(Ll, Lu, smlen, imin, imax)
    assert [float(x) for x in mindom] == [cfg.mindom[k] for k in range(3)]
    assert [int(x) for x in n] == [cfg.ncell[k] for k in range(3)] and d == cfg.delta
    assert set(np.flatnonzero(mine["type"] != 0).tolist()) == set(int(x) for x in lin)
    for f in ("type", "has_involved", "bucket_index", "bnd"):
        assert np.array_equal(mine[f][lin], attr[f][lin]), f


@pytest.mark.parametrize("name", ["plume3d_small", "jpue3d"])
def test_wall_ghost_images_match_the_reference_table(name):
    """mesh.wall_ghost_images (search_bnd_images for a flat ground in closed form, bnd_image.cc:27-270) against the
    Image_table the unmodified reference held at the end of step 1: the same wall ghosts get an image, with bit-identical
    mirror points, inside the same buckets"""
    import oracle_py as op
    from common import golden, cell_attrs, image_cells, dump_to_particles
    from plume_sph_b200 import mesh
    z = golden(name)
    cfg = op.config_for(name, z)
    attr, _ = cell_attrs(cfg, z, "b2")
    rows = dump_to_particles(z["bc_pre"])
    img, sel = mesh.wall_ghost_images(rows, [cfg.mindom[k] for k in range(3)], [cfg.ncell[k] for k in range(3)], cfg.delta, attr)
    o = np.argsort(z["bc_img_row"])
    assert np.array_equal(sel, z["bc_img_row"][o])
    assert np.array_equal(img["coord"], z["bc_img_coord"][o])
    assert np.array_equal(img["cell"], image_cells(cfg, z)[o])
    assert np.array_equal(img["ghost_key"], z["bc_pre"]["key"][sel])


def test_default_lattice_has_the_bucket_count_of_the_reference_run():
    """the default 3-D case (bin/simulation.data: smoothing length 100, parameters.h box) has 93 775 buckets
    (counted in the reference run); 55 x 55 x 31 from preprocess.cc:99-133"""
    from plume_sph_b200 import mesh
    mindom, n, d = mesh.lattice((-20800., -20800., 1666.), (20800., 20800., 24000.), 100.0)
    assert [int(x) for x in n] == [55, 55, 31] and int(np.prod(n)) == 93775 and d == 800.0
    assert [float(x) for x in mindom] == [-22000.0, -22000.0, 466.0]
