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


def test_run_loop_grows_the_domain_in_the_reference_order():
    """with growth on, run_loop scans the outer layer every step (main.cc:435) and runs the adapt block at the start of the
    step after the scan raised `adapt` (:256-304), in the reference's order: adapt_domain, add_pressure_ghost,
    [shift_brief_buck], add_wall_ghost, [update_bgmesh, image search], apply_bcond"""
    from plume_sph_b200 import driver, mesh
    import plume_sph_b200 as psph

    class Cfg:
        adaptive_sml, dim, sml_update_int, use_gsph, have_turbulence_lans = 1, 3, 50, 0, 1

    n = [3, 3, 3]

    class Rec:
        cfg = Cfg()

        def __init__(self):
            self.calls, self.scans = [], 0

        def get_cells(self):
            self.calls.append("get_cells")
            a = np.zeros(27, psph.CELL_ATTR_DTYPE)
            a["type"] = 12
            return a

        def cells(self):
            self.calls.append("cells")
            return np.zeros(4, np.int32), np.zeros(4, psph.PARTICLE_DTYPE)

        def download(self):
            self.calls.append("download")
            return np.zeros(4, psph.PARTICLE_DTYPE)

        def scan_outside_layer(self):
            self.calls.append("scan_outside_layer")
            self.scans += 1
            return 1 if self.scans == 1 else 0

        def __getattr__(self, name):
            def f(*a):
                self.calls.append(name)
                return 0.25 if name == "timestep" else 0
            return f

    _, _, _, latent = mesh.background_mesh((0., 0., 0.), (24., 24., 24.), 1.0, (-1e9,) * 3, (1e9,) * 3)
    growth = dict(latent=np.resize(latent, 27), mindom=[0., 0., 0.], ncell=n, delta=8.0)
    s = Rec()
    assert driver.run_loop(s, driver.TimeProps(max_time=10.0, max_steps=100, timeoutput=100.0), max_steps=2, growth=growth) == 2
    stages = ["set_dt", "search_neighs_consth", "mom_engr_update", "smooth_density", "smooth_velocity", "update_pos",
              "update_particles_in_bucket", "scan_outside_layer", "apply_bcond"]
    block = ["adapt_domain", "add_pressure_ghost", "get_cells", "set_cells", "add_wall_ghost", "get_cells", "cells", "set_cells",
             "set_images", "apply_bcond"]
    assert s.calls == ["timestep"] + stages + ["timestep"] + block + stages


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


def test_shift_brief_buck_matches_the_reference():
    """mesh.shift_brief_buck against the buckets the unmodified reference created in the growth step of the small plume
    (tests/golden/grow3d.npz: bucket dumps before and after shift_brief_buck): the same 120 cells, with bit-identical
    type, bucket_index and bnd (the closed-form attributes of mesh.background_mesh), not involved, no particle flags"""
    import oracle_py as op
    from common import golden, grow_cell_attrs
    from plume_sph_b200 import mesh
    z = golden("grow3d")
    cfg = op.config_for("grow3d")
    n = [cfg.ncell[k] for k in range(3)]
    before, l0 = grow_cell_attrs(cfg, z, "pghost_post")
    after, l1 = grow_cell_attrs(cfg, z, "wghost_pre")
    Ll, Lu = z["param_domain"][:3], z["param_domain"][3:]
    _, _, _, latent = mesh.background_mesh(Ll, Lu, float(z["param_const"][1]), Ll - 1e9, Lu + 1e9)   # every cell as a bucket
    new = mesh.shift_brief_buck(before, latent, n)
    assert np.array_equal(np.sort(new), np.setdiff1d(l1, l0)) and len(new) == 120
    for f in ("type", "has_involved", "flags", "bucket_index", "bnd", "active"):
        assert np.array_equal(before[f], after[f]), f


def test_update_bgmesh_rule():
    """mesh.update_bgmesh: active = a neighbour bucket holds a particle (mesh_update.cc:50-63).  (No reference pin: in the
    small-plume fixture the reference's own call stops half way with "neigh-bucket missing")"""
    from plume_sph_b200 import mesh
    import plume_sph_b200 as psph
    n = [4, 4, 3]
    attr = np.zeros(48, psph.CELL_ATTR_DTYPE)
    attr["type"] = 12
    attr["type"][0] = 0                                       # a brief bucket is never active
    cnt = np.zeros(48, np.int64)
    cnt[(1 * 4 + 1) * 3 + 1] = 5                              # one occupied bucket at (1, 1, 1)
    mesh.update_bgmesh(attr, cnt, n)
    idx = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    near = (np.abs(idx - 1).max(axis=1) <= 1) & ~((idx == 1).all(axis=1))
    want = near & (attr["type"] != 0)
    assert np.array_equal(attr["active"].astype(bool), want)


def test_default_lattice_has_the_bucket_count_of_the_reference_run():
    """the default 3-D case (bin/simulation.data: smoothing length 100, parameters.h box) has 93 775 buckets
    (counted in the reference run); 55 x 55 x 31 from preprocess.cc:99-133"""
    from plume_sph_b200 import mesh
    mindom, n, d = mesh.lattice((-20800., -20800., 1666.), (20800., 20800., 24000.), 100.0)
    assert [int(x) for x in n] == [55, 55, 31] and int(np.prod(n)) == 93775 and d == 800.0
    assert [float(x) for x in mindom] == [-22000.0, -22000.0, 466.0]


def test_bench_plume_setup_loader(tmp_path):
    """bench.py's plume leg: the state files the hooked reference driver writes (its own set-up of bin/simulation.data with a
    +-1000 m initial domain, 1 step) load into what the C ABI takes — rows, bucket attributes over the default lattice, the
    wall-ghost image table (every image points at an existing wall ghost, inside a bucket that exists) and the vent record"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oracle", "ref_build"))
    import ref_run
    if not ref_run.have_inputs():
        pytest.skip("oracle/_ref is not built in this checkout (python -c 'import __graft_entry__ as g; g.build()')")
    import bench
    import plume_sph_b200 as psph
    dump = str(tmp_path / "dump")
    ref_run.time_plume3d(1, 0, replicas=1, edits=ref_run.MEDIUM, keep=str(tmp_path / "setup"), dump=dump)
    cfg = psph.config_default_3d()
    rows, attr, img, er, boxes = bench.load_reference_setup(dump, psph, cfg)
    assert len(rows) == 52112 and int((rows["bc_type"] == 100).sum()) == 6912
    assert len(attr) == 55 * 55 * 31 and int((attr["type"] != 0).sum()) == 245
    keys = {tuple(k) for k in rows["key"][rows["bc_type"] == 2]}
    assert len(img) > 0 and all(tuple(k) in keys for k in img["ghost_key"])
    assert (attr["type"][img["cell"]] != 0).all()
    assert er.sml == 40.0 and er.rv == 120.0 and boxes.shape == (2, 6) and er.kmax[2] > er.kmin[2]
