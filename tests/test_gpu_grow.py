"""-m gpu: domain growth on the device against the reference's own calls in the small plume
(tests/golden/grow3d.npz, produced by the unmodified reference: tests/golden/make_golden.py::golden_grow):

  psph_scan_outside_layer   scan_outside_layer.cc:44-121  (end of step 1: the vent's bucket becomes has_involved = 3, adapt = 1)
  psph_adapt_domain         adapt_domain.cc:41-194        (start of step 2: 6144 pressure ghosts turn real, bucket flags)
  psph_add_pressure_ghost   add_pressure_ghost.cc:36-316  (20480 new pressure ghosts whose state comes from the Meteo sounding)
  psph_add_wall_ghost       add_wall_ghost.cc:44-293      (12288 new wall ghosts under the grown domain, after shift_brief_buck)

Everything here is integer / flag work or explicit IEEE arithmetic in the reference's order: the bar is BIT-EXACT — bucket
attributes, bc_type / involved of every row, and for the new rows keys, positions, mass, density, energy and every EOS output,
at the places the reference's particle table (THashTable iteration order) has them.
"""
import numpy as np
import pytest

import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, grow_cell_attrs, grow_cell_mincrd
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE

pytestmark = pytest.mark.gpu


def growth_record(z):
    g, mz, mrho, mp = op.growth_from_golden(z)
    import ctypes as C
    pg = psph.PsphGrowth()
    assert C.sizeof(pg) == C.sizeof(g)
    C.memmove(C.byref(pg), C.byref(g), C.sizeof(pg))
    return pg, mz, mrho, mp


@pytest.mark.parametrize("with_bounds", [True, False])
def test_domain_growth_matches_reference(with_bounds):
    z = golden("grow3d")
    ocfg = op.config_for("grow3d")
    cfg = to_product_cfg(ocfg)
    attr0, lin = grow_cell_attrs(ocfg, z, "scan_pre")
    table = z["table"]
    s = psph.Solver(cfg)
    try:
        s.set_cells(attr0)
        rng = np.random.default_rng(11)
        s.upload(dump_to_particles(table)[rng.permutation(len(table))])      # the result must not depend on the upload order
        s.search_neighs_consth()                                             # (rows are binned; the calls read I_CELL)

        def check(tag, active_of=None):
            want, _ = grow_cell_attrs(ocfg, z, tag)
            got = s.get_cells()
            for f in ("type", "has_involved", "flags"):
                assert np.array_equal(got[f], want[f]), (tag, f)
            assert np.array_equal(got["active"][lin], want["active"][lin]), (tag, "active")

        # ---- scan_outside_layer
        assert s.scan_outside_layer() == int(z["adapt"][0]) == 1
        check("scan_post")
        # (update_bgmesh ran between the scan and the adapt block: active flags of the fixture's next dump)
        a = s.get_cells()
        a["active"] = grow_cell_attrs(ocfg, z, "adapt_pre")[0]["active"]
        s.set_cells(a)
        # ---- adapt_domain
        assert s.adapt_domain() == len(z["adapt_rows"]) == 6144
        check("adapt_post")
        rows = s.download(ORDER_REFERENCE)
        assert np.array_equal(rows["key"], table["key"])
        bc, inv = table["bc_type"].copy(), table["involved"].copy()
        bc[z["adapt_rows"]] = z["adapt_bc_type"]
        inv[z["adapt_rows"]] = z["adapt_involved"]
        assert np.array_equal(rows["bc_type"], bc) and np.array_equal(rows["involved"], inv)
        for f, d in (("coords", "coord"), ("rho", "rho"), ("engr", "engr"), ("mass", "mass"), ("smlen", "smlen")):
            assert np.array_equal(rows[f], table[d]), f                      # nothing else moved
        # ---- add_pressure_ghost
        g, mz, mrho, mp = growth_record(z)
        s.set_growth(g, mz, mrho, mp, grow_cell_mincrd(ocfg, z, lin) if with_bounds else None)
        added = s.add_pressure_ghost(int(z["param_step"][0]))
        want = z["pghost_rows"]
        assert added == len(want) == 20480
        check("pghost_post")
        rows = s.download(ORDER_REFERENCE)
        assert len(rows) == len(table) + added
        pos = z["pghost_pos"]
        old = np.ones(len(rows), bool)
        old[pos] = False
        if with_bounds:
            # the new rows sit where the reference's table has them, bit for bit
            assert np.array_equal(rows["key"][old], table["key"])
            got = rows[pos]
            for f, d in (("key", "key"), ("bc_type", "bc_type"), ("phase_num", "phase"), ("involved", "involved"), ("guest", "guest"),
                         ("coords", "coord"), ("mass", "mass"), ("smlen", "smlen"), ("mass_frac", "mass_frac"), ("rho", "rho"),
                         ("engr", "engr"), ("smoothed_e", "smoothed_e"), ("pressure", "pressure"), ("gamma", "gamma"),
                         ("sound_speed", "sound_speed"), ("temperature", "temperature"), ("specific_heat_p", "specific_heat_p"),
                         ("phase_density", "phase_density"), ("vel", "vel"), ("smoothed_v", "smoothed_v")):
                assert np.array_equal(got[f], want[d]), f
        else:
            # without the driver's stored bounds the lattice sits on (mindom + i del) / len_scale: same particles up to the
            # few float ulps by which the reference's stored x bounds differ from that formula
            new = rows[rows["key"][:, 2] == int(z["param_step"][0])]
            new = new[new["bc_type"] == 1]
            assert len(new) == added
            a = new[np.lexsort(np.rint(new["coords"]).T[::-1])]
            b = want[np.lexsort(np.rint(want["coord"]).T[::-1])]
            assert np.abs(a["coords"] - b["coord"]).max() < 1e-3
            assert np.array_equal(a["rho"], b["rho"]) and np.array_equal(a["engr"], b["engr"])   # functions of z only
        # ---- add_wall_ghost, after the driver's shift_brief_buck (the fixture's bucket set at that point)
        attr_w, lin_w = grow_cell_attrs(ocfg, z, "wghost_pre")
        s.set_cells(attr_w)
        if with_bounds:
            s.set_growth(g, mz, mrho, mp, grow_cell_mincrd(ocfg, z, lin_w, "wghost_pre"))
        nwall = s.add_wall_ghost(int(z["param_step"][0]))
        wantw = z["wghost_rows"]
        assert nwall == len(wantw) == 12288
        want_a, _ = grow_cell_attrs(ocfg, z, "wghost_post")
        got_a = s.get_cells()
        for f in ("type", "has_involved", "flags", "active"):
            assert np.array_equal(got_a[f], want_a[f]), ("wghost_post", f)
        rows2 = s.download(ORDER_REFERENCE)
        assert len(rows2) == len(rows) + nwall
        if with_bounds:
            posw = z["wghost_pos"]
            oldw = np.ones(len(rows2), bool)
            oldw[posw] = False
            assert np.array_equal(rows2["key"][oldw], rows["key"])
            got = rows2[posw]
            for f, d in (("key", "key"), ("bc_type", "bc_type"), ("phase_num", "phase"), ("involved", "involved"), ("coords", "coord"),
                         ("mass", "mass"), ("smlen", "smlen"), ("rho", "rho"), ("engr", "engr"), ("pressure", "pressure"),
                         ("gamma", "gamma"), ("sound_speed", "sound_speed"), ("temperature", "temperature"),
                         ("specific_heat_p", "specific_heat_p"), ("phase_density", "phase_density"), ("vel", "vel")):
                assert np.array_equal(got[f], wantw[d]), ("wall ghost", f)
        else:
            assert int((rows2["bc_type"] == 2).sum()) == int((rows["bc_type"] == 2).sum()) + nwall
        # the grown table steps: new ghosts are binned, real rows see them
        dt, lost, ierr = s.step(0.0, 1)
        assert ierr == 0 and dt > 0
    finally:
        s.close()


def test_whole_adapt_block_matches_the_reference():
    """The adapt block of main.cc:256-304 end to end — psph_adapt_domain, psph_add_pressure_ghost, [shift_brief_buck on the host],
    psph_add_wall_ghost, [update_bgmesh and the image search on the host], psph_apply_bcond — as plume_sph_b200.driver.grow_domain
    runs it, from the reference's table at the end of step 1 (grow3d) to the table the reference starts the stages of step 2 from
    (plume3d_grow `pre2`): the same 52 112 particles in the reference's table order; flags, positions and every field of the
    non-wall rows bit for bit; the wall ghosts' interpolated states (apply_bcond over old AND new ghosts) within RTOL_STAGE."""
    from common import relerr, RTOL_STAGE
    from plume_sph_b200 import mesh, driver
    z, g2 = golden("grow3d"), golden("plume3d_grow")
    ocfg = op.config_for("grow3d")
    cfg = to_product_cfg(ocfg)
    attr0, lin = grow_cell_attrs(ocfg, z, "scan_pre")
    _, lin_w = grow_cell_attrs(ocfg, z, "wghost_pre")
    s = psph.Solver(cfg)
    try:
        s.set_cells(attr0)
        s.upload(dump_to_particles(z["table"]))
        s.search_neighs_consth()
        assert s.scan_outside_layer() == 1                                   # main.cc:435 at the end of step 1
        a = s.get_cells()
        a["active"] = grow_cell_attrs(ocfg, z, "adapt_pre")[0]["active"]
        s.set_cells(a)
        Ll, Lu = z["param_domain"][:3], z["param_domain"][3:]
        mindom, n, d, latent = mesh.background_mesh(Ll, Lu, float(z["param_const"][1]), Ll - 1e9, Lu + 1e9)
        assert [int(x) for x in n] == [ocfg.ncell[k] for k in range(3)]
        g, mz, mrho, mp = growth_record(z)
        s.set_growth(g, mz, mrho, mp, grow_cell_mincrd(ocfg, z, lin_w, "wghost_pre"))
        tp = driver.TimeProps()
        tp.step = int(z["param_step"][0])
        assert driver.grow_domain(s, tp, dict(latent=latent, mindom=[float(x) for x in mindom], ncell=n, delta=d)) == 0
        assert s.errors() == 0
        rows = s.download(ORDER_REFERENCE)
        want = g2["pre2"]
        assert len(rows) == len(want) == 52112
        assert np.array_equal(rows["key"], want["key"])
        for f, w in (("bc_type", "bc_type"), ("involved", "involved"), ("phase_num", "phase"), ("coords", "coord"), ("mass", "mass"),
                     ("smlen", "smlen"), ("mass_frac", "mass_frac")):
            assert np.array_equal(rows[f], want[w]), f
        wall = want["bc_type"] == 2
        for f, w in (("rho", "rho"), ("engr", "engr"), ("pressure", "pressure"), ("vel", "vel"), ("sound_speed", "sound_speed"),
                     ("temperature", "temperature")):
            assert np.array_equal(rows[f][~wall], want[w][~wall]), f
            e = relerr(rows[f][wall], want[w][wall])
            assert e <= RTOL_STAGE, (f, e)
        # the host's bucket bookkeeping: the buckets of the reference after shift_brief_buck, with their attributes
        got_a = s.get_cells()
        want_a, _ = grow_cell_attrs(ocfg, z, "wghost_post")
        for f in ("type", "has_involved", "flags"):
            assert np.array_equal(got_a[f], want_a[f]), f
    finally:
        s.close()
