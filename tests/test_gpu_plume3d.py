"""-m gpu: CUDA path (through the C ABI) against the reference's own dumps of the 3-D cases
(tests/golden/plume3d_small.npz, jpue3d.npz, and plume3d_small_gsph.npz = the plume built with USE_GSPH 1: 3-D gradients,
reconstructed Riemann states, Roe solver; produced by the unmodified reference — see tests/golden/make_golden.py):
every stage of time step 1.

Bars: neighbour lists, particle ordering, re-binning, flags: bit-exact.  Fields: relative
tolerance RTOL_STAGE (tests/common.py) per stage, stated against the field's max magnitude.
"""
import numpy as np
import pytest

import oracle_py as op
from common import (golden, to_product_cfg, dump_to_particles, cell_attrs, relerr, match_rows, RTOL_STAGE, D2P)
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE, ORDER_DEVICE

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["plume3d_small", "jpue3d", "plume3d_small_gsph"])
def case(request):
    z = golden(request.param)
    ocfg = op.config_for(request.param, z)
    cfg = to_product_cfg(ocfg)
    s = psph.Solver(cfg)
    attr, _ = cell_attrs(ocfg, z)
    s.set_cells(attr)
    s.upload(dump_to_particles(z["pre"]))
    yield z, ocfg, s
    s.close()


def check_stage(s, z, stage, fields, exact=()):
    out = s.download(ORDER_REFERENCE)
    pre = z["pre"]
    assert np.array_equal(out["key"], pre["key"])
    worst = {}
    for f in fields:
        ref = z["s1_%s_%s" % (stage, f)]
        got = out[D2P[f]]
        if f in exact:
            assert np.array_equal(got, ref), (stage, f)
        else:
            worst[f] = relerr(got, ref)
            assert worst[f] <= RTOL_STAGE, (stage, f, worst[f])
    return worst


def test_step1_all_stages(case):
    z, ocfg, s = case
    pre = z["pre"]
    # (1) particle ordering: the REFERENCE-order download reproduces THashTable iteration order
    out = s.download(ORDER_REFERENCE)
    assert np.array_equal(out["key"], pre["key"])
    assert np.array_equal(out["coords"], pre["coord"])
    # timestep: CFL dt, bit-exact (a min of exact quotients)
    dt = s.timestep()
    assert dt == float(z["dt"][0])
    s.set_dt(dt, 1, 1)
    # (2) neighbour search: list lengths of every particle + full sets of the sampled ones
    s.search_neighs_consth()
    off, idx, rows = s.neighbours()
    dev2ref = match_rows(pre, rows)            # device row -> row of `pre`
    counts = np.zeros(len(pre), np.int64)
    counts[dev2ref] = np.diff(off)
    assert np.array_equal(counts, z["list_counts"])
    ref2dev = np.zeros(len(pre), np.int64)
    ref2dev[dev2ref] = np.arange(len(pre))
    loff, lidx = z["list_off"], z["list_idx"]
    for q, r in enumerate(z["list_rows"]):
        d = ref2dev[r]
        mine = np.sort(dev2ref[idx[off[d]:off[d + 1]]])
        assert np.array_equal(mine, lidx[loff[q]:loff[q + 1]]), r
    # (3) the summation stages (a USE_GSPH build first forms the gradients the Riemann states are reconstructed from:
    # calc_gradients, main.cc:341-350; every non-guest row, ghosts included)
    if ocfg.use_gsph:
        s.calc_gradients()
        check_stage(s, z, "grad", ["grad"])
    s.mom_engr_update()
    check_stage(s, z, "mom", ["vel", "engr", "pressure", "sound_speed", "temperature", "specific_heat_p", "gamma",
                              "phase_density"])
    s.smooth_density()
    check_stage(s, z, "dens", ["rho", "mass_frac", "engr", "pressure", "sound_speed", "temperature", "involved"],
                exact=("involved",))
    s.smooth_velocity()
    check_stage(s, z, "svel", ["smoothed_v"])
    # (4) integrator + re-bin
    s.update_pos()
    check_stage(s, z, "pos", ["coord", "bc_type", "smlen", "involved"], exact=("bc_type", "smlen", "involved"))
    lost = s.update_particles_in_bucket()
    assert lost == 0
    assert s.errors() == 0
    cells, rows = s.cells()
    dev2ref = match_rows(pre, rows)
    got = np.zeros(len(pre), np.int64)
    got[dev2ref] = cells
    attr, lin = cell_attrs(ocfg, z, "b1")
    want = np.full(len(pre), -1, np.int64)
    boff, brows = z["b1_off"], z["b1_rows"]
    for b, c in enumerate(lin):
        want[brows[boff[b]:boff[b + 1]]] = c
    assert np.array_equal(got, want)


def test_fused_step_equals_staged(case):
    """psph_step (no host round trip between stages) == the staged calls above"""
    z, ocfg, s = case
    staged = s.download(ORDER_REFERENCE)
    cfg = to_product_cfg(ocfg)
    s2 = psph.Solver(cfg)
    attr, _ = cell_attrs(ocfg, z)
    s2.set_cells(attr)
    s2.upload(dump_to_particles(z["pre"]))
    dt, lost, ierr = s2.step(0.0, 1)
    assert dt == float(z["dt"][0]) and lost == 0 and ierr == 0
    fused = s2.download(ORDER_REFERENCE)
    for f in ("coords", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure", "smlen", "bc_type"):
        assert np.array_equal(fused[f], staged[f]), f
    s2.close()


def test_sph_rows_round_trip():
    """psph_upload_rows / psph_download_rows with PSPH_ROW_BYTES_SPH rows (psph_particle without the GSPH gradient block,
    what an SPH build of the reference exchanges, pack_data.h:68-75): the same table as the full rows, in both orders"""
    z = golden("plume3d_small")
    ocfg = op.config_for("plume3d_small", z)
    full = dump_to_particles(z["pre"])
    sph = np.zeros(len(full), psph.PARTICLE_SPH_DTYPE)
    for f in psph.PARTICLE_SPH_DTYPE.names:
        sph[f] = full[f]
    a, b = psph.Solver(to_product_cfg(ocfg)), psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, z)
        for s, rows in ((a, full), (b, sph)):
            s.set_cells(attr)
            s.upload(rows)
        b.append(sph[:0])
        for order in (ORDER_REFERENCE, ORDER_DEVICE):
            fa, fb = a.download(order), b.download(order)
            ca, cb = a.download(order, dtype=psph.PARTICLE_SPH_DTYPE), b.download(order, dtype=psph.PARTICLE_SPH_DTYPE)
            assert len(fa) == len(fb) == len(ca) == len(cb) == len(full)
            for f in psph.PARTICLE_SPH_DTYPE.names:
                assert np.array_equal(fa[f], fb[f]) and np.array_equal(fa[f], ca[f]) and np.array_equal(fa[f], cb[f]), f
            assert not fb["grad"].any()
        # one step from either upload: identical tables
        ra, rb = a.step(0.0, 1), b.step(0.0, 1)
        assert ra == rb
        fa, cb = a.download(ORDER_REFERENCE), b.download(ORDER_REFERENCE, dtype=psph.PARTICLE_SPH_DTYPE)
        for f in psph.PARTICLE_SPH_DTYPE.names:
            assert np.array_equal(fa[f], cb[f]), f
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("name", ["plume3d_small", "jpue3d"])
def test_h5part_columns_match_the_writer(name):
    """psph_download_h5part == what write_h5part (h5part_writer.cc:40-269) would put into a Step group from this table:
    the reference's own dump is in THashTable iteration order (the writer's loop), the writer keeps the non-guest real
    rows; and == the full-row download in reference order, filtered the same way, after a time step"""
    z = golden(name)
    ocfg = op.config_for(name, z)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, z)
        s.set_cells(attr)
        pre = z["pre"]
        rng = np.random.default_rng(3)
        s.upload(dump_to_particles(pre)[rng.permutation(len(pre))])     # the table order must not depend on the upload order
        cols = s.download_h5part()
        w = (pre["bc_type"] == 100) & (pre["guest"] == 0)
        assert len(cols["x"]) == int(w.sum()) and np.array_equal(cols["ID"], np.arange(int(w.sum())))
        for k, (f, c) in dict(x=("coord", 0), y=("coord", 1), z=("coord", 2), Vx=("vel", 0), Vy=("vel", 1), Vz=("vel", 2)).items():
            assert np.array_equal(cols[k], pre[f][w][:, c]), k
        for k, f in dict(Rho="rho", engr="engr", mssfrc="mass_frac", prss="pressure", phase="phase", bctp="bc_type", guest="guest").items():
            assert np.array_equal(cols[k], pre[f][w]), k
        allc = s.download_h5part(with_ghosts=True)                      # a WRITE_GHOSTS build writes every particle
        assert np.array_equal(allc["x"], pre["coord"][:, 0]) and np.array_equal(allc["bctp"], pre["bc_type"])
        dt, lost, ierr = s.step(0.0, 1)
        assert ierr == 0
        rows = s.download(ORDER_REFERENCE)
        w = (rows["bc_type"] == 100) & (rows["guest"] == 0)
        cols = s.download_h5part()
        assert np.array_equal(cols["z"], rows["coords"][w][:, 2]) and np.array_equal(cols["prss"], rows["pressure"][w])
        assert np.array_equal(cols["Vy"], rows["vel"][w][:, 1]) and np.array_equal(cols["mssfrc"], rows["mass_frac"][w])
    finally:
        s.close()
