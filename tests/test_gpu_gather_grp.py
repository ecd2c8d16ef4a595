"""-m gpu: the grouped pair gather (csrc/psph_gather3.cuh, the default of the 3-D SPH stages) against
the oracle and against the per-row gather (csrc/psph_gather.cuh, PSPH_GATHER=row) on inputs that
reach its rare paths:
  * a dense cube (h = 1.7 lattice spacings, ~1000 pairs per particle): queues fill up before the scan of
    a group ends -> early drains with parked partial sums;
  * two smoothing lengths with hmax > 1.83 h_i: the neighbour-list / in_support limits bite, so every
    candidate of the small-h rows goes through the FP64 gates;
  * positions placed exactly ON and one ulp either side of a cutoff face: the FP32 pre-filter must hand
    them to the FP64 gates (the pair set is bit-exact, so a wrong decision shows up as an O(1e-6) error).
Both gathers are CUDA; the oracle (CPU) is the checker.
"""
import os
import numpy as np
import pytest

import plume_sph_b200 as psph
from plume_sph_b200 import cases
from plume_sph_b200.solver import ORDER_REFERENCE
from common import to_oracle_cfg, particles_to_oracle, relerr, RTOL_STAGE

pytestmark = pytest.mark.gpu

FIELDS = ("vel", "engr", "rho", "mass_frac", "smoothed_v", "pressure")


SWITCHES = ("PSPH_GATHER", "PSPH_NBR", "PSPH_SCAN", "PSPH_WCACHE", "PSPH_DENSNEAR")


def run_stages(cfg, rows, mode, nbr=None, env=None):
    """mode: None (default gathers) or "row" (per-row gather); nbr: None (default pair lists), "0" (fused grouped
    gather only), "1" / "3" (row pair lists: uniform h / any h), "2" (block-tiled slot lists, csrc/psph_tile.cuh);
    env: further switches of the block-tiled path (PSPH_SCAN, PSPH_WCACHE, PSPH_DENSNEAR).  The switches are read by
    psph_create."""
    old = {k: os.environ.get(k) for k in SWITCHES}
    want = dict(env or {})
    want.update({"PSPH_GATHER": mode, "PSPH_NBR": nbr})
    for k in SWITCHES:
        v = want.get(k)
        if v:
            os.environ[k] = v
        else:
            os.environ.pop(k, None)
    try:
        s = psph.Solver(cfg)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    s.upload(rows)
    dt = s.timestep()
    s.set_dt(dt, 1, 0)
    s.search_neighs_consth()
    s.mom_engr_update()
    s.smooth_density()
    s.smooth_velocity()
    bad = s.errors()
    out = s.download(ORDER_REFERENCE)
    s.close()
    return dt, bad, out


def oracle_stages(cfg, rows):
    o = particles_to_oracle(to_oracle_cfg(cfg), rows)
    o.bin()
    dt = o.timestep()
    o.search()
    o.mom_engr_update(dt)
    o.smooth_density()
    o.smooth_velocity()
    perm = o.table_order()
    return dt, {"vel": o.a["v"][perm], "engr": o.a["e"][perm], "rho": o.a["rho"][perm], "mass_frac": o.a["ks"][perm],
                "smoothed_v": o.a["sv"][perm], "pressure": o.a["p"][perm]}


def dense_rows(n=20, hfac=1.7, two_h=False, seed=5):
    cfg, rows = cases.cube(n, seed=seed)
    real = rows["bc_type"] == 100
    rows["smlen"] = hfac
    if two_h:  # erupted material carries a smaller smoothing length (sml_of_phase2)
        rows["smlen"][rows["phase_num"] == 2] = 0.45 * hfac
        rows["smlen"][real & (rows["coords"][:, 0] < 6.0)] = 0.5 * hfac
    return cfg, rows


@pytest.mark.parametrize("two_h", [False, True])
def test_dense_cube_grouped_vs_row_vs_oracle(two_h):
    cfg, rows = dense_rows(two_h=two_h)
    dt_g, bad_g, grp = run_stages(cfg, rows, None)
    dt_r, bad_r, row = run_stages(cfg, rows, "row")
    assert dt_g == dt_r and bad_g == 0 and bad_r == 0
    for f in FIELDS:
        assert relerr(grp[f], row[f]) <= RTOL_STAGE, (f, relerr(grp[f], row[f]))
    dt_o, ora = oracle_stages(cfg, rows)
    assert dt_o == dt_g
    for f in FIELDS:
        assert relerr(grp[f], ora[f]) <= 3 * RTOL_STAGE, (f, relerr(grp[f], ora[f]))  # three stages in a row


@pytest.mark.parametrize("nbr", ["0", "1", "3"])
def test_other_pair_paths_vs_default(nbr):
    """the fused grouped gather alone (PSPH_NBR=0), the uniform-h row lists (1) and the general tagged row lists (3)
    against the default (block tiles) on the synthetic cube with its erupted sphere, and on the two-h dense cube
    (mixed smoothing lengths, list limits that bite, queue overflow)"""
    for cfg, rows in (cases.cube(24, seed=9), dense_rows(n=16, two_h=True)):
        dt_a, bad_a, a = run_stages(cfg, rows, None)
        dt_b, bad_b, b = run_stages(cfg, rows, None, nbr=nbr)
        assert dt_a == dt_b and bad_a == 0 and bad_b == 0
        for f in FIELDS:
            assert relerr(a[f], b[f]) <= RTOL_STAGE, (nbr, f, relerr(a[f], b[f]))


def near_rows(n=24, seed=13):
    """the synthetic cube with partners INSIDE the energy-smoothing kernel (support 0.75 h): 300 particles moved to
    0.2-0.6 h from a lattice neighbour (one near partner each way), and two clusters of 10 particles within 0.4 h
    of one another (more near partners than a row's near list holds: smooth_density must fall back to the list
    walk for the step)"""
    cfg, rows = cases.cube(n, seed=seed)
    real = np.flatnonzero(rows["bc_type"] == 100)
    rng = np.random.default_rng(seed)
    pos = rows["coords"]
    pick = rng.choice(real, 320, replace=False)
    for p in pick[:300]:
        q = real[np.argmin(np.abs(pos[real] - (pos[p] + [1.0, 0.0, 0.0])).sum(axis=1))]
        pos[p] = pos[q] + rng.uniform(-1, 1, 3) * 0.35
    return cfg, rows, pick[300:]


@pytest.mark.parametrize("cluster", [False, True])
def test_near_partners_and_path_switches(cluster):
    """block-tiled path with each of its shortcuts switched off in turn (lane-per-candidate search, no pair-weight
    cache, smooth_density walking the lists) against the default and the oracle, on inputs with partners inside
    the energy-smoothing kernel; cluster=True overflows the near lists (the device falls back by itself)"""
    cfg, rows, spare = near_rows()
    if cluster:
        pos = rows["coords"]
        rng = np.random.default_rng(4)
        for c in range(2):
            centre = pos[spare[10 * c]].copy()
            for p in spare[10 * c:10 * c + 10]:
                pos[p] = centre + rng.uniform(-1, 1, 3) * 0.2
    dt_o, ora = oracle_stages(cfg, rows)
    dt_a, bad_a, a = run_stages(cfg, rows, None)
    assert bad_a == 0 and dt_a == dt_o
    for f in FIELDS:
        assert relerr(a[f], ora[f]) <= 3 * RTOL_STAGE, (f, relerr(a[f], ora[f]))
    for env in ({"PSPH_SCAN": "1"}, {"PSPH_WCACHE": "0"}, {"PSPH_DENSNEAR": "0"}):
        dt_b, bad_b, b = run_stages(cfg, rows, None, env=env)
        assert bad_b == 0 and dt_b == dt_a
        for f in FIELDS:
            assert relerr(a[f], b[f]) <= RTOL_STAGE, (env, f, relerr(a[f], b[f]))


def test_pairs_on_the_cutoff_face():
    """neighbours exactly at |dx| = 3 h, and one ulp inside / outside, along every axis"""
    cfg, rows = cases.cube(16, seed=3, jitter=0.0, vjit=0.02)
    real = np.flatnonzero(rows["bc_type"] == 100)
    rng = np.random.default_rng(11)
    pos = rows["coords"]
    # unjittered lattice: every real particle already has neighbours at exactly 3 h along each axis
    # (|dx|/h == 3 passes the reference's test `fabs(dx/h) > 3`); move a sample by +-1 ulp / +-1e-7 of it
    pick = rng.choice(real, 600, replace=False)
    for q, p in enumerate(pick):
        k = q % 3
        x = pos[p, k]
        pos[p, k] = [np.nextafter(x, np.inf), np.nextafter(x, -np.inf), x + 3e-8, x - 3e-8][(q // 3) % 4]
    dt_g, bad_g, grp = run_stages(cfg, rows, None)
    dt_o, ora = oracle_stages(cfg, rows)
    assert bad_g == 0 and dt_o == dt_g
    for f in FIELDS:
        assert relerr(grp[f], ora[f]) <= 3 * RTOL_STAGE, (f, relerr(grp[f], ora[f]))


# ---- per-row smoothing lengths: the sub-tiled mixed-h gather (csrc/psph_mix.cuh) ----------------------------------------
def mixed_h_rows(n=24, seed=21, spread=0.03, hfac=1.5, cluster=0):
    """the synthetic cube as adaptive_sml leaves it: every real particle with its own h = hfac (1 +- spread) lattice
    spacings (the ghost shell keeps h = 1); cluster > 0: that many extra-close particles around one site, so that their
    near lists (partners inside the energy-smoothing kernel, 0.75 h) overflow"""
    cfg, rows = cases.cube(n, seed=seed)
    real = np.flatnonzero(rows["bc_type"] == 100)
    rng = np.random.default_rng(seed)
    rows["smlen"][real] = hfac * (1.0 + spread * rng.uniform(-1, 1, len(real)))
    if cluster:
        pos = rows["coords"]
        pick = rng.choice(real, cluster, replace=False)
        centre = pos[pick[0]].copy()
        for p in pick:
            pos[p] = centre + rng.uniform(-1, 1, 3) * 0.3
    return cfg, rows


@pytest.mark.parametrize("cluster", [0, 60])
def test_mixed_h_subtiled_gather_vs_oracle(cluster):
    """default path on a table with per-row h (the device picks the mixed-h gather: StepState::pair_mode 4), the same
    forced (PSPH_NBR=4), and the fused grouped gather alone (PSPH_NBR=0), against the oracle; cluster: 60 particles
    within 0.3 spacings of one site — more near partners than a near list holds, those rows are summed by themselves"""
    cfg, rows = mixed_h_rows(cluster=cluster)
    dt_o, ora = oracle_stages(cfg, rows)
    runs = {}
    for nbr in (None, "4", "0"):
        dt, bad, out = run_stages(cfg, rows, None, nbr=nbr)
        assert bad == 0 and dt == dt_o
        for f in FIELDS:
            assert relerr(out[f], ora[f]) <= 3 * RTOL_STAGE, (nbr, f, relerr(out[f], ora[f]))
        runs[nbr] = out
    for f in FIELDS:   # default == forced mixed-h path, bit for bit (same kernels)
        assert np.array_equal(runs[None][f], runs["4"][f]), f


def test_mixed_h_pairs_on_the_cutoff_face():
    """unjittered lattice with h = 4/3 spacings for every real particle: partners sit exactly on the cutoff faces
    3 (h_i + h_j) / 2 = 4 spacings and 3 h_i = 4 spacings; a sample is moved one ulp / 3e-8 either side.  The FP32
    pre-filter of the mixed-h search must hand these to the FP64 gates (a wrong decision is an O(1e-6) error)."""
    cfg, rows = cases.cube(16, seed=3, jitter=0.0, vjit=0.02)
    real = np.flatnonzero(rows["bc_type"] == 100)
    rows["smlen"][real] = 4.0 / 3.0
    rng = np.random.default_rng(12)
    rows["smlen"][real[::7]] = 4.0 / 3.0 * (1 + 1e-9)      # a second length, so that pairs differ between the two gate families
    pos = rows["coords"]
    pick = rng.choice(real, 600, replace=False)
    for q, p in enumerate(pick):
        k = q % 3
        x = pos[p, k]
        pos[p, k] = [np.nextafter(x, np.inf), np.nextafter(x, -np.inf), x + 3e-8, x - 3e-8][(q // 3) % 4]
    dt_o, ora = oracle_stages(cfg, rows)
    for nbr in ("4", "0"):
        dt, bad, out = run_stages(cfg, rows, None, nbr=nbr)
        assert bad == 0 and dt == dt_o
        for f in FIELDS:
            assert relerr(out[f], ora[f]) <= 3 * RTOL_STAGE, (nbr, f, relerr(out[f], ora[f]))
