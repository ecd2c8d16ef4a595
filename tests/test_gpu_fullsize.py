"""-m gpu: the hot path at BASELINE's full single-GPU size (216^3 = 1.0e7 real particles + ghost shell, the
bench.py workload) checked through properties that do not depend on the size (no oracle run of that size is
affordable: the reference needs ~30 min per 1e6-particle step):

on the UNJITTERED lattice with all particles at rest, equal masses and uniform pressure
  * smooth_density must give every real particle the same density, the lattice sum
        rho = m * sum_{d in {-3..3}^3} pi^-3/2 exp(-|d|^2)          (h = lattice spacing = 1, box cutoff 3 h)
    computed here in 343 terms — pressure ghosts contribute like real particles (PGHOST_CONTRIBUTE_DES = 1);
  * mom_engr_update's pair forces cancel by symmetry: velocities stay zero and, away from the erupted sphere,
    the specific energy keeps its initial value; smooth_velocity returns zero; update_pos moves nothing;
  * every real particle is updated, none is lost, no invariant is violated, and the CFL dt is the analytic one.
The run goes through the default pair path (block tiles) at the size the bench is quoted on.
"""
import numpy as np
import pytest

import plume_sph_b200 as psph
from plume_sph_b200 import cases
from plume_sph_b200.solver import ORDER_DEVICE

pytestmark = pytest.mark.gpu


def lattice_sum():
    d = np.arange(-3, 4, dtype=np.float64)
    X, Y, Z = np.meshgrid(d, d, d, indexing="ij")
    return float((np.pi ** -1.5 * np.exp(-(X * X + Y * Y + Z * Z))).sum())


@pytest.mark.parametrize("n", [216])
def test_lattice_known_answers_at_full_size(n):
    cfg, rows = cases.cube(n, seed=1, jitter=0.0, vjit=0.0)
    real = rows["bc_type"] == 100
    c_air = float(rows["sound_speed"][real & (rows["phase_num"] == 1)][0])
    h = float(rows["smlen"][0])
    s = psph.Solver(cfg)
    s.upload(rows)
    assert s.num_updated == n ** 3
    dt, lost, bad = s.step(0.0, 0)
    assert lost == 0 and bad == 0
    c_max = float(rows["sound_speed"][real].max())
    assert dt == cfg.cfl * (h / c_max)                      # timestep.cc:95-135: CFL * min h / c
    out = s.download(ORDER_DEVICE)
    s.close()
    assert len(out) == len(rows)
    # device order != input order: match by the birth key
    def key64(a):
        return (a["key"][:, 0].astype(np.uint64) << np.uint64(32)) | a["key"][:, 1].astype(np.uint64)
    out = out[np.argsort(key64(out))]
    rows = rows[np.argsort(key64(rows))]
    real = rows["bc_type"] == 100
    # density: one number for all 1.0e7 real particles
    rho_ref = cases.RHO0 * h ** 3 * lattice_sum() / h ** 3
    rel = np.abs(out["rho"][real] / rho_ref - 1.0).max()
    assert rel <= 2e-13, rel
    # forces cancel: velocities (and the filtered ones) stay at rest relative to the sound speed
    assert np.abs(out["vel"][real]).max() <= 1e-9 * c_air
    assert np.abs(out["smoothed_v"][real]).max() <= 1e-9 * c_air
    assert np.abs(out["coords"][real] - rows["coords"][real]).max() <= 1e-12 * h
    # energy: untouched away from the erupted sphere (whose energy differs and is smoothed across its surface)
    ctr = 0.5 * n * h
    r = np.sqrt(((rows["coords"] - ctr) ** 2).sum(1))
    far = real & (r > n * h / 8.0 + 4.0 * h)
    assert np.abs(out["engr"][far] / rows["engr"][far] - 1.0).max() <= 2e-13
    # ghosts are static
    gh = ~real
    assert np.array_equal(out["coords"][gh], rows["coords"][gh]) and np.array_equal(out["rho"][gh], rows["rho"][gh])
