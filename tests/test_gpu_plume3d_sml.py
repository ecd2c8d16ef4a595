"""-m gpu: the regime AFTER adaptive_sml — every particle with its own smoothing length — against the reference's own
dumps (tests/golden/plume3d_sml.npz: the small 3-D plume with adaptive_sml called every step, time steps 1 and 2; see
tests/golden/make_golden.py::golden_3d_sml).  A default 3-D run is in this regime from time step 50 on
(main.cc:319-322, SML_UPDATE_INT constant.h:189).

Step 1: adaptive_sml from uniform h, then mom_engr_update / smooth_density / smooth_velocity with the new lengths.
Step 2: starts from 675 distinct smoothing lengths among the real particles (and freshly converted real particles that
still carry the initial one: the domain grew), adaptive_sml again, all stages with mixed h.

Bars: RTOL_STAGE per stage against the field maximum; flags and bc_type exact.
"""
import numpy as np
import pytest

import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs, relerr, sparse_field, RTOL_STAGE, D2P
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE

pytestmark = pytest.mark.gpu

STAGES = (("sml", ["smlen"], ()),
          ("grad", ["grad"], ()),                      # USE_GSPH builds only (calc_gradients, main.cc:341-350)
          ("mom", ["vel", "engr", "pressure", "sound_speed", "temperature", "specific_heat_p", "gamma", "phase_density"], ()),
          ("dens", ["rho", "mass_frac", "engr", "pressure", "sound_speed", "temperature", "involved"], ("involved",)),
          ("svel", ["smoothed_v"], ()),
          ("pos", ["coord", "bc_type", "smlen", "involved"], ("bc_type", "involved")))


def run_step(z, step, name="plume3d_sml", sml_int=1):
    ocfg = op.config_for(name, z)
    cfg = to_product_cfg(ocfg)
    cfg.sml_update_int = sml_int
    pre = z["pre%d" % step]
    s = psph.Solver(cfg)
    try:
        attr, _ = cell_attrs(ocfg, z, "b%d" % step)
        s.set_cells(attr)
        s.upload(dump_to_particles(pre))
        dt = float(z["dt"][step - 1])
        if step == 1:
            assert s.timestep() == dt
        s.set_dt(dt, step, 1)
        s.search_neighs_consth()
        cur = {f: pre[f] for f in pre.dtype.names}
        calls = dict(sml=s.adaptive_sml, grad=s.calc_gradients, mom=s.mom_engr_update, dens=s.smooth_density, svel=s.smooth_velocity, pos=s.update_pos)
        worst = {}
        for stage, fields, exact in STAGES:
            if (stage == "sml" and sml_int != 1) or (stage == "grad" and not ocfg.use_gsph):
                continue
            calls[stage]()
            out = s.download(ORDER_REFERENCE)
            assert np.array_equal(out["key"], pre["key"])
            for f in fields:
                want = sparse_field(z, "s%d_%s_%s" % (step, stage, f), cur[f])
                got = out[D2P[f]]
                if f in exact:
                    assert np.array_equal(got, want), (step, stage, f)
                else:
                    worst[(stage, f)] = relerr(got, want)
                    assert worst[(stage, f)] <= RTOL_STAGE, (step, stage, f, worst[(stage, f)])
                cur[f] = want
        assert s.errors() == 0
        return worst
    finally:
        s.close()


@pytest.mark.parametrize("step", [1, 2])
def test_adaptive_sml_and_mixed_h_stages(step):
    z = golden("plume3d_sml")
    if step == 2:
        pre = z["pre2"]
        assert len(np.unique(pre["smlen"][pre["bc_type"] == 100])) > 100     # the regime under test
    worst = run_step(z, step)
    print("step", step, "worst relative errors:", {k: "%.1e" % v for k, v in worst.items()})


@pytest.mark.parametrize("step", [1, 2])
def test_gsph_3d_two_steps(step):
    """tests/golden/plume3d_gsph2.npz: the same two steps of the reference built with USE_GSPH 1 and its 3-D option set
    (options.h:544-663: Riemann states reconstructed from the gradients of calc_gradients with the cubic interface position,
    MINI_DERIVATIVE_COND 10, modified monotonicity, Roe solver; sph_lib.cc:2247-2600, :1667-1749).  Step 2 carries velocities,
    velocity gradients and per-particle smoothing lengths into the reconstruction."""
    z = golden("plume3d_gsph2")
    worst = run_step(z, step, name="plume3d_gsph2")
    print("step", step, "worst relative errors:", {k: "%.1e" % v for k, v in worst.items()})


@pytest.mark.parametrize("step", [1, 2])
def test_uniform_h_stages_after_domain_growth(step):
    """tests/golden/plume3d_grow.npz: the same two steps with the reference's own SML_UPDATE_INT (uniform h, block-tiled
    path).  Step 2 follows the adapt block of main.cc:257-303 (pressure ghosts turned real, new ghost layers and buckets)
    and keeps hundreds of partners exactly ON the cutoff faces of the unmoved lattice."""
    z = golden("plume3d_grow")
    worst = run_step(z, step, name="plume3d_grow", sml_int=50)
    print("step", step, "worst relative errors:", {k: "%.1e" % v for k, v in worst.items()})


def test_fused_step_crosses_adaptive_sml():
    """psph_step with sml_update_int = 1 from the state at the start of step 2 == the staged calls (bit for bit)"""
    z = golden("plume3d_sml")
    ocfg = op.config_for("plume3d_sml", z)
    cfg = to_product_cfg(ocfg)
    cfg.sml_update_int = 1
    pre = z["pre2"]
    attr, _ = cell_attrs(ocfg, z, "b2")
    outs = []
    for fused in (False, True):
        s = psph.Solver(cfg)
        try:
            s.set_cells(attr)
            s.upload(dump_to_particles(pre))
            if fused:
                dt, lost, ierr = s.step(0.0, 1)
                assert ierr == 0
            else:
                dt = s.timestep()
                s.set_dt(dt, 2, 1)
                s.search_neighs_consth(); s.adaptive_sml(); s.mom_engr_update(); s.smooth_density(); s.smooth_velocity()
                s.update_pos(); s.update_particles_in_bucket()
            outs.append((dt, s.download(ORDER_REFERENCE)))
        finally:
            s.close()
    assert outs[0][0] == outs[1][0]
    for f in ("coords", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure", "smlen", "bc_type"):
        assert np.array_equal(outs[0][1][f], outs[1][1][f]), f
