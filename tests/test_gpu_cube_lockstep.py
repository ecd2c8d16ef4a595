"""-m gpu: 3-D lock-step with the oracle over a run that crosses adaptive_sml.

The synthetic cube (24^3 real particles + ghost shell, 3-D SPH defaults) runs 31 time steps on the GPU with
adaptive_sml every 10th step (the reference calls it every SML_UPDATE_INT = 50th, main.cc:319-322; the interval is a
configuration value here so that a short run reaches the regime every step after the 50th of a default run is in).
Along the GPU's own trajectory selected steps are repeated by the oracle FROM THE GPU'S STATE and must agree to
RTOL_STEP in every field: a uniform-h step, the first adaptive_sml step, mixed-h steps, an adaptive_sml step that
starts from mixed h.  The pair path the GPU takes differs between those (block tiles / mixed-h lists), so this is
also the multi-step check of each.
"""
import numpy as np
import pytest

import plume_sph_b200 as psph
from plume_sph_b200 import cases
from plume_sph_b200.driver import TimeProps, run_loop
from plume_sph_b200.solver import ORDER_REFERENCE
from common import to_oracle_cfg, particles_to_oracle, relerr, RTOL_STEP

pytestmark = pytest.mark.gpu

CHECK_STEPS = (3, 10, 11, 15, 20, 21, 31)     # steps whose result is compared (10, 20: adaptive_sml runs)
FIELDS = (("coords", "x"), ("vel", "v"), ("rho", "rho"), ("engr", "e"), ("pressure", "p"), ("smlen", "h"),
          ("smoothed_v", "sv"), ("mass_frac", "ks"))


def test_cube_lock_step_across_adaptive_sml():
    cfg, rows = cases.cube(24, seed=11)
    cfg.sml_update_int = 10
    ocfg = to_oracle_cfg(cfg)
    s = psph.Solver(cfg)
    try:
        s.upload(rows)
        tp = TimeProps(max_time=1e9, max_steps=10**8, timeoutput=1e9)
        worst = {}
        for target in CHECK_STEPS:
            run_loop(s, tp, max_steps=target - 1 - tp.step)
            assert tp.step == target - 1
            before = s.download(ORDER_REFERENCE)
            o = particles_to_oracle(ocfg, before)
            o.bin()
            dt_o = o.timestep()
            run_loop(s, tp, max_steps=1)
            assert tp.step == target
            dt = tp.dtime
            assert dt == dt_o, (target, dt, dt_o)          # CFL minimum of the same state: bit-exact
            o.search()
            if target % cfg.sml_update_int == 0:
                o.adaptive_sml()
            o.mom_engr_update(dt); o.smooth_density(); o.smooth_velocity(); o.update_pos(dt, 0)
            out = s.download(ORDER_REFERENCE)
            assert np.array_equal(out["key"], before["key"])
            if target > cfg.sml_update_int:
                real = before["bc_type"] == 100
                assert len(np.unique(before["smlen"][real])) > 1000      # every particle its own h
            for f, of in FIELDS:
                e = relerr(out[f], o.a[of])
                worst[(target, f)] = e
                assert e <= RTOL_STEP, (target, f, e)
            assert np.array_equal(out["involved"], o.a["involved"])
        assert s.errors() == 0
        print("worst per step:", {t: "%.1e" % max(v for (tt, f), v in worst.items() if tt == t) for t in CHECK_STEPS})
    finally:
        s.close()
