"""-m gpu: 1-D shock tubes (BASELINE configs 0/1) — CUDA path against dumps of the unmodified
reference (tests/golden/*_sph.npz, *_gsph.npz): stage by stage over the first three steps, and
the state after hundreds of steps of the full loop driven by the host mirror of main.cc.
"""
import numpy as np
import pytest

import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, relerr, RTOL_STAGE, RTOL_MID, RTOL_LONG, D2P
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_REFERENCE
from plume_sph_b200.driver import TimeProps, run_loop

pytestmark = pytest.mark.gpu

# name -> GSPH build?  (inputs: the reference's bin/simulation_1D_*.data, see make_golden.py)
CASES = {"sod_sph": False, "shu_sph": False, "sod_gsph": True, "sod0_gsph": True, "shu_gsph": True, "blast_gsph": True}


@pytest.mark.parametrize("name", list(CASES))   # a missing golden fails (golden() raises), it is not filtered out
def test_first_steps_stage_by_stage(name):
    gsph = CASES[name]
    z = golden(name)
    cfg = to_product_cfg(op.config_1d(gsph=gsph))
    s = psph.Solver(cfg)
    s.upload(dump_to_particles(z["init"]))
    for step in (1, 2, 3):
        dt = float(z["dt"][step - 1])
        cfl = s.timestep()
        # the CFL dt is a min of quotients of the current state: bit-exact at step 1, within the
        # state's own rounding differences afterwards
        assert (cfl == dt) if step == 1 else abs(cfl - dt) <= 1e-12 * dt
        s.set_dt(dt, step, 1)
        s.search_neighs_consth()
        seq = [("sml", s.adaptive_sml, ["smlen"])]
        if gsph:
            seq.append(("grad", s.calc_gradients, ["grad"]))
        seq += [("mom", s.mom_engr_update, ["vel", "engr", "pressure", "sound_speed"]),
                ("dens", s.smooth_density, ["rho", "mass_frac", "engr", "pressure", "sound_speed"]),
                ("pos", s.update_pos, ["coord"])]
        for stage, fn, fields in seq:
            fn()
            out = s.download(ORDER_REFERENCE)
            for f in fields:
                ref = z["s%d_%s_%s" % (step, stage, f)]
                e = relerr(out[D2P[f]], ref)
                # errors of earlier steps carry over: allow them to add up
                assert e <= RTOL_STAGE * 10 * step, (name, step, stage, f, e)
    assert s.errors() == 0
    s.close()


@pytest.mark.parametrize("name", list(CASES))   # a missing golden fails (golden() raises), it is not filtered out
def test_full_loop_against_reference(name):
    """The whole loop (host mirror of main.cc driving the GPU stages) against the reference's state
    after 30 steps (tight: rounding differences only add up) and at the end of the run (per-field
    tolerances from the reference algorithm's own measured sensitivity, common.RTOL_LONG)."""
    gsph = CASES[name]
    z = golden(name)
    cfg = to_product_cfg(op.config_1d(gsph=gsph))
    s = psph.Solver(cfg)
    s.upload(dump_to_particles(z["init"]))
    tp = TimeProps(max_time=float(z["t_end"][0]), max_steps=10**8, timeoutput=float(z["t_out"][0]))
    done = 0
    for tag in ("mid", "final"):
        nsteps = int(z[tag + "_steps"][0])
        run_loop(s, tp, max_steps=nsteps - done)
        done = nsteps
        assert tp.step == nsteps
        assert abs(tp.time - float(z[tag + "_time"][0])) <= (RTOL_MID if tag == "mid" else 1e-6) * tp.time
        out = s.download(ORDER_REFERENCE)
        ref = z[tag]
        assert np.array_equal(out["key"], ref["key"])
        for f in ("coord", "vel", "rho", "engr", "pressure", "smlen"):
            a = out[D2P[f]][:, 0] if out[D2P[f]].ndim == 2 else out[D2P[f]]
            b = ref[f][:, 0] if ref[f].ndim == 2 else ref[f]
            e = relerr(a, b)
            tol = RTOL_MID if tag == "mid" else RTOL_LONG[f]
            assert e <= tol, (name, tag, f, e)
    assert s.errors() == 0
    s.close()


@pytest.mark.parametrize("name", list(CASES))   # a missing golden fails (golden() raises), it is not filtered out
def test_lock_step_with_oracle(name):
    """Long-run parity stated tightly: along 300 steps of the GPU's own trajectory, every 20th
    step is repeated by the oracle FROM THE GPU'S STATE and must agree to RTOL_STEP in every field."""
    from common import particles_to_oracle, to_oracle_cfg, RTOL_STEP
    z = golden(name)
    gsph = CASES[name]
    ocfg = op.config_1d(gsph=gsph)
    cfg = to_product_cfg(ocfg)
    s = psph.Solver(cfg)
    s.upload(dump_to_particles(z["init"]))
    tp = TimeProps(max_time=float(z["t_end"][0]), max_steps=10**8, timeoutput=float(z["t_out"][0]))
    nsteps = min(300, int(z["final_steps"][0])) if len(z["init"]) < 1000 else 60
    for k in range(0, nsteps, 20):
        run_loop(s, tp, max_steps=19)
        before = s.download(ORDER_REFERENCE)
        o = particles_to_oracle(ocfg, before)
        o.bin(); o.search()
        dt_o = o.timestep()
        step0 = tp.step
        run_loop(s, tp, max_steps=1)
        assert tp.step == step0 + 1
        dt = tp.dtime
        assert dt <= dt_o * (1 + 1e-15)
        o.adaptive_sml()
        if gsph:
            o.calc_gradients()
        o.mom_engr_update(dt); o.smooth_density(); o.update_pos(dt, 1)
        out = s.download(ORDER_REFERENCE)
        for f, of in (("coords", "x"), ("vel", "v"), ("rho", "rho"), ("engr", "e"), ("pressure", "p"), ("smlen", "h")) + \
                ((("grad", "grad"),) if gsph else ()):
            e = relerr(out[f], o.a[of])
            assert e <= RTOL_STEP, (name, tp.step, f, e)
    s.close()
