"""-m gpu: the drop-in as a program.  oracle/_ref/3d_plume/particler_b200 is the UNMODIFIED reference driver (its main.cc,
input parsing, set-up, domain growth, ghost generation, image search, output) with the per-step stage calls forwarded to
libpsph.so (oracle/ref_build/ref_b200.cc); particler_hooked is the same driver with the reference's own stages.  Both run
the small plume for a few steps from the same inputs; the final particle tables are compared: the same particles
(keys bit-exact, so the eruption injection and the losses agree), fields within RTOL_DROPIN of the field maximum.

A third case sets PSPH_SML_INT=1 (the hook TUs' override of SML_UPDATE_INT, constant.h:189) so that both programs call
adaptive_sml every step: two steps, the second one in the per-particle-h regime and across a domain growth (the adapt
block of main.cc:256-304 runs on the host in both programs; the forwarded driver uploads the regenerated table).

Free-running end states are compared only where they can be: with the reference's own SML_UPDATE_INT the air is an exact
lattice whose spacing equals h, so hundreds of partners sit exactly ON the cutoff face |dx| = 3 h, and whether such a pair
is inside at the next step depends on whether y + dt * v (v ~ 1e-13 m/s of noise) rounds back to the lattice value or
one ulp off it — a last-bit difference in a velocity (summation order) flips pairs worth exp(-9) of a kernel sum
(measured: 240 rows with coordinates one ulp apart after step 1 -> 53 rows with 4e-4 velocity differences in step 2;
profiles/dropin_diff.py).  That is the reference algorithm's own sensitivity, not a device error, so the multi-step
statement is the LOCK-STEP one: PSPH_LOCKSTEP=1 makes the forwarded driver run the reference's stages on the host beside
the device's, reload the device with the host's table at every step, and report the per-field difference after each
whole step (growth, injection, wall ghosts and, with PSPH_SML_INT=1, adaptive_sml included) — bar RTOL_STEP."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
import refio  # noqa: E402
import ref_run as rr  # noqa: E402

pytestmark = pytest.mark.gpu
RTOL_DROPIN = 1e-9      # three steps, summation-order differences amplified by the stages that follow


def prepare(top, tags):
    shutil.copytree(os.path.join(rr.INPUTS, "Meteo_data"), os.path.join(top, "Meteo_data"))
    txt = open(os.path.join(rr.INPUTS, "run", "simulation.data")).read()
    for a, b in rr.SMALL:
        txt = txt.replace(a, b)
    smlen = [ln for ln in txt.splitlines() if ln.startswith("smoothing length")][0].split(":")[1]
    for tag in tags:
        wd = os.path.join(top, tag)
        os.makedirs(os.path.join(wd, "dump"))
        shutil.copy(os.path.join(rr.INPUTS, "run", "scale.data"), wd)
        open(os.path.join(wd, "simulation.data"), "w").write(txt)
        subprocess.run([os.path.join(rr.OUT, "3d_plume", "preprocess"), "1", smlen, "0"], cwd=wd, check=True, stdout=subprocess.DEVNULL)


# (the UNMODIFIED reference itself dies with SIGSEGV on this small domain at its second domain growth: step 10 with its
# own SML_UPDATE_INT, step 3 with PSPH_SML_INT=1 — "neigh-bucket missing", mesh_update.cc:77 — so the runs stop before)
@pytest.mark.parametrize("steps,sml_int", [(9, 0), (2, 1)])
def test_lock_step_with_the_reference_driver(steps, sml_int):
    """every step of the reference's own run (small plume: domain growth at the start of step 2, injection from step 2 on,
    wall ghosts every step; sml_int = 1: adaptive_sml every step, per-particle h from step 2 on) repeated on the device from
    the host's state: same rows, same flags, fields within RTOL_STEP of the field maximum after the whole step"""
    exe = os.path.join(rr.OUT, "3d_plume", "particler_b200")
    if not os.path.exists(exe) or not rr.have_inputs():
        pytest.skip("oracle/_ref not built")
    from common import RTOL_STEP
    top = tempfile.mkdtemp(prefix="psph_lock_")
    try:
        prepare(top, ["particler_b200"])
        wd = os.path.join(top, "particler_b200")
        env = dict(os.environ, PSPH_LOCKSTEP="1", PSPH_MAX_STEPS=str(steps))
        if sml_int:
            env["PSPH_SML_INT"] = str(sml_int)
        r = subprocess.run([exe], cwd=wd, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "device lost" not in r.stderr and "device added" not in r.stderr, r.stderr[-2000:]
        lines = [ln.split() for ln in r.stderr.splitlines() if ln.startswith("LOCKSTEP")]
        seen, worst = set(), {}
        for ln in lines:
            st = int(ln[1])
            if ln[2] == "rows":
                kv = dict(x.split("=") for x in ln[3:])
                assert kv["device"] == kv["host"] and kv["missing"] == "0" and kv["flags"] == "0", ln
                seen.add(st)
            elif ln[2] in ("dt", "mom_engr_update"):
                raise AssertionError(" ".join(ln))
            else:
                worst[ln[2]] = max(worst.get(ln[2], 0.0), float(ln[3]))
                assert float(ln[3]) <= RTOL_STEP, ln
        assert seen == set(range(1, steps + 1)), seen
        print("lock-step", steps, "steps, sml_int", sml_int, "worst:", {k: "%.1e" % v for k, v in worst.items()})
    finally:
        shutil.rmtree(top, ignore_errors=True)


@pytest.mark.parametrize("steps,sml_int", [(1, 0), (2, 1)])
def test_reference_driver_with_forwarded_stages(steps, sml_int):
    exe = {k: os.path.join(rr.OUT, "3d_plume", k) for k in ("particler_hooked", "particler_b200", "preprocess")}
    if not all(os.path.exists(p) for p in exe.values()) or not rr.have_inputs():
        pytest.skip("oracle/_ref not built")
    top = tempfile.mkdtemp(prefix="psph_dropin_")
    try:
        shutil.copytree(os.path.join(rr.INPUTS, "Meteo_data"), os.path.join(top, "Meteo_data"))
        txt = open(os.path.join(rr.INPUTS, "run", "simulation.data")).read()
        for a, b in rr.SMALL:
            txt = txt.replace(a, b)
        smlen = [ln for ln in txt.splitlines() if ln.startswith("smoothing length")][0].split(":")[1]
        final = {}
        for tag in ("particler_hooked", "particler_b200"):
            wd = os.path.join(top, tag)
            os.makedirs(os.path.join(wd, "dump"))
            shutil.copy(os.path.join(rr.INPUTS, "run", "scale.data"), wd)
            open(os.path.join(wd, "simulation.data"), "w").write(txt)
            subprocess.run([exe["preprocess"], "1", smlen, "0"], cwd=wd, check=True, stdout=subprocess.DEVNULL)
            env = dict(os.environ, PSPH_DUMP_DIR=os.path.join(wd, "dump"), PSPH_DUMP_STEPS="9999-9999", PSPH_MAX_STEPS=str(steps))
            if sml_int:
                env["PSPH_SML_INT"] = str(sml_int)
            r = subprocess.run([exe[tag]], cwd=wd, env=env, capture_output=True, text=True)
            assert r.returncode == 0, (tag, r.stderr[-2000:])
            assert "device lost" not in r.stderr and "device added" not in r.stderr, r.stderr[-2000:]
            _, final[tag] = refio.read_dump(os.path.join(wd, "dump", "s%04d_final_post.bin" % steps))
        a, b = final["particler_hooked"], final["particler_b200"]
        assert len(a) == len(b) and np.array_equal(a["key"], b["key"])
        assert np.array_equal(a["bc_type"], b["bc_type"]) and np.array_equal(a["involved"], b["involved"])
        for f in ("coord", "vel", "rho", "engr", "mass_frac", "pressure", "smlen", "smoothed_v"):
            scale = max(np.abs(a[f]).max(), 1e-300)
            d = np.abs(a[f] - b[f])
            d = d if d.ndim == 1 else d.max(axis=1)
            err = float(d.max() / scale)
            badrows = np.nonzero(d > RTOL_DROPIN * scale)[0]
            assert err <= RTOL_DROPIN, (f, err, "rows off by bc_type:", dict(zip(*np.unique(a["bc_type"][badrows], return_counts=True))),
                                        "first:", a[f][badrows[:3]], b[f][badrows[:3]])
    finally:
        shutil.rmtree(top, ignore_errors=True)
