"""-m gpu: the drop-in as a program.  oracle/_ref/3d_plume/particler_b200 is the UNMODIFIED reference driver (its main.cc,
input parsing, set-up, domain growth, ghost generation, image search, output) with the per-step stage calls forwarded to
libpsph.so (oracle/ref_build/ref_b200.cc); particler_hooked is the same driver with the reference's own stages.  Both run
the small plume for a few steps from the same inputs; the final particle tables are compared: the same particles
(keys bit-exact, so the eruption injection and the losses agree), fields within RTOL_DROPIN of the field maximum.

A third case sets PSPH_SML_INT=1 (the hook TUs' override of SML_UPDATE_INT, constant.h:189) so that both programs call
adaptive_sml every step: two steps, the second one in the per-particle-h regime and across a domain growth (the adapt
block of main.cc:256-304 runs on the host in both programs; the forwarded driver uploads the regenerated table)."""
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


@pytest.mark.parametrize("steps,sml_int", [(1, 0), (3, 0), (2, 1)])
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
