#!/usr/bin/env python3
"""Debugging aid for the drop-in program (tests/test_gpu_dropin.py): runs particler_hooked (reference stages) and
particler_b200 (stages forwarded to libpsph.so) side by side with per-stage dumps of one time step and reports, stage by
stage, the largest relative difference of every field and which rows carry it.   usage: dropin_diff.py STEP [SML_INT]"""
import os, shutil, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "ref_build")]
import refio, ref_run as rr

step = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sml = sys.argv[2] if len(sys.argv) > 2 else ""
top = tempfile.mkdtemp(prefix="psph_dd_")
shutil.copytree(os.path.join(rr.INPUTS, "Meteo_data"), os.path.join(top, "Meteo_data"))
txt = open(os.path.join(rr.INPUTS, "run", "simulation.data")).read()
for a, b in rr.SMALL:
    txt = txt.replace(a, b)
smlen = [ln for ln in txt.splitlines() if ln.startswith("smoothing length")][0].split(":")[1]
dumps = {}
for tag in ("particler_hooked", "particler_b200"):
    wd = os.path.join(top, tag)
    os.makedirs(os.path.join(wd, "dump"))
    shutil.copy(os.path.join(rr.INPUTS, "run", "scale.data"), wd)
    open(os.path.join(wd, "simulation.data"), "w").write(txt)
    subprocess.run([os.path.join(rr.OUT, "3d_plume", "preprocess"), "1", smlen, "0"], cwd=wd, check=True, stdout=subprocess.DEVNULL)
    env = dict(os.environ, PSPH_DUMP_DIR=os.path.join(wd, "dump"), PSPH_DUMP_STEPS="%d-%d" % (step, step), PSPH_MAX_STEPS=str(step))
    if sml:
        env["PSPH_SML_INT"] = sml
    r = subprocess.run([os.path.join(rr.OUT, "3d_plume", tag)], cwd=wd, env=env, capture_output=True, text=True)
    print(tag, "rc", r.returncode, r.stderr[-300:].replace("\n", " | "))
    dumps[tag] = os.path.join(wd, "dump")
for stage in ("mom_pre", "sml_post", "mom_post", "dens_post", "svel_post", "pos_post", "bcond_post", "final_post"):
    fa, fb = (os.path.join(dumps[t], "s%04d_%s.bin" % (step, stage)) for t in ("particler_hooked", "particler_b200"))
    if not (os.path.exists(fa) and os.path.exists(fb)):
        print(stage, "missing", os.path.exists(fa), os.path.exists(fb))
        continue
    _, a = refio.read_dump(fa)
    _, b = refio.read_dump(fb)
    if len(a) != len(b) or not np.array_equal(a["key"], b["key"]):
        print(stage, "row sets differ", len(a), len(b))
        continue
    out = []
    for f in ("coord", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure", "smlen", "mass", "gamma", "sound_speed",
              "temperature", "specific_heat_p", "phase_density", "smoothed_e"):
        d = np.abs(a[f] - b[f])
        d = d if d.ndim == 1 else d.max(axis=1)
        sc = max(np.abs(a[f]).max(), 1e-300)
        bad = np.nonzero(d > 1e-11 * sc)[0]
        out.append("%s %.1e (%d rows, bc %s)" % (f, d.max() / sc, len(bad), dict(zip(*np.unique(a["bc_type"][bad], return_counts=True)))))
        if len(bad) and os.environ.get("DD_DETAIL") == stage:
            for r in bad[:12]:
                print("   ", f, "row", r, "bc", a["bc_type"][r], "phase", a["phase"][r], "inv", a["involved"][r], "x", a["coord"][r], "h", a["smlen"][r],
                      "ref", a[f][r], "b200", b[f][r])
    for f in ("bc_type", "involved", "phase", "guest"):
        out.append("%s neq %d" % (f, int((a[f] != b[f]).sum())))
    print(stage, "; ".join(out))
keep = os.environ.get("DD_KEEP")
if keep:   # stage dumps of both programs, for a closer look elsewhere
    for t in dumps:
        for stage in keep.split(","):
            src = os.path.join(dumps[t], "s%04d_%s.bin" % (step, stage))
            if os.path.exists(src):
                os.makedirs(os.path.join(ROOT, "gpurun_out", "dd"), exist_ok=True)
                shutil.copy(src, os.path.join(ROOT, "gpurun_out", "dd", "%s_s%04d_%s.bin" % (t, step, stage)))
shutil.rmtree(top, ignore_errors=True)
