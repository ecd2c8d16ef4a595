#!/usr/bin/env python3
"""Run the compiled, UNMODIFIED reference (oracle/_ref/<variant>/particler_hooked) on its own
3-D plume case and time its loop.  Test / baseline infrastructure only (bench.py
--impl reference, bench.py cpu_baseline): nothing in the product imports this.

`prepare_inputs()` runs in the build container (needs /root/reference): it places the
reference's run-time INPUT files (bin/simulation.data, bin/scale.data and the Meteo sounding
the data file names) under oracle/_ref/inputs/ so the timing can be repeated on the GPU box,
where /root/reference does not exist.  oracle/_ref/ is git-ignored build output.

`time_plume3d()` copies those inputs to a scratch directory, shrinks the INITIAL domain (the
reference's own `initial domain` lines of simulation.data; the mesh, atmosphere, eruption
boundary, parameters and code are untouched), runs the reference's `preprocess 1 <h> 0` and
then P independent single-rank copies of the reference loop side by side (there is no MPI
in this image, SURVEY.md Appendix D, so "all host cores" = P replicas of the 1-rank build).
Per-step wall-clock stamps come from the hook TU's PSPH_STAGE_TIMES output.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
INPUTS = os.path.join(OUT, "inputs")
REF = os.environ.get("PSPH_REFERENCE", "/root/reference")

# initial-domain edits of bin/simulation.data (default: x,y in [-2000,2000], z top 2700)
SMALL = [("initial domain:-2000", "initial domain:-350"), ("initial domain:2000", "initial domain:350"),
         ("initial domain:2700", "initial domain:2400")]
MEDIUM = [("initial domain:-2000", "initial domain:-1000"), ("initial domain:2000", "initial domain:1000")]


def prepare_inputs():
    if not os.path.isdir(os.path.join(REF, "bin")):
        return False
    os.makedirs(os.path.join(INPUTS, "run"), exist_ok=True)
    for f in ("simulation.data", "scale.data"):
        shutil.copy(os.path.join(REF, "bin", f), os.path.join(INPUTS, "run", f))
    txt = open(os.path.join(REF, "bin", "simulation.data")).read()
    m = re.search(r"(\.\./Meteo_data/\S+)", txt)
    os.makedirs(os.path.join(INPUTS, "Meteo_data"), exist_ok=True)
    if m:
        name = os.path.basename(m.group(1))
        shutil.copy(os.path.join(REF, "Meteo_data", name), os.path.join(INPUTS, "Meteo_data", name))
    return True


def have_inputs():
    return os.path.exists(os.path.join(INPUTS, "run", "simulation.data")) and \
        os.path.exists(os.path.join(OUT, "3d_plume", "particler_hooked"))


def _parse(log):
    stamps, real, total, stages, reals, totals = {}, None, None, {}, {}, {}
    for ln in log.splitlines():
        if ln.startswith("PSPH_STEP "):
            f = ln.split()
            stamps[int(f[1])] = float(f[2])
            reals[int(f[1])] = int(f[3].split("=")[1])
            totals[int(f[1])] = int(f[4].split("=")[1])
        elif ln.startswith("PSPH_PARTICLES"):
            real = int(re.search(r"real=(\d+)", ln).group(1))
            total = int(re.search(r"total=(\d+)", ln).group(1))
        elif ln.startswith("PSPH_STAGE "):
            _, name, t = ln.split()
            stages[name] = float(t)
    return stamps, real, total, stages, reals, totals


def time_plume3d(steps, warmup=0, replicas=1, edits=SMALL, variant="3d_plume", keep=None, dump=None):
    """-> dict(seconds=[per replica time of the `steps` timed steps], real=, total=, stages=).
    dump: directory that receives the hook dumps of steps 0-1 of replica 0 (the state after the reference's own
    initialisation, its bucket attributes, image table and vent record: the set-up a drop-in run starts from)"""
    top = keep or tempfile.mkdtemp(prefix="psph_ref_")
    try:
        shutil.copytree(os.path.join(INPUTS, "Meteo_data"), os.path.join(top, "Meteo_data"), dirs_exist_ok=True)
        txt = open(os.path.join(INPUTS, "run", "simulation.data")).read()
        for a, b in edits:
            assert a in txt, a
            txt = txt.replace(a, b)
        smlen = [l for l in txt.splitlines() if l.startswith("smoothing length")][0].split(":")[1]
        procs = []
        for r in range(replicas):
            wd = os.path.join(top, "run%d" % r)
            os.makedirs(wd, exist_ok=True)
            shutil.copy(os.path.join(INPUTS, "run", "scale.data"), wd)
            open(os.path.join(wd, "simulation.data"), "w").write(txt)
            if r == 0:
                subprocess.run([os.path.join(OUT, variant, "preprocess"), "1", smlen, "0"], cwd=wd, check=True,
                               stdout=subprocess.DEVNULL)
            else:
                for f in os.listdir(os.path.join(top, "run0")):
                    if f.startswith("funky"):
                        os.link(os.path.join(top, "run0", f), os.path.join(wd, f))
        env = dict(os.environ, PSPH_STAGE_TIMES="1", PSPH_MAX_STEPS=str(steps + warmup))
        for r in range(replicas):
            wd = os.path.join(top, "run%d" % r)
            if dump and r == 0:
                os.makedirs(dump, exist_ok=True)
                env = dict(env, PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="0-1")
            procs.append(subprocess.Popen([os.path.join(OUT, variant, "particler_hooked")], cwd=wd, env=env,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
        secs, stages, updates, total = [], {}, 0, 0
        for p in procs:
            _, err = p.communicate()
            stamps, _, _, stages, reals, totals = _parse(err)
            if (steps + warmup) not in stamps or warmup not in stamps:
                raise RuntimeError("reference run did not finish: " + err[-400:])
            secs.append(stamps[steps + warmup] - stamps[warmup])
            # particles advanced by the timed steps (the population changes when the domain grows)
            updates = sum(reals[k] for k in range(warmup, warmup + steps))
            total = max(totals[k] for k in range(warmup, warmup + steps))
        return dict(seconds=secs, updates=updates, real=updates // steps, total=total, stages=stages)
    finally:
        if keep is None:
            shutil.rmtree(top, ignore_errors=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "prepare":
        print("inputs prepared:", prepare_inputs())
    else:
        print(time_plume3d(int(sys.argv[1]) if len(sys.argv) > 1 else 1, replicas=int(sys.argv[2]) if len(sys.argv) > 2 else 1))
