#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

The reference (Plume-SPH) has no tests or golden vectors of its own, so the
pins are outputs of the reference itself: oracle/_ref/<variant>/particler_hooked is the
reference's own main.cc + stage functions compiled from /root/reference/src in place
(recipe: oracle/ref_build/build_ref.py); the hooks only dump the particle table around each
stage call.  This script runs those binaries on the reference's own input files
(/root/reference/bin/simulation_*.data) and stores what the tests need as compressed .npz.

Run in the build container only (needs /root/reference and oracle/_ref):
    python oracle/ref_build/build_ref.py && python tests/golden/make_golden.py
The fixtures are committed; nothing at test time reads /root/reference.
"""
import os
import shutil
import subprocess
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refio  # noqa: E402

REF = os.environ.get("PSPH_REFERENCE", "/root/reference")
REFBIN = os.path.join(ROOT, "oracle", "_ref")

NMID = 30

STAGE_FIELDS = {
    "sml": ["smlen"],
    "grad": ["grad"],
    "mom": ["vel", "engr", "pressure", "sound_speed", "temperature", "specific_heat_p", "gamma", "phase_density"],
    "dens": ["rho", "mass_frac", "engr", "pressure", "sound_speed", "temperature", "involved", "gamma"],
    "svel": ["smoothed_v"],
    "pos": ["coord", "bc_type", "smlen", "involved"],
}


def run_ref(variant, workdir, env):
    exe = os.path.join(REFBIN, variant, "particler_hooked")
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    with open(os.path.join(workdir, "out.log"), "w") as f:
        subprocess.run([exe], cwd=workdir, env=e, stdout=f, stderr=subprocess.STDOUT, check=False)


def stage_dumps(dump, step, stages):
    out = {}
    for st in stages:
        fn = os.path.join(dump, "s%04d_%s_post.bin" % (step, st))
        if not os.path.exists(fn):
            continue
        _, rows = refio.read_dump(fn)
        for f in STAGE_FIELDS[st]:
            out["s%d_%s_%s" % (step, st, f)] = rows[f]
    return out


def golden_1d(name, variant, simdata, nfinal):
    with tempfile.TemporaryDirectory() as wd:
        shutil.copy(os.path.join(REF, "bin", "scale.data"), wd)
        shutil.copy(os.path.join(REF, "bin", simdata), os.path.join(wd, "simulation.data"))
        dump = os.path.join(wd, "dump")
        os.makedirs(dump)
        run_ref(variant, wd, dict(PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="1-3", PSPH_MAX_STEPS=nfinal))
        h0, init = refio.read_dump(os.path.join(dump, "s0000_init_post.bin"))
        out = dict(init=init)
        dts = []
        for step in (1, 2, 3):
            h, _ = refio.read_dump(os.path.join(dump, "s%04d_timestep_post.bin" % step))
            dts.append(h["dt"])
            out.update(stage_dumps(dump, step, ["sml", "grad", "mom", "dens", "pos"]))
        hf, fin = refio.read_dump(os.path.join(dump, "s%04d_final_post.bin" % nfinal))
        # state after NMID steps (before rounding differences can flip one of the algorithm's
        # discrete decisions, see tests/test_oracle_vs_ref.py::test_reference_algorithm_is_sensitive)
        dump2 = os.path.join(wd, "dump2")
        os.makedirs(dump2)
        run_ref(variant, wd, dict(PSPH_DUMP_DIR=dump2, PSPH_DUMP_STEPS="0-0", PSPH_MAX_STEPS=NMID))
        hm, mid = refio.read_dump(os.path.join(dump2, "s%04d_final_post.bin" % NMID))
        out["mid"] = mid
        out["mid_time"] = np.asarray([hm["time"]])
        out["mid_steps"] = np.asarray([NMID])
        sim = [ln.split(":", 1)[1].strip() for ln in open(os.path.join(wd, "simulation.data")) if ":" in ln]
        out["t_end"] = np.asarray([float(np.float32(float(sim[0])))])   # dataread.cc parses with stof
        out["t_out"] = np.asarray([float(np.float32(float(sim[2])))])
        out["dt"] = np.asarray(dts)
        out["final"] = fin
        out["final_time"] = np.asarray([hf["time"]])
        out["final_steps"] = np.asarray([nfinal])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n =", len(init), "final time", hf["time"], "->", os.path.getsize(os.path.join(HERE, name + ".npz")), "B")


def golden_3d(name, variant, edits, list_stride=4, simdata="simulation.data"):
    """Small 3-D plume: the reference's default case (bin/simulation.data, Eyjafjallajokull
    parameters, Meteo sounding) with a smaller initial domain; state around every stage of step 1."""
    with tempfile.TemporaryDirectory() as top:
        wd = os.path.join(top, "run")
        os.makedirs(wd)
        shutil.copytree(os.path.join(REF, "Meteo_data"), os.path.join(top, "Meteo_data"))
        shutil.copy(os.path.join(REF, "bin", "scale.data"), wd)
        txt = open(os.path.join(REF, "bin", simdata)).read()
        for a, b in edits:
            assert a in txt, a
            txt = txt.replace(a, b)
        open(os.path.join(wd, "simulation.data"), "w").write(txt)
        smlen = [l for l in txt.splitlines() if l.startswith("smoothing length")][0].split(":")[1]
        subprocess.run([os.path.join(REFBIN, variant, "preprocess"), "1", smlen, "0"], cwd=wd, check=True,
                       stdout=subprocess.DEVNULL)
        dump = os.path.join(wd, "dump")
        os.makedirs(dump)
        run_ref(variant, wd, dict(PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="1-1", PSPH_DUMP_NEIGH=1, PSPH_MAX_STEPS=1))
        h, pre = refio.read_dump(os.path.join(dump, "s0001_timestep_post.bin"))
        out = dict(pre=pre, dt=np.asarray([h["dt"]]))
        out.update(stage_dumps(dump, 1, ["grad", "mom", "dens", "svel", "pos"]))
        # buckets: attributes + plists (rows of `pre`) before the step and after the re-bin
        for tag, fn in (("b0", "s0001_timestep_buckets.bin"), ("b1", "s0001_rebin_buckets.bin")):
            bk = refio.read_buckets(os.path.join(dump, fn))
            out[tag + "_mincrd"] = np.asarray([b["mincrd"] for b in bk])
            out[tag + "_key"] = np.asarray([b["key"] for b in bk], np.uint32)
            out[tag + "_attr"] = np.asarray([[b["type"], b["has_involved"], b["active"], b["guest"]] for b in bk], np.int32)
            out[tag + "_bindex"] = np.asarray([b["bucket_index"] for b in bk], np.int32)
            out[tag + "_bnd"] = np.asarray([b["bnd"] for b in bk], np.float64)
            out[tag + "_off"] = np.cumsum([0] + [len(b["rows"]) for b in bk]).astype(np.int64)
            out[tag + "_rows"] = np.concatenate([b["rows"] for b in bk]).astype(np.int32)
        # apply_bcond at the end of step 1 (main.cc:468): state before, wall-ghost images, bucket plists, and the wall
        # ghosts' state after
        _, bpre = refio.read_dump(os.path.join(dump, "s0001_bcond_pre.bin"))
        _, bpost = refio.read_dump(os.path.join(dump, "s0001_bcond_post.bin"))
        img = refio.read_images(os.path.join(dump, "s0001_images.bin"))
        bk = refio.read_buckets(os.path.join(dump, "s0001_bcond_buckets.bin"))
        out["bc_pre"] = bpre
        rowof = {tuple(k): i for i, k in enumerate(bpre["key"])}
        bmin = {tuple(b["key"]): b["mincrd"] for b in bk}
        out["bc_img_row"] = np.asarray([rowof[tuple(k)] for k in img["ghost_key"]], np.int32)
        out["bc_img_coord"] = np.asarray(img["coord"], np.float64)
        out["bc_img_bucket_mincrd"] = np.asarray([bmin[tuple(k)] for k in img["bucket_key"]], np.float64)
        out["b2_mincrd"] = np.asarray([b["mincrd"] for b in bk])
        out["b2_attr"] = np.asarray([[b["type"], b["has_involved"], b["active"], b["guest"]] for b in bk], np.int32)
        out["b2_bindex"] = np.asarray([b["bucket_index"] for b in bk], np.int32)
        out["b2_bnd"] = np.asarray([b["bnd"] for b in bk], np.float64)
        out["b2_off"] = np.cumsum([0] + [len(b["rows"]) for b in bk]).astype(np.int64)
        out["b2_rows"] = np.concatenate([b["rows"] for b in bk]).astype(np.int32)
        wall = np.nonzero(bpre["bc_type"] == 2)[0]
        out["bc_wall_rows"] = wall.astype(np.int32)
        for f in ("rho", "vel", "engr", "mass_frac", "pressure", "sound_speed", "temperature", "phase_density", "gamma",
                  "specific_heat_p"):
            out["bc_post_" + f] = bpost[f][wall]
        other = np.ones(len(bpre), bool)
        other[wall] = False
        assert all(np.array_equal(bpre[f][other], bpost[f][other]) for f in bpre.dtype.names), "apply_bcond touched a non-wall row"
        # neighbour lists of every list_stride-th listed particle
        off, idx = refio.read_neigh(os.path.join(dump, "s0001_neigh.bin"))
        listed = np.nonzero(np.diff(off) > 0)[0]
        out["nlisted"] = np.asarray([len(listed)])
        out["list_counts"] = np.diff(off).astype(np.int32)
        sel = listed[::list_stride]
        out["list_rows"] = sel.astype(np.int32)
        out["list_off"] = np.cumsum([0] + [off[i + 1] - off[i] for i in sel]).astype(np.int64)
        out["list_idx"] = np.concatenate([np.sort(idx[off[i]:off[i + 1]]) for i in sel]).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "n =", len(pre), "listed", len(listed), "->", os.path.getsize(os.path.join(HERE, name + ".npz")), "B")


def pack_buckets(out, tag, fn):
    bk = refio.read_buckets(fn)
    out[tag + "_mincrd"] = np.asarray([b["mincrd"] for b in bk])
    out[tag + "_key"] = np.asarray([b["key"] for b in bk], np.uint32)
    out[tag + "_attr"] = np.asarray([[b["type"], b["has_involved"], b["active"], b["guest"]] for b in bk], np.int32)
    out[tag + "_bindex"] = np.asarray([b["bucket_index"] for b in bk], np.int32)
    out[tag + "_bnd"] = np.asarray([b["bnd"] for b in bk], np.float64)
    out[tag + "_off"] = np.cumsum([0] + [len(b["rows"]) for b in bk]).astype(np.int64)
    out[tag + "_rows"] = np.concatenate([b["rows"] for b in bk]).astype(np.int32)
    return bk


def golden_3d_sml(name="plume3d_sml", variant="3d_plume", edits=None, simdata="simulation.data", sml_int=1):
    """The regime AFTER adaptive_sml (neighs.cc:479-693; main.cc:319-322 calls it every SML_UPDATE_INT = 50 steps): the
    small plume with PSPH_SML_INT=1 (the hook TU's override of that constant, nothing else changes), time steps 1 and 2.
    Step 1 runs adaptive_sml from the uniform initial smoothing lengths; step 2 starts with every real particle carrying
    its own h (and, the domain having grown in its adapt block, with freshly converted real particles that still carry
    the initial one), runs adaptive_sml again from there and every summation stage with mixed smoothing lengths.
    Per step: the state search_neighs_consth sees (after the adapt block) with the bucket plists, dt, and what every
    stage changed — stored sparse: <key>__rows = the rows whose value differs from the state before the stage,
    <key> = their new values (tests/common.py::sparse_field puts them back)."""
    edits = SMALL_3D if edits is None else edits
    with tempfile.TemporaryDirectory() as top:
        wd = os.path.join(top, "run")
        os.makedirs(wd)
        shutil.copytree(os.path.join(REF, "Meteo_data"), os.path.join(top, "Meteo_data"))
        shutil.copy(os.path.join(REF, "bin", "scale.data"), wd)
        txt = open(os.path.join(REF, "bin", simdata)).read()
        for a, b in edits:
            assert a in txt, a
            txt = txt.replace(a, b)
        open(os.path.join(wd, "simulation.data"), "w").write(txt)
        smlen = [l for l in txt.splitlines() if l.startswith("smoothing length")][0].split(":")[1]
        subprocess.run([os.path.join(REFBIN, variant, "preprocess"), "1", smlen, "0"], cwd=wd, check=True,
                       stdout=subprocess.DEVNULL)
        dump = os.path.join(wd, "dump")
        os.makedirs(dump)
        env = dict(PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="1-2", PSPH_MAX_STEPS=2)
        if sml_int:
            env["PSPH_SML_INT"] = sml_int       # (sml_int = 0: the reference's own SML_UPDATE_INT, no adaptive_sml in steps 1-2)
        run_ref(variant, wd, env)
        out = {}
        dts = []
        for st in (1, 2):
            h, pre = refio.read_dump(os.path.join(dump, "s%04d_search_post.bin" % st))
            dts.append(h["dt"])
            out["pre%d" % st] = pre
            pack_buckets(out, "b%d" % st, os.path.join(dump, "s%04d_search_buckets.bin" % st))
            cur = {f: pre[f].copy() for f in pre.dtype.names}
            for stage in ("sml", "grad", "mom", "dens", "svel", "pos"):
                fn = os.path.join(dump, "s%04d_%s_post.bin" % (st, stage))
                if (stage == "sml" and not sml_int) or (stage == "grad" and "gsph" not in variant):
                    continue
                assert os.path.exists(fn), fn
                _, rows = refio.read_dump(fn)
                assert np.array_equal(rows["key"], pre["key"])
                for f in rows.dtype.names:
                    a, b = cur[f], rows[f]
                    ch = (a != b) if a.ndim == 1 else (a != b).any(axis=1)
                    if f in STAGE_FIELDS[stage]:
                        idx = np.nonzero(ch)[0].astype(np.int32)
                        out["s%d_%s_%s__rows" % (st, stage, f)] = idx
                        out["s%d_%s_%s" % (st, stage, f)] = b[idx]
                    cur[f] = b.copy()
        out["dt"] = np.asarray(dts)
        real = out["pre2"]["bc_type"] == 100
        print(name, "n =", len(out["pre1"]), len(out["pre2"]), "real", int(real.sum()), "distinct h of real rows at the start of step 2:",
              len(np.unique(out["pre2"]["smlen"][real])))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "->", os.path.getsize(os.path.join(HERE, name + ".npz")), "B")


def golden_grow(name="grow3d", variant="3d_plume", edits=None, simdata="simulation.data"):
    """Domain growth in the small plume: scan_outside_layer at the end of step 1 raises `adapt`
    (scan_outside_layer.cc:44-121), the adapt block at the start of step 2 (main.cc:256-304) turns the pressure ghosts of the
    outer layer real (adapt_domain.cc:41-194) and fills the newly exposed buckets with pressure ghosts whose state comes from
    the Meteo sounding (add_pressure_ghost.cc:36-316, sph_lib.cc:724-766,831-872).  Stored: the table the three calls start
    from, the bucket attributes (with particles_type, erupt_flag) around each call, what adapt_domain changed, the rows
    add_pressure_ghost added with their positions in the table's iteration order, and the call's parameters."""
    edits = SMALL_3D if edits is None else edits
    with tempfile.TemporaryDirectory() as top:
        wd = os.path.join(top, "run")
        os.makedirs(wd)
        shutil.copytree(os.path.join(REF, "Meteo_data"), os.path.join(top, "Meteo_data"))
        shutil.copy(os.path.join(REF, "bin", "scale.data"), wd)
        txt = open(os.path.join(REF, "bin", simdata)).read()
        for a, b in edits:
            assert a in txt, a
            txt = txt.replace(a, b)
        open(os.path.join(wd, "simulation.data"), "w").write(txt)
        smlen = [l for l in txt.splitlines() if l.startswith("smoothing length")][0].split(":")[1]
        subprocess.run([os.path.join(REFBIN, variant, "preprocess"), "1", smlen, "0"], cwd=wd, check=True,
                       stdout=subprocess.DEVNULL)
        dump = os.path.join(wd, "dump")
        os.makedirs(dump)
        run_ref(variant, wd, dict(PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="1-2", PSPH_MAX_STEPS=2))
        out = {}
        out["adapt"] = np.asarray([int(open(os.path.join(dump, "s%04d_scan_adapt.txt" % k)).read()) for k in (1, 2)])
        _, t0 = refio.read_dump(os.path.join(dump, "s0001_scan_pre.bin"))
        _, t1 = refio.read_dump(os.path.join(dump, "s0002_adapt_pre.bin"))
        _, t2 = refio.read_dump(os.path.join(dump, "s0002_adapt_post.bin"))
        _, t3 = refio.read_dump(os.path.join(dump, "s0002_pghost_post.bin"))
        for f in ("key", "bc_type", "involved", "guest", "coord"):   # (apply_bcond ran in between: wall-ghost states only)
            assert np.array_equal(t0[f], t1[f]), f
        out["table"] = t1
        ch = (t1["bc_type"] != t2["bc_type"]) | (t1["involved"] != t2["involved"])
        assert all(np.array_equal(t1[f], t2[f]) for f in t1.dtype.names if f not in ("bc_type", "involved"))
        out["adapt_rows"] = np.nonzero(ch)[0].astype(np.int32)
        out["adapt_bc_type"] = t2["bc_type"][ch]
        out["adapt_involved"] = t2["involved"][ch]
        old = {tuple(k) for k in t2["key"]}
        isnew = np.asarray([tuple(k) not in old for k in t3["key"]])
        assert np.array_equal(t3[~isnew], t2)                        # the old rows keep their relative order, untouched
        out["pghost_pos"] = np.nonzero(isnew)[0].astype(np.int32)
        out["pghost_rows"] = t3[isnew]
        # add_wall_ghost (add_wall_ghost.cc:44-293), after shift_brief_buck: the rows it added and where they sit
        _, t4 = refio.read_dump(os.path.join(dump, "s0002_wghost_post.bin"))
        old3 = {tuple(k) for k in t3["key"]}
        isnew4 = np.asarray([tuple(k) not in old3 for k in t4["key"]])
        assert np.array_equal(t4[~isnew4], t3)
        out["wghost_pos"] = np.nonzero(isnew4)[0].astype(np.int32)
        out["wghost_rows"] = t4[isnew4]
        # (wghost_pre = the buckets after shift_brief_buck, shift_brief_buck.cc:33-129; bgmesh_post = after update_bgmesh,
        # mesh_update.cc:32-245, which follows add_wall_ghost at main.cc:286)
        for tag in ("wghost_pre", "wghost_post", "bgmesh_post"):
            bk = refio.read_bucketsx(os.path.join(dump, "s0002_" + tag + "_bucketsx.bin"))
            out[tag + "_mincrd"] = np.asarray([b["mincrd"] for b in bk])
            out[tag + "_attr"] = np.asarray([[b["type"], b["has_involved"], b["active"], b["guest"], b["particles_type"],
                                             b["erupt_flag"]] for b in bk], np.int32)
            out[tag + "_bindex"] = np.asarray([b["bucket_index"] for b in bk], np.int32)
            out[tag + "_bnd"] = np.asarray([b["bnd"] for b in bk], np.float64)
            out[tag + "_count"] = np.asarray([len(b["rows"]) for b in bk], np.int32)
        for tag, fn in (("scan_pre", "s0001_scan_pre"), ("scan_post", "s0001_scan_post"), ("adapt_pre", "s0002_adapt_pre"),
                        ("adapt_post", "s0002_adapt_post"), ("pghost_post", "s0002_pghost_post")):
            bk = refio.read_bucketsx(os.path.join(dump, fn + "_bucketsx.bin"))
            out[tag + "_mincrd"] = np.asarray([b["mincrd"] for b in bk])
            out[tag + "_attr"] = np.asarray([[b["type"], b["has_involved"], b["active"], b["guest"], b["particles_type"],
                                             b["erupt_flag"]] for b in bk], np.int32)
            if tag == "scan_pre":
                out["bindex"] = np.asarray([b["bucket_index"] for b in bk], np.int32)
                out["bnd"] = np.asarray([b["bnd"] for b in bk], np.float64)
                out["neigh_proc"] = np.asarray([b["neigh_proc"] for b in bk], np.int32)
                out["plist_off"] = np.cumsum([0] + [len(b["rows"]) for b in bk]).astype(np.int64)
                out["plist_rows"] = np.concatenate([b["rows"] for b in bk]).astype(np.int32)
        g = refio.read_grow_params(os.path.join(dump, "s0002_grow_params.bin"))
        out["param_step"] = np.asarray([g["step"], g["atmosphere_type"]], np.int32)
        out["param_const"] = np.asarray([g["mass"], g["smlen"], g["gamma"]])
        out["param_keybox"] = np.concatenate([g["mindom"], g["maxdom"]])
        out["param_domain"] = np.concatenate([g["Ll"], g["Lu"]])
        out["meteo_z"] = g["meteo_z"]
        out["meteo"] = g["meteo"]
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "adapt", out["adapt"], "rows", len(t1), "turned", int(ch.sum()), "added", int(isnew.sum()), "wall ghosts",
              int(isnew4.sum()), "buckets", len(out["wghost_pre_mincrd"]), "->",
              os.path.getsize(os.path.join(HERE, name + ".npz")), "B")


ERUPT_PARAMS = ("time", "dt", "step", "bot", "sml_of_phase2", "mass_of_phase2", "rv", "pos_v0", "pos_v1", "Vv0", "ev0", "rhov",
                "pv0", "gamma_v", "sndspd", "vel_prof", "ratio_sml_dx", "kmin0", "kmin1", "kmin2", "kmax0", "kmax1", "kmax2",
                "iferupt")


def read_erupt_params(fn):
    raw = open(fn, "rb").read()
    v = np.frombuffer(raw, np.float64, 24)
    nb = int(np.frombuffer(raw, np.int64, 1, 192)[0])
    rec = np.frombuffer(raw, np.dtype([("mincrd", "<f8", 3), ("maxcrd", "<f8", 3), ("guest", "<i4"), ("pad", "<i4")]), nb, 200)
    return v.copy(), rec.copy()


def golden_erupt(name="erupt3d", variant="3d_plume", edits=None, steps=(2, 3), simdata="simulation.data"):
    """add_new_erupt (setup_erupt.cc:390-670) of the small 3-D plume at time steps 2 and 3 (step 1 adds nothing: dt is
    still shorter than the injection period of every duct column): what the function reads (arguments, vent constants,
    key range of the particle table, erupt buckets) and the particles it added, in THashTable order."""
    edits = SMALL_3D if edits is None else edits
    with tempfile.TemporaryDirectory() as top:
        wd = os.path.join(top, "run")
        os.makedirs(wd)
        shutil.copytree(os.path.join(REF, "Meteo_data"), os.path.join(top, "Meteo_data"))
        shutil.copy(os.path.join(REF, "bin", "scale.data"), wd)
        txt = open(os.path.join(REF, "bin", simdata)).read()
        for a, b in edits:
            assert a in txt, a
            txt = txt.replace(a, b)
        open(os.path.join(wd, "simulation.data"), "w").write(txt)
        smlen = [l for l in txt.splitlines() if l.startswith("smoothing length")][0].split(":")[1]
        subprocess.run([os.path.join(REFBIN, variant, "preprocess"), "1", smlen, "0"], cwd=wd, check=True,
                       stdout=subprocess.DEVNULL)
        dump = os.path.join(wd, "dump")
        os.makedirs(dump)
        run_ref(variant, wd, dict(PSPH_DUMP_DIR=dump, PSPH_DUMP_STEPS="%d-%d" % (min(steps), max(steps)),
                                  PSPH_MAX_STEPS=max(steps)))
        out = dict(param_names=np.asarray(ERUPT_PARAMS))
        for st in steps:
            _, a = refio.read_dump(os.path.join(dump, "s%04d_rebin_post.bin" % st))
            _, b = refio.read_dump(os.path.join(dump, "s%04d_erupt_post.bin" % st))
            v, cells = read_erupt_params(os.path.join(dump, "s%04d_erupt_params.bin" % st))
            ka = {tuple(k): i for i, k in enumerate(a["key"])}
            new = np.asarray([i for i, k in enumerate(b["key"]) if tuple(k) not in ka], np.int64)
            old = np.asarray([i for i, k in enumerate(b["key"]) if tuple(k) in ka], np.int64)
            assert len(old) == len(a) and all(np.array_equal(a[f], b[f][old]) for f in a.dtype.names), "add_new_erupt touched an old row"
            out["s%d_params" % st] = v
            out["s%d_cell_min" % st] = cells["mincrd"]
            out["s%d_cell_max" % st] = cells["maxcrd"]
            out["s%d_cell_guest" % st] = cells["guest"]
            out["s%d_new" % st] = b[new]
            out["s%d_nold" % st] = np.asarray([len(a)])
            print(name, "step", st, "added", len(new), "of", len(b), "erupt buckets", len(cells))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


SMALL_3D = [("initial domain:-2000", "initial domain:-350"), ("initial domain:2000", "initial domain:350"),
            ("initial domain:2700", "initial domain:2400")]

JPUE_3D = [("initial domain:-100", "initial domain:-60"), ("initial domain:100", "initial domain:60"),
           ("initial domain:140", "initial domain:100")]

if __name__ == "__main__":
    which = sys.argv[1:] or ["sod_sph", "sod_gsph", "sod0_gsph", "shu_sph", "shu_gsph", "blast_gsph",
                             "plume3d_small", "plume3d_small_gsph"]
    table = {
        "sod_sph": lambda: golden_1d("sod_sph", "1d_sod_sph", "simulation_1D_shocktube_Sod.data", 1000),
        "sod_gsph": lambda: golden_1d("sod_gsph", "1d_sod_gsph", "simulation_1D_shocktube_Sod.data", 1000),
        "sod0_gsph": lambda: golden_1d("sod0_gsph", "1d_sod0_gsph", "simulation_1D_shocktube_Sod_GSPH.data", 300),
        "shu_sph": lambda: golden_1d("shu_sph", "1d_shu_sph", "simulation_1D_shocktube_shu-Osher.data", 200),
        "shu_gsph": lambda: golden_1d("shu_gsph", "1d_shu_gsph", "simulation_1D_shocktube_shu-Osher.data", 200),
        # (the reference's own plain-SPH build aborts on the strong blast at step 4: assert c>0, timestep.cc:167)
        "blast_gsph": lambda: golden_1d("blast_gsph", "1d_blast_gsph", "simulation_1D_shocktube_Strong-Blast.data", 300),
        "grow3d": lambda: golden_grow(),
        "erupt3d": lambda: golden_erupt(),
        "erupt3d_jpue": lambda: golden_erupt("erupt3d_jpue", "3d_jpue", JPUE_3D, steps=(2, 3, 4),
                                             simdata="simulation_JPUE_incomp_back.data"),
        "plume3d_small": lambda: golden_3d("plume3d_small", "3d_plume", SMALL_3D),
        "plume3d_small_gsph": lambda: golden_3d("plume3d_small_gsph", "3d_plume_gsph", SMALL_3D, list_stride=16),
        "plume3d_sml": lambda: golden_3d_sml(),
        # the USE_GSPH 1 build over two steps with adaptive_sml in each: step 2 has velocities, velocity gradients and
        # per-particle smoothing lengths in the reconstructed Riemann states
        "plume3d_gsph2": lambda: golden_3d_sml("plume3d_gsph2", "3d_plume_gsph"),
        # the same two steps with the reference's own SML_UPDATE_INT: uniform h, and a step (2) that follows a domain growth
        "plume3d_grow": lambda: golden_3d_sml("plume3d_grow", sml_int=0),
        # BASELINE config 2: JPUE weakly-compressible jet (Tait EOS, g = 0), the reference's own input file
        "jpue3d": lambda: golden_3d("jpue3d", "3d_jpue", JPUE_3D, list_stride=16, simdata="simulation_JPUE_incomp_back.data"),
    }
    for w in which:
        table[w]()
