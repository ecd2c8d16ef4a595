#!/usr/bin/env python3
"""Build the UNMODIFIED reference (read in place from /root/reference) into oracle/_ref/.

Test infrastructure only (see oracle/README.md). Outputs go to oracle/_ref/<variant>/
(git-ignored, but shipped to the GPU box by gpurun like any built artefact).
No reference source file is copied: every TU is compiled from /root/reference/src
where it lies. What we add is
  * stubs/mpi.h, stubs/hdf5.h, stubs/cfg{1d,3d}/config.h  (single-rank MPI, no HDF5),
  * gh5_flat.cc   (stand-in for src/plume/hdf5calls.cc: flat record files),
  * ref_hooks.cc  (textually includes src/plume/main.cc with its stage calls renamed
                   to dumping/timing wrappers; the wrappers call the reference functions).

Variants (compile-time option sets):
  1d_sod_sph     CODE_DIMENSION=1 USE_GSPH=0 SHOCK_TUBE_TESTS=1
  1d_sod_gsph    CODE_DIMENSION=1 (defaults: GSPH/HLLC)  SHOCK_TUBE_TESTS=1
  1d_sod0_gsph   CODE_DIMENSION=1 defaults, SHOCK_TUBE_TESTS=0 (…_Sod_GSPH.data)
  1d_shu_sph / 1d_shu_gsph       SHOCK_TUBE_TESTS=2
  1d_blast_sph / 1d_blast_gsph   SHOCK_TUBE_TESTS=4
  3d_plume       CODE_DIMENSION=3 defaults (+ preprocess)

Usage: build_ref.py [variant ...]   (default: all);  -j N parallel compiles.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PSPH_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(os.path.dirname(HERE), "_ref")

COMMON = ["-O2", "-std=c++11", "-DHAVE_CONFIG_H", "-w", "-fpermissive", "-fno-stack-protector",
          "-include", "cmath",
          "-include", f"{SRC}/datstr/bucket.h", "-include", f"{SRC}/datstr/particle.h",
          "-include", f"{SRC}/plume/parameters.h", "-include", f"{SRC}/datstr/hilbert.h"]
INCS = [f"-I{SRC}/datstr", f"-I{SRC}/plume", f"-I{SRC}/multiproc", f"-I{SRC}/preproc", f"-I{SRC}/debug"]

TU_1D = ["datstr/particle.cc", "datstr/thashtab.cc", "datstr/hashtab.cc", "datstr/meteo.cc", "datstr/bucket.cc",
         "plume/sph_lib.cc", "plume/neighs.cc", "plume/density_massfrac.cc", "plume/mom_engr_update.cc",
         "plume/calculate_gradient.cc", "plume/smooth_velocity.cc", "plume/update_pos.cc", "plume/timestep.cc",
         "plume/add_air.cc", "plume/dataread.cc", "plume/write_output.cc", "plume/h5part_writer.cc",
         "plume/matlab_writer.cc"]


def tu_3d():
    out = []
    for d in ("datstr", "multiproc", "plume"):
        for f in sorted(os.listdir(os.path.join(SRC, d))):
            if f.endswith(".cc") and f not in ("hdf5calls.cc", "mpi_ind_struct.cc", "main.cc"):
                out.append(f"{d}/{f}")
    return out


VARIANTS = {
    "1d_sod_sph": dict(dim=1, defs=["-DUSE_GSPH=0", "-DSHOCK_TUBE_TESTS=1"]),
    "1d_sod_gsph": dict(dim=1, defs=["-DSHOCK_TUBE_TESTS=1"]),
    "1d_sod0_gsph": dict(dim=1, defs=["-DSHOCK_TUBE_TESTS=0"]),
    "1d_shu_sph": dict(dim=1, defs=["-DUSE_GSPH=0", "-DSHOCK_TUBE_TESTS=2"]),
    "1d_shu_gsph": dict(dim=1, defs=["-DSHOCK_TUBE_TESTS=2"]),
    "1d_blast_sph": dict(dim=1, defs=["-DUSE_GSPH=0", "-DSHOCK_TUBE_TESTS=4"]),
    "1d_blast_gsph": dict(dim=1, defs=["-DSHOCK_TUBE_TESTS=4"]),
    "3d_plume": dict(dim=3, defs=[]),
    "3d_plume_gsph": dict(dim=3, defs=["-DUSE_GSPH=1"]),
    # JPUE weakly-compressible jet (user guide "A JPUE simulation"): Tait equation of state, no gravity,
    # and the parameters.h block at :404-562 instead of the Eyjafjallajokull one (:875-1046)
    "3d_jpue": dict(dim=3, defs=["-DFLUID_COMPRESSIBILITY=1", "-DATMOSPHERE_TYPE=2"], params="jpue"),
}


def patched_parameters(odir, which):
    """parameters.h keeps its alternative parameter sets as commented-out blocks; selecting one
    means editing the header.  /root/reference is read-only and must stay unmodified, so a patched COPY is
    generated into the build output (oracle/_ref/<variant>/include/parameters.h, git-ignored) and force-
    included first; its include guard then shadows the original.  Edits: uncomment the JPUE block, put the
    default (Eyjafjallajokull) block under #if 0."""
    assert which == "jpue"
    lines = open(os.path.join(SRC, "plume", "parameters.h"), encoding="latin-1").read().split("\n")
    find = lambda text, start=0: next(i for i in range(start, len(lines)) if text in lines[i])
    j0 = find("These parameters are for JPUE of incompressible flow")
    j1 = find("Parameters for running RP1 with sml1=200", j0)
    e0 = find("These parameters are for Eyjafjallaj", j1)
    e1 = find("#else //ASH_SIMULATION", e0)
    # block comments start one line above the text line and end two lines below it
    for i in range(j0 + 2, j1 - 1):
        rest = lines[i][2:].strip() if lines[i].startswith("//") else None
        if rest is not None and (rest == "" or rest.startswith(("const", "#", "//"))):
            lines[i] = lines[i][2:]
    lines.insert(e1, "#endif /* default parameter block disabled: JPUE selected */")
    lines.insert(e0 + 2, "#if 0")
    inc = os.path.join(odir, "include")
    os.makedirs(inc, exist_ok=True)
    out = os.path.join(inc, "parameters.h")
    open(out, "w", encoding="latin-1").write("\n".join(lines))
    return out


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(1)


def build(name, jobs):
    v = VARIANTS[name]
    odir = os.path.join(OUT, name)
    os.makedirs(os.path.join(odir, "obj"), exist_ok=True)
    cfg = os.path.join(HERE, "stubs", "cfg1d" if v["dim"] == 1 else "cfg3d")
    common = list(COMMON)
    if v.get("params"):
        ph = patched_parameters(odir, v["params"])
        # first of all force-includes: a -DDEBUG build (preprocess) reaches parameters.h through other headers, and
        # whichever copy is seen first defines the include guard
        k = common.index(f"{SRC}/plume/parameters.h")
        del common[k - 1:k + 1]
        common[common.index("-include"):common.index("-include")] = ["-include", "cmath", "-include", ph]
    flags = common + v["defs"] + [f"-I{cfg}", f"-I{HERE}/stubs"] + INCS
    tus = TU_1D if v["dim"] == 1 else tu_3d()
    objs = []
    cmds = []
    for tu in tus:
        o = os.path.join(odir, "obj", tu.replace("/", "_") + ".o")
        objs.append(o)
        cmds.append(["g++"] + flags + ["-c", os.path.join(SRC, tu), "-o", o])
    o = os.path.join(odir, "obj", "gh5_flat.o")
    objs.append(o)
    cmds.append(["g++"] + flags + ["-c", os.path.join(HERE, "gh5_flat.cc"), "-o", o])
    o = os.path.join(odir, "obj", "hilbert.o")
    objs.append(o)
    cmds.append(["gcc", "-O2", "-w", "-c", os.path.join(SRC, "datstr/hilbert.c"), "-o", o])
    # two drivers: plain (main.cc as is) and hooked (main.cc included by ref_hooks.cc)
    o_main = os.path.join(odir, "obj", "main_plain.o")
    cmds.append(["g++"] + flags + ["-c", os.path.join(SRC, "plume/main.cc"), "-o", o_main])
    o_hook = os.path.join(odir, "obj", "main_hooked.o")
    cmds.append(["g++"] + flags + [f'-DPSPH_REF_MAIN_CC="{SRC}/plume/main.cc"',
                                  "-c", os.path.join(HERE, "ref_hooks.cc"), "-o", o_hook])
    with ThreadPoolExecutor(jobs) as ex:
        list(ex.map(run, cmds))
    run(["g++", "-o", os.path.join(odir, "particler")] + objs + [o_main, "-lm"])
    run(["g++", "-o", os.path.join(odir, "particler_hooked")] + objs + [o_hook, "-lm"])
    # third driver (3-D SPH builds, when the product library is built): main.cc with its stage calls forwarded to
    # libpsph.so (ref_b200.cc) — the drop-in as a program; needs a GPU to run
    libdir = os.path.join(os.path.dirname(os.path.dirname(HERE)), "plume_sph_b200")
    if v["dim"] == 3 and "-DUSE_GSPH=1" not in v["defs"] and not v.get("gsph") and os.path.exists(os.path.join(libdir, "libpsph.so")):
        o_b200 = os.path.join(odir, "obj", "main_b200.o")
        inc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")
        run(["g++"] + flags + [f"-I{inc}", f'-DPSPH_REF_MAIN_CC="{SRC}/plume/main.cc"',
                               "-c", os.path.join(HERE, "ref_b200.cc"), "-o", o_b200])
        run(["g++", "-o", os.path.join(odir, "particler_b200")] + objs + [o_b200, "-lm", f"-L{libdir}", "-lpsph",
             "-Wl,-rpath,$ORIGIN/../../../plume_sph_b200", "-Wl,--allow-shlib-undefined"])
    if v["dim"] == 3:
        # preprocess only compiles with -DDEBUG
        pre = []
        pcmds = []
        for tu in ("preproc/preprocess.cc", "preproc/createfunky.cc"):
            o = os.path.join(odir, "obj", "pre_" + os.path.basename(tu) + ".o")
            pre.append(o)
            pcmds.append(["g++"] + flags + ["-DDEBUG", "-c", os.path.join(SRC, tu), "-o", o])
        with ThreadPoolExecutor(jobs) as ex:
            list(ex.map(run, pcmds))
        run(["g++", "-o", os.path.join(odir, "preprocess")] + pre +
            [os.path.join(odir, "obj", "gh5_flat.o"), os.path.join(odir, "obj", "hilbert.o"), "-lm"])
    print("built", name, "->", odir)


def main():
    args = sys.argv[1:]
    jobs = os.cpu_count() or 4
    if "-j" in args:
        i = args.index("-j")
        jobs = int(args[i + 1])
        del args[i:i + 2]
    names = args or list(VARIANTS)
    if not os.path.isdir(SRC):
        print("reference sources not present at", SRC, "- nothing to build (prebuilt oracle/_ref is used as is)")
        return 0
    for n in names:
        build(n, jobs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
