#!/usr/bin/env python3
"""bench.py — particle-updates/s of the Plume-SPH per-time-step hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--edge EDGE] [--impl b200|reference]

Workload (BASELINE.json config 4): the synthetic 3-D particle cube, 3-D SPH
defaults (artificial-viscosity SPH + LANS velocity filter + energy smoothing, ideal-gas
mixture EOS, g = 0): EDGE^3 real particles on a jittered unit lattice (h = 1), a central
sphere of erupted material, one bucket-thick (8 h) shell of static pressure ghosts.
One "step" = one full time step of the hot path in the reference's order
(main.cc:217-408): timestep -> neighbour search (cell-list build) -> mom_engr_update ->
smooth_density -> smooth_velocity -> update_pos -> update_particles_in_bucket.
`value` = real particles advanced per second, whole job, inputs resident in HBM; `e2e` =
the same through the C ABI with HOST rows: psph_upload (pinned host AoS -> device SoA) +
psph_step + psph_download every step.

--impl reference: the UNMODIFIED reference compiled in oracle/_ref (its own 3-D plume case,
small initial domain) timed on the host cores; see oracle/ref_build/ref_run.py.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle-updates/sec (3D SPH hot path, synthetic particle cube)"
UNIT = "particle-updates/s"
# algorithmic bytes per real-particle-update of each pair gather
ALG_BYTES = {"gather_mom_engr": 228, "gather_density": 148, "gather_velocity": 100}
ALG_BYTES_STEP = 1184
# kernel behind each timer, per regime (3-D SPH defaults)
KERNEL_OF = {"uniform_h": {"gather_mom_engr": "k_tile_gather<MomStage<3>, WM=1>", "gather_density": "k_dens_near",
                           "gather_velocity": "k_tile_gather<VelStage<3>, WM=2>"},
             "mixed_h": {"gather_mom_engr": "k_mix_gather<MomStage<3>>", "gather_density": "k_dens_near<false>",
                         "gather_velocity": "k_mix_gather<VelStage<3>>"}}
# Figures that come from ncu captures, NOT from this run: reported under roofline.from_profile with their source file
# and problem size, never mixed into the live numbers.
PROFILE_FIGURES = os.path.join(ROOT, "profiles", "ncu_figures.json")


def measured_peaks():
    try:
        j = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(j["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# device-side generator of the synthetic cube (same shape as plume_sph_b200.cases.cube, which
# builds host rows for the tests)
# ------------------------------------------------------------------------------------------
def cube_columns(torch, cfg, n, dev, seed=20260101, smlen=1.0, lo=None, hi=None):
    """SoA columns (PSPH_COL_* order) + bc/phase of the block of lattice sites [lo, hi) of the
    (n + 16)^3 cube-with-shell; default = everything."""
    from plume_sph_b200 import cases
    g = 8
    lo = lo if lo is not None else (-g, -g, -g)
    hi = hi if hi is not None else (n + g, n + g, n + g)
    ax = [torch.arange(lo[k], hi[k], device=dev, dtype=torch.float64) for k in range(3)]
    I, J, K = torch.meshgrid(ax[0], ax[1], ax[2], indexing="ij")
    I, J, K = I.reshape(-1), J.reshape(-1), K.reshape(-1)
    N = I.numel()
    real = (I >= 0) & (I < n) & (J >= 0) & (J < n) & (K >= 0) & (K < n)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed + 7919 * (lo[0] + g))
    jit = (torch.rand((3, N), device=dev, dtype=torch.float64, generator=gen) * 2 - 1) * (0.1 * smlen)
    pos = [(c + 0.5) * smlen + torch.where(real, jit[k], torch.zeros_like(c)) for k, c in enumerate((I, J, K))]
    ctr = 0.5 * n * smlen
    r2 = (pos[0] - ctr) ** 2 + (pos[1] - ctr) ** 2 + (pos[2] - ctr) ** 2
    ph2 = real & (r2 < (n * smlen / 8.0) ** 2)
    ks = ph2.to(torch.float64)
    rho = torch.full((N,), cases.RHO0, device=dev, dtype=torch.float64)
    e_air = cases.P0 / (0.4 * cases.RHO0)
    ng = cfg.ng0
    cvm2 = (1 - ng) * cfg.Cvs + ng * cfg.Cvg
    p_air = cfg.Ra * cases.RHO0 * e_air / cfg.Cva
    e = torch.where(ph2, torch.full_like(rho, p_air * cvm2 / (cfg.Rg * ng * cases.RHO0)), torch.full_like(rho, e_air))
    # Particle::update_second_var (particle.cc:528-604), ideal-gas mixture
    na, ngk = 1 - ks, ng * ks
    ns = ks - ngk
    Cvm = ns * cfg.Cvs + ngk * cfg.Cvg + na * cfg.Cva
    Rm = ngk * cfg.Rg + na * cfg.Ra
    gamma = 1 + Rm / Cvm
    desmg = rho * (na + ngk)
    p = (Rm / (ngk + na)) * desmg * e / Cvm
    c = torch.sqrt(gamma * p / desmg)
    c_air = float((cfg.Ra / cfg.Cva + 1) * p_air / cases.RHO0) ** 0.5
    vj = (torch.rand((3, N), device=dev, dtype=torch.float64, generator=gen) * 2 - 1) * (0.05 * c_air)
    vel = [torch.where(real, vj[k], torch.zeros_like(rho)) for k in range(3)]
    mass = torch.full_like(rho, cases.RHO0 * smlen ** 3)
    h = torch.full_like(rho, smlen)
    cols = pos + vel + [v.clone() for v in vel] + [rho, e, mass, h, ks, p, gamma, c, e / Cvm, Cvm]
    bc = torch.where(real, 100, 1).to(torch.int32)
    phase = torch.where(ph2, 2, 1).to(torch.int32)
    return [x.contiguous() for x in cols], bc.contiguous(), phase.contiguous(), int(real.sum().item())


def run_b200(args):
    import torch
    import torch.distributed as dist
    import plume_sph_b200 as psph
    from plume_sph_b200 import cases
    from plume_sph_b200.solver import ORDER_DEVICE

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    if world > 1 and args.weak:   # weak scaling: the global cube holds world x the per-GPU cube
        n = int(round(args.n * world ** (1.0 / 3.0)))
    cfg = cases.cube_config(n)
    if args.sml_int > 0:               # adaptive_sml every k-th step (reference default 50): measures the regime after it
        cfg.sml_update_int = args.sml_int
    cfg.device = local
    cfg.rank, cfg.nranks = rank, world
    from plume_sph_b200.slabs import setup_cube_ranks  # slab decomposition across ranks (no-op for 1 rank)
    s, n_real_local, n_rows_local = setup_cube_ranks(torch, dist, psph, cfg, n, dev, rank, world, cube_columns)

    stream = torch.cuda.ExternalStream(s.stream(), device=dev)
    peak, peak_src = measured_peaks()

    def barrier():
        s.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def measure(regime, warmup, steps, sample_clocks):
        """W untimed + K timed steps of psph_run (CUDA events on the context stream, barrier + synchronize on both sides,
        max over ranks), then a separate pass with events around every kernel for the per-kernel / per-stage times."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(warmup):
            s.run(1, erupting=0)
        barrier()
        s.launch_count(reset=True)
        clocks = ClockSampler(local)
        if rank == 0 and sample_clocks:
            clocks.start()
        ev0.record(stream)
        for _ in range(steps):
            s.run(1, erupting=0)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = s.launch_count(reset=True)
        clk = clocks.stop() if (rank == 0 and sample_clocks) else None
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        cnt = torch.tensor([float(n_real_local), float(launches)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms = float(t.item())
        n_real = int(cnt[0].item())
        # per-kernel device times: separate pass, not the timed region
        s.set_timing(True)
        ksteps = max(1, min(2, steps))
        for _ in range(ksteps):
            s.run(1, erupting=0)
        s.sync()
        kms, sms = s.kernel_ms(), s.stage_ms()
        s.set_timing(False)
        pstats = s.pair_stats()
        dom = max(ALG_BYTES, key=lambda k: kms[k][0])
        dom_ms = kms[dom][0] / max(kms[dom][1], 1)
        achieved = ALG_BYTES[dom] * n_real_local / (dom_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": KERNEL_OF[regime][dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5),
                "traffic": None,   # DRAM bytes per launch come from ncu captures only: see from_profile
                "alg_bytes_per_update": ALG_BYTES[dom], "updates_per_launch": n_real_local,
                "ms_per_launch": round(dom_ms, 4), "peak_source": peak_src,
                "step_hbm": {"alg_bytes_per_update": ALG_BYTES_STEP,
                             "achieved_GBps": round(ALG_BYTES_STEP * n_real / (ms / steps * 1e-3) / 1e9, 2)},
                "stage_ms_per_step": {k: round(v / ksteps, 4) for k, v in sms.items() if v > 0},
                "kernel_ms_per_launch": {k: round(v[0] / max(v[1], 1), 4) for k, v in kms.items() if v[1] > 0}}
        return {"value": round(n_real * steps / (ms * 1e-3), 1), "ms_per_step": round(ms / steps, 4), "steps": steps,
                "warmup": warmup, "n_real": n_real, "gpu_launches": int(cnt[1].item()), "clocks": clk, "roofline": roof,
                "pair_path_rank0": pstats}

    regimes = {}
    # ---- regime 1: uniform h (time steps 1-49 of a default run: before the first adaptive_sml call, main.cc:319-322)
    if args.regime in ("both", "uniform"):
        regimes["uniform_h"] = measure("uniform_h", args.warmup, args.steps, args.headline == "uniform")
    # ---- regime 2: per-particle h (every step after the first adaptive_sml call: the production regime).  The bench
    # makes the call the reference makes at step 50 (search_neighs_consth + adaptive_sml) and keeps stepping.
    sml_ms = None
    if args.regime in ("both", "mixed"):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.search_neighs_consth()
        barrier()
        ev0.record(stream)
        s.adaptive_sml()
        ev1.record(stream)
        barrier()
        sml_ms = ev0.elapsed_time(ev1)
        if world > 1:
            s.exchange_ghosts()                                    # main.cc:326
        regimes["mixed_h"] = measure("mixed_h", args.warmup, args.steps, args.headline == "mixed")
        regimes["mixed_h"]["adaptive_sml_ms_per_call"] = round(sml_ms, 3)
    head = regimes["mixed_h" if args.headline == "mixed" else "uniform_h"]
    value, ms_step, n_real, clk = head["value"], head["ms_per_step"], head["n_real"], head["clocks"]
    roofline = dict(head["roofline"])
    roofline["note"] = ("FP64 pair loop over per-step pair lists, partners' records staged in shared memory: bound by the FP64 "
                        "pipe and the issue slots, not by HBM (SURVEY.md §8d): the HBM fraction by ALGORITHMIC bytes is low by "
                        "construction; from_profile carries the ncu figures (FP64 pipe, DRAM bytes per launch) with their source")
    try:
        roofline["from_profile"] = json.load(open(PROFILE_FIGURES))
    except Exception:
        roofline["from_profile"] = None
    # the roofline that does bind the pair sums: the FP64 pipe.  Peak = the DFMA micro-benchmark of this pool's B200
    # (profiles/r02d_dfma_peak.json); the kernel's share of it is an ncu figure (a profile run, not this run) and says so.
    fp = roofline["from_profile"] or {}
    kname = roofline["kernel"].replace(", WM=1>", ",WM=1>").replace(", WM=2>", ",WM=2>")
    kfig = fp.get(kname) or {}
    pk = fp.get("_fp64_peak") or {}
    if kfig.get("fp64_pipe_pct") is not None and pk:
        roofline["fp64"] = {"bound": "fp64 pipe", "peak_dfma_per_s": pk["dfma_per_s"], "peak_tflops": pk["fp64_fma_tflops"],
                            "peak_source": pk["source"], "kernel": roofline["kernel"],
                            "pipe_busy_pct_from_profile": kfig["fp64_pipe_pct"], "profile_source": kfig.get("fp64_source"),
                            "what": "share of the FP64 pipe's issue capacity the dominant kernel uses (ncu "
                                    "sm__inst_executed_pipe_fp64, committed capture at the size named in profile_source); the "
                                    "time of THIS run's launch is roofline.ms_per_launch"}
    bad = s.errors()

    # ---- end to end through the C ABI with host rows (pinned): upload + step + download per step
    e2e = None
    if not args.no_e2e:
        # rows as an SPH build of the reference exchanges them: ParticlePack carries no gradients unless USE_GSPH
        # (pack_data.h:68-75) -> psph_particle without its gradient block, 208 B per row (PSPH_ROW_BYTES_SPH)
        host = s.host_rows(n_rows_local + 1024, dtype=psph.PARTICLE_SPH_DTYPE)
        out = s.download(ORDER_DEVICE, out=host)
        nrow = len(out)
        rowb = host.dtype.itemsize
        ke = max(1, min(args.steps, 3))
        import ctypes as C
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            s._ck(s.L.psph_upload_rows(s.h, host.ctypes.data, nrow, rowb), "psph_upload_rows")
            s.step(0.0, 0)
            got = C.c_size_t(0)
            s._ck(s.L.psph_download_rows(s.h, host.ctypes.data, host.size, C.byref(got), ORDER_DEVICE, rowb), "psph_download_rows")
            nrow = got.value
        barrier()
        te = time.perf_counter() - t0
        tt = torch.tensor([te], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        e2e = {"value": round(n_real * ke / te, 1), "unit": UNIT, "regime": "mixed_h" if args.headline == "mixed" else "uniform_h", "h2d_bytes_per_step": nrow * rowb,
               "d2h_bytes_per_step": nrow * rowb, "steps": ke,
               "row_bytes": rowb,
               "what": "psph_upload_rows(host AoS rows, pinned; SPH rows = psph_particle without the GSPH gradient block) + "
                       "psph_step + psph_download_rows(host AoS rows) every step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_port(args.cpu_n)
    s.close()
    s = None
    plume = None
    if not args.no_plume:
        try:
            plume = run_plume(args, torch, dist, psph, rank, world, local, dev)
        except Exception as ex:                            # the cube line must not die with the companion case
            plume = {"error": "%s: %s" % (type(ex).__name__, ex)} if rank == 0 else None

    if rank == 0:
        other = {k: {kk: vv for kk, vv in v.items() if kk != "clocks"} for k, v in regimes.items()}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak" if args.weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic 3-D particle cube, %d^3 real (%d^3 per GPU) + 8h pressure-ghost shell, "
                                       "3-D SPH defaults (art. viscosity + LANS + energy smoothing), g=0" % (n, args.n),
                           "regime": ("per-particle smoothing lengths: the state after adaptive_sml (every step of a default run "
                                      "from step 50 on)" if args.headline == "mixed" else
                                      "uniform smoothing length: steps 1-49 of a default run, before the first adaptive_sml call"),
                           "real_particles": n_real, "rows_rank0": n_rows_local, "edge": n,
                           "l2": "inputs (%.1f GB of SoA columns) far larger than the 126 MB L2" %
                                 (n_rows_local * 700 / 1e9),
                           "parallelism": "key-range slabs x%d" % world if world > 1 else "1 GPU"},
                "roofline": roofline, "gpu_launches": head["gpu_launches"], "clocks": clk, "e2e": e2e,
                "regimes": other, "invariant_violations": bad}
        if cpu:
            line["cpu_baseline"] = cpu
        if plume:
            line["plume"] = plume
        print(json.dumps(line), flush=True)
    if s is not None:
        s.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# The reference's own 3-D plume case, like for like (BASELINE.json config 3): the unmodified reference driver sets the case
# up (input parsing, background mesh, Meteo atmosphere, air / pressure-ghost / wall-ghost / vent particles, image search —
# everything north_star leaves with the driver) and is timed on the host cores; the state it starts its loop from is then
# advanced by libpsph.so through the same per-step calls (psph_step, psph_add_new_erupt, psph_apply_bcond,
# psph_scan_outside_layer), strong-scaled over the slabs when there are several GPUs.
# ---------------------------------------------------------------------------------------------------------------------
DUMP_ROW = [("key", "<u4", 3), ("bc_type", "<i4"), ("phase", "<i4"), ("involved", "<i4"), ("guest", "<i4"), ("pad", "<i4"),
            ("mass", "<f8"), ("smlen", "<f8"), ("mass_frac", "<f8"), ("pressure", "<f8"), ("gamma", "<f8"), ("sound_speed", "<f8"),
            ("temperature", "<f8"), ("specific_heat_p", "<f8"), ("phase_density", "<f8", 2), ("coord", "<f8", 3), ("vel", "<f8", 3),
            ("smoothed_v", "<f8", 3), ("rho", "<f8"), ("engr", "<f8"), ("smoothed_e", "<f8"), ("grad", "<f8", 15)]


def load_reference_setup(dump, psph, cfg):
    """the state files the hooked reference driver wrote (oracle/ref_build/ref_hooks.cc formats) -> what the C ABI takes:
    particle rows, bucket attributes, wall-ghost images, vent record + erupt buckets"""
    import struct
    import numpy as np
    import ctypes as C
    with open(os.path.join(dump, "s0000_init_post.bin"), "rb") as f:
        f.read(12)
        n, = struct.unpack("<q", f.read(8))
        f.read(16)
        d = np.fromfile(f, dtype=np.dtype(DUMP_ROW), count=n)
    rows = np.zeros(n, psph.PARTICLE_DTYPE)
    for a, b in (("key", "key"), ("bc_type", "bc_type"), ("phase_num", "phase"), ("involved", "involved"), ("guest", "guest"),
                 ("mass", "mass"), ("smlen", "smlen"), ("mass_frac", "mass_frac"), ("pressure", "pressure"), ("gamma", "gamma"),
                 ("sound_speed", "sound_speed"), ("temperature", "temperature"), ("specific_heat_p", "specific_heat_p"),
                 ("phase_density", "phase_density"), ("coords", "coord"), ("vel", "vel"), ("smoothed_v", "smoothed_v"),
                 ("rho", "rho"), ("engr", "engr"), ("smoothed_e", "smoothed_e")):
        rows[a] = d[b]
    n1, n2 = int(cfg.ncell[1]), int(cfg.ncell[2])
    attr = np.zeros(int(cfg.ncell[0]) * n1 * n2, psph.CELL_ATTR_DTYPE)
    key2cell = {}
    mind = np.asarray([cfg.mindom[k] for k in range(3)])
    with open(os.path.join(dump, "s0000_init_buckets.bin"), "rb") as f:
        nb, = struct.unpack("<q", f.read(8))
        for _ in range(nb):
            key = struct.unpack("<II", f.read(8))
            mn = np.asarray(struct.unpack("<ddd", f.read(24)))
            f.read(24)
            typ, inv, act, guest, npart = struct.unpack("<iiiii", f.read(20))
            bidx = struct.unpack("<6i", f.read(24))
            bnd = struct.unpack("<6d", f.read(48))
            f.read(4 * npart)
            idx = np.rint((mn * cfg.len_scale - mind) / cfg.delta).astype(np.int64)
            lin = int((idx[0] * n1 + idx[1]) * n2 + idx[2])
            key2cell[key] = lin
            attr[lin] = (typ, inv, act, bidx, 0, bnd)
    idt = np.dtype([("buckproc", "<i4"), ("partproc", "<i4"), ("bucket_key", "<u4", 2), ("ghost_key", "<u4", 3), ("coord", "<f8", 3)])
    with open(os.path.join(dump, "s0001_images.bin"), "rb") as f:
        ni, = struct.unpack("<q", f.read(8))
        im = np.fromfile(f, dtype=idt, count=ni)
    img = np.zeros(ni, psph.IMAGE_DTYPE)
    img["ghost_key"] = im["ghost_key"]
    img["cell"] = [key2cell[tuple(int(x) for x in k)] for k in im["bucket_key"]]
    img["coord"] = im["coord"]
    raw = open(os.path.join(dump, "s0001_erupt_params.bin"), "rb").read()
    v = np.frombuffer(raw, np.float64, 24)
    nbx = int(np.frombuffer(raw, np.int64, 1, 192)[0])
    rec = np.frombuffer(raw, np.dtype([("mincrd", "<f8", 3), ("maxcrd", "<f8", 3), ("guest", "<i4"), ("pad", "<i4")]), nbx, 200)
    e = psph.PsphErupt()
    (_, _, _, e.bot, e.sml, e.mass, e.rv, e.pos_v[0], e.pos_v[1], e.Vv0, e.ev0, e.rhov, e.pv0, e.gamma_v, e.sndspd) = [float(x) for x in v[:15]]
    e.vel_prof, e.ratio_sml_dx = int(v[15]), float(v[16])
    for k in range(3):
        e.kmin[k], e.kmax[k] = float(v[17 + k]), float(v[20 + k])
    boxes = np.concatenate([rec["mincrd"], rec["maxcrd"]], axis=1)
    return rows, attr, img, e, boxes


def run_plume(args, torch, dist, psph, rank, world, local, dev):
    """-> the "plume" object of the JSON line (rank 0), or None when the compiled reference is not there"""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
    import ref_run                                         # (the reference arm's runner: CPU baseline + set-up, rank 0 only)
    if not ref_run.have_inputs():
        return None
    import tempfile
    top = os.path.join(tempfile.gettempdir(), "psph_plume_case_%s" % os.environ.get("MASTER_PORT", "single"))   # same on every rank
    dump = os.path.join(top, "dump")
    ncpu = os.cpu_count() or 1
    cpu = None
    if rank == 0:
        import shutil
        shutil.rmtree(top, ignore_errors=True)
        os.makedirs(top, exist_ok=True)
        # set-up run (1 step, with the state files) — not timed
        ref_run.time_plume3d(1, 0, replicas=1, edits=ref_run.MEDIUM, keep=os.path.join(top, "setup"), dump=dump)
        if world == 1:
            # CPU arm, bounded: 2 steps of the whole reference loop on every host core (single-rank replicas: no MPI here)
            r = ref_run.time_plume3d(2, 0, replicas=ncpu, edits=ref_run.MEDIUM)
            t = max(r["seconds"])
            cpu = {"value": round(r["updates"] * ncpu / t, 1), "unit": UNIT, "cores": ncpu, "kind": "reference", "steps": 2,
                   "one_replica_updates_per_s": round(r["updates"] / t, 1), "real": r["real"], "total": r["total"],
                   "sample": "unmodified reference, whole loop, %d single-rank replicas side by side, 2 steps" % ncpu}
    if world > 1:
        dist.barrier()
    cfg = psph.config_default_3d()                         # bin/simulation.data = the defaults: Eyjafjallajokull box, h = 100 m
    cfg.device, cfg.rank, cfg.nranks = local, rank, world
    rows, attr, img, er, boxes = load_reference_setup(dump, psph, cfg)
    cfg.smlen_real = er.sml                                # sml_of_phase2 (setup_erupt.cc:95-101)
    s = psph.Solver(cfg)
    s.set_cells(attr)
    part = np.array_split(np.argsort(rows["coords"][:, 0], kind="stable"), world)[rank]
    s.upload(rows[part])
    s.set_images(img)
    s.set_erupt(er, boxes)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(psph.Solver.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        s.comm_init(bytes(idt.cpu().numpy().tobytes()))
        s.repartition()
    stream = torch.cuda.ExternalStream(s.stream(), device=dev)
    warm, steps = 3, max(args.steps, 10)
    t_sim, k, updates, adapt_at = 0.0, 0, 0, None

    def one_step():
        nonlocal t_sim, k
        k += 1
        dt, lost, ierr = s.step(0.0, 1)                    # main.cc:220-408 (dt comes back: TimeProps stays on the host)
        t_sim += dt
        s.add_new_erupt(t_sim, dt, k)                      # :414-417
        adapt = s.scan_outside_layer()                     # :435
        s.apply_bcond()                                    # :468
        if world > 1:
            s.exchange_ghosts()                            # :471
        return adapt, ierr

    for _ in range(warm):
        one_step()
    s.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    bad = 0
    for _ in range(steps):
        updates += s.num_updated
        adapt, ierr = one_step()
        bad += ierr
        if adapt and adapt_at is None:
            adapt_at = k
    ev1.record(stream)
    s.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    u = torch.tensor([float(updates), float(s.n)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    s.close()
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    import shutil
    shutil.rmtree(top, ignore_errors=True)
    ms = float(t.item())
    out = {"workload": "the reference's own 3-D dusty-gas plume (bin/simulation.data + Meteo sounding, initial domain +-1000 m), set up "
                       "by the unmodified reference driver; per step: psph_step + psph_add_new_erupt + psph_scan_outside_layer + "
                       "psph_apply_bcond" + (" + psph_exchange_ghosts" if world > 1 else ""),
           "value": round(float(u[0].item()) / (ms * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms / steps, 4), "steps": steps,
           "warmup": warm, "n_gpus": world, "scaling": "strong", "real_particles": int(float(u[0].item()) / steps),
           "rows": int(u[1].item()) if world == 1 else None, "invariant_violations": bad,
           "domain_growth_requested_at_step": adapt_at,
           "note": "a step is launch-bound at this size (a few thousand real particles): the figure is the like-for-like companion "
                   "of cpu_reference, not a throughput claim; the dt round trip per step is the drop-in's (TimeProps on the host)"}
    if cpu:
        out["cpu_reference"] = cpu
        out["speedup_vs_all_host_cores"] = round(out["value"] / cpu["value"], 1)
    return out



def cpu_baseline_port(n):
    """oracle/psph_oracle.c (scalar C restatement of the reference algorithm, 1 thread) on an
    n^3 sample of the same synthetic cube: one full step."""
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    from plume_sph_b200 import cases
    from common import to_oracle_cfg, particles_to_oracle
    cfg, rows = cases.cube(n, seed=1)
    o = particles_to_oracle(to_oracle_cfg(cfg), rows)
    nreal = int((rows["bc_type"] == 100).sum())
    o.bin()
    t0 = time.perf_counter()
    dt = o.timestep()
    o.search()
    o.mom_engr_update(dt)
    o.smooth_density()
    o.smooth_velocity()
    o.update_pos(dt, 0)
    o.rebin()
    t = time.perf_counter() - t0
    return {"value": round(nreal / t, 1), "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d^3 real-particle sample of the same synthetic cube (%d rows with ghost shell), 1 step, "
                      "oracle/psph_oracle.c, 1 thread, %.1f s" % (n, len(rows), t)}


def run_reference(args):
    """The unmodified reference (oracle/_ref/3d_plume, compiled from /root/reference/src in place)
    on its own 3-D dusty-gas plume case with a small initial domain, 1 rank per replica, as many
    replicas side by side as there are host cores (no MPI in this image)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_build"))
    import ref_run
    ncpu = os.cpu_count() or 1
    if ref_run.have_inputs():
        reps = max(1, min(ncpu, args.ref_replicas if args.ref_replicas > 0 else ncpu))
        # CPU code: no warm-up effect, every step is timed; bounded sample: at most 2 steps (a step is ~7 s on one core)
        steps_run = min(max(1, args.steps), 2)
        r = ref_run.time_plume3d(steps_run, 0, replicas=reps, edits=ref_run.MEDIUM)
        path = ("timestep", "search_neighs_consth", "adaptive_sml", "mom_engr_update", "smooth_density",
                "smooth_velocity", "update_pos", "update_particles_in_bucket")
        t_loop = max(r["seconds"])
        # raw value of the reference's whole loop; the share of it that the B200 path replaces is reported beside it
        t_path = sum(r["stages"].get(k, 0.0) for k in path)
        frac = t_path / max(sum(r["stages"].values()), 1e-30)
        value = r["updates"] * reps / t_loop
        sample = ("unmodified reference, its own 3-D plume case (bin/simulation.data, initial domain shrunk to "
                  "+-1000 m: %d real / %d total particles), %d step(s) timed (steps_requested %d), %d single-rank replicas side "
                  "by side (no MPI in the image); value = whole loop; the stages the B200 path replaces are %.0f%% of it; "
                  "loop wall %.1f s" % (r["real"], r["total"], steps_run, args.steps, reps, 100 * frac, t_loop))
        kind, cores, ms = "reference", reps, t_loop * 1e3 / steps_run
        extra = {"steps_run": steps_run, "hot_path_share_of_loop": round(frac, 4)}
    else:
        cpu = cpu_baseline_port(args.cpu_n)
        value, sample, kind, cores, ms = cpu["value"], cpu["sample"], "port", 1, None
        steps_run, extra = 1, {"steps_run": 1}
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps_run, "warmup": 0, "steps_requested": args.steps, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "3-D dusty-gas plume hot path on host cores (bounded sample)"},
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--edge", dest="n", type=int, default=216, help="cube edge in particles per GPU^(1/3) (216 -> 1.0e7 real)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strong", dest="weak", action="store_false", help="fixed total cube (default: weak, per-GPU cube)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-plume", action="store_true", help="skip the reference's own plume case (the like-for-like companion)")
    ap.add_argument("--cpu-n", type=int, default=32)
    ap.add_argument("--ref-replicas", type=int, default=0)
    ap.add_argument("--sml-int", type=int, default=0, help="run adaptive_sml every k-th step (default: the reference's 50)")
    ap.add_argument("--regime", default="both", choices=["both", "uniform", "mixed"],
                    help="which regimes to time: uniform h (steps 1-49 of a default run), per-particle h (after adaptive_sml), both")
    ap.add_argument("--headline", default="mixed", choices=["uniform", "mixed"],
                    help="regime whose numbers are the line's value / roofline / e2e (the other one is under `regimes`)")
    args = ap.parse_args()
    if args.regime != "both":
        args.headline = args.regime
    try:
        if args.impl == "reference":
            return run_reference(args)
        return run_b200(args)
    except BaseException:
        # a failed rank must not leave its peers waiting in a collective: die hard so the launcher tears the job down
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    main()
