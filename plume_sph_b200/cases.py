"""Synthetic inputs of the BASELINE shapes.

``cube(n)``: the weak-scaling particle cube — an n^3 lattice of real air particles (spacing =
smoothing length h = 1, eight per bucket edge as the reference's add_air lattice,
src/plume/add_air.cc:592-598), positions jittered by U(-0.1h, 0.1h), velocities U(-0.05c, 0.05c)
per component, a central sphere of radius n/8 of erupted material (phase 2, mass fraction 1, in
pressure equilibrium), wrapped in one bucket-thick (8h) shell of static pressure ghosts
(bc_type 1, add_pressure_ghost.cc:95-215).  3-D SPH defaults with ATMOSPHERE_TYPE 2 (g = 0).
Secondary variables come from the reference's equation of state (particle.cc:528-604).
"""
import numpy as np
from . import lib as _l

RHO0, P0 = 1.104, 1.0125e5


def cube_config(n, smlen=1.0, origin=(0.0, 0.0, 0.0)):
    cfg = _l.config_default_3d()
    cfg.g = 0.0                       # ATMOSPHERE_TYPE == 2 (parameters.h:894-898)
    d = 8.0 * smlen
    for k in range(3):
        cfg.mindom[k] = origin[k] - 2 * d          # ghost shell + one spare bucket layer
        cfg.ncell[k] = int(np.ceil(n * smlen / d)) + 4
    cfg.delta = d
    cfg.len_scale = 1.0
    cfg.smlen_real = smlen
    cfg.erupt_top = -1.0e300
    cfg.rhoa0, cfg.pa0 = RHO0, P0
    return cfg


def eos(cfg, rho, ks, e):
    """Particle::update_second_var (particle.cc:528-604), ideal-gas mixture branch, vectorised"""
    na = 1 - ks
    ng = cfg.ng0 * ks
    ns = ks - ng
    Cvm = ns * cfg.Cvs + ng * cfg.Cvg + na * cfg.Cva
    Rm = ng * cfg.Rg + na * cfg.Ra
    gamma = 1 + Rm / Cvm
    desmg = rho * (na + ng)
    p = (Rm / (ng + na)) * desmg * e / Cvm
    c = np.sqrt(gamma * p / desmg)
    return dict(pressure=p, gamma=gamma, sound_speed=c, temperature=e / Cvm, specific_heat_p=Cvm,
                rph0=rho * na, rph1=rho * ks)


def cube(n, seed=20260101, smlen=1.0, ghosts=True, jitter=0.1, vjit=0.05):
    """-> (cfg, rows) host rows (plume_sph_b200.PARTICLE_DTYPE) of the synthetic cube."""
    cfg = cube_config(n, smlen)
    g = 8 if ghosts else 0
    m = n + 2 * g
    idx = np.arange(m) - g
    I, J, K = np.meshgrid(idx, idx, idx, indexing="ij")
    I, J, K = I.ravel(), J.ravel(), K.ravel()
    N = I.size
    real = (I >= 0) & (I < n) & (J >= 0) & (J < n) & (K >= 0) & (K < n)
    rng = np.random.default_rng(seed)
    rows = np.zeros(N, _l.PARTICLE_DTYPE)
    pos = (np.stack([I, J, K], 1) + 0.5) * smlen
    pos[real] += rng.uniform(-jitter * smlen, jitter * smlen, (int(real.sum()), 3))
    rows["coords"] = pos
    ctr = 0.5 * n * smlen
    r2 = ((pos - ctr) ** 2).sum(1)
    ph2 = real & (r2 < (n * smlen / 8.0) ** 2)
    ks = np.where(ph2, 1.0, 0.0)
    rho = np.full(N, RHO0)
    # air: e = p0 / (0.4 rho0); erupted material: same pressure (p = Rg*rho*ng*e/Cvm)
    e = np.full(N, P0 / (0.4 * RHO0))
    ng = cfg.ng0
    cvm2 = (1 - ng) * cfg.Cvs + ng * cfg.Cvg
    p_air = cfg.Ra * RHO0 * e[0] / cfg.Cva
    e[ph2] = p_air * cvm2 / (cfg.Rg * ng * RHO0)
    sec = eos(cfg, rho, ks, e)
    c_air = float(sec["sound_speed"][~ph2][0])
    vel = np.zeros((N, 3))
    vel[real] = rng.uniform(-vjit * c_air, vjit * c_air, (int(real.sum()), 3))
    rows["vel"] = vel
    rows["smoothed_v"] = vel
    rows["rho"] = rho
    rows["engr"] = e
    rows["smoothed_e"] = e
    rows["mass"] = RHO0 * smlen ** 3
    rows["smlen"] = smlen
    rows["mass_frac"] = ks
    for f in ("pressure", "gamma", "sound_speed", "temperature", "specific_heat_p"):
        rows[f] = sec[f]
    rows["phase_density"][:, 0] = sec["rph0"]
    rows["phase_density"][:, 1] = sec["rph1"]
    rows["bc_type"] = np.where(real, 100, 1)
    rows["phase_num"] = np.where(ph2, 2, 1)
    rows["involved"] = np.where(ph2, 2, np.where(real, 1, 0))
    ids = np.arange(N, dtype=np.uint64)
    rows["key"][:, 0] = (ids >> np.uint64(32)).astype(np.uint32)
    rows["key"][:, 1] = (ids & np.uint64(0xffffffff)).astype(np.uint32)
    return cfg, rows
