"""Host side of the time loop, mirroring the reference driver (src/plume/main.cc).

``TimeProps`` restates the reference's time bookkeeping (src/datstr/properties.h:20-150): it
stays on the host in the drop-in design, exactly where the reference keeps it; the values come
from simulation.data through ``read_simulation_data`` (src/plume/dataread.cc:109-235, every
number parsed through C ``stof``, i.e. rounded to float first).  ``run_loop`` is the 1-D /
3-D loop of main.cc:217-486 / :676-751 with each stage call forwarded to the GPU.
"""
import numpy as np


def stof(s):
    """dataread.cc parses every field with std::stof: the value is rounded to float32 first."""
    return float(np.float32(float(s)))


class TimeProps:
    """src/datstr/properties.h:20-150"""

    def __init__(self, max_time=180.0, max_steps=10, timeoutput=1.0, stat_erupt=0.0, end_erupt=7200.0, time_scale=1.0):
        self.time = 0.0
        self.dtime = 0.0
        self.max_time = max_time
        self.TIME_SCALE = time_scale
        self.ndmax_time = max_time / time_scale      # dataread.cc:145
        self.timeoutput = timeoutput
        self.ndtimeoutput = timeoutput / time_scale  # dataread.cc:147
        self.stat_erupt = stat_erupt
        self.end_erupt = end_erupt
        self.step = 0
        self.ioutput = 0
        self.max_steps = max_steps

    def incrtime(self, dt):
        """properties.h:71-82: clip dt to hit output/end time exactly, advance; returns the dt used"""
        if self.time + dt > self.ndtimeoutput:
            dt = self.ndtimeoutput - self.time
        if self.time + dt > self.ndmax_time:
            dt = self.ndmax_time - self.time
        self.time += dt
        self.dtime = dt
        self.step += 1
        return dt

    def ifend(self):
        return self.time >= self.ndmax_time or self.step > self.max_steps

    def ifoutput(self):
        if self.time >= self.ndtimeoutput or self.time == self.ndmax_time:
            self.ioutput += 1
            self.ndtimeoutput = ((self.ioutput + 1) * self.timeoutput) / self.TIME_SCALE
            return True
        return False

    def iferupt(self):
        t = self.time * self.TIME_SCALE
        return not (t < self.stat_erupt or t > self.end_erupt)


def read_simulation_data(path):
    """simulation.data: 18 `label:value` lines read positionally (dataread.cc:109-235)."""
    vals = [ln.split(":", 1)[1].strip() for ln in open(path) if ":" in ln]
    d = dict(max_time=stof(vals[0]), max_steps=int(stof(vals[1])), timeoutput=stof(vals[2]), stat_erupt=stof(vals[3]),
             end_erupt=stof(vals[4]), format=int(stof(vals[5])), P_CONSTANT=stof(vals[6]), GAMMA=stof(vals[7]),
             smoothing_length=stof(vals[8]), Lx=(stof(vals[9]), stof(vals[10])), Ly=(stof(vals[11]), stof(vals[12])),
             Lz=(stof(vals[13]), stof(vals[14])))
    return d


def timeprops_from(d, time_scale=1.0):
    return TimeProps(d["max_time"], d["max_steps"], d["timeoutput"], d["stat_erupt"], d["end_erupt"], time_scale)


def grow_domain(solver, tp, growth):
    """The adapt block of main.cc:256-304 for a flat ground: the particle work on the device (adapt_domain, add_pressure_ghost,
    add_wall_ghost, apply_bcond), the bucket bookkeeping on the host on the attribute array (shift_brief_buck, update_bgmesh,
    the image search).  growth: dict(latent=attr array of mesh.background_mesh with every cell a bucket, mindom, ncell, delta,
    len_scale); the Solver must have been given psph_set_growth."""
    from . import mesh
    n = [int(x) for x in growth["ncell"]]
    solver.adapt_domain()                                             # :265
    solver.add_pressure_ghost(tp.step)                                # :271
    attr = solver.get_cells()
    mesh.shift_brief_buck(attr, growth["latent"], n)                  # :274
    solver.set_cells(attr)
    solver.add_wall_ghost(tp.step)                                    # :277
    attr = solver.get_cells()
    cells, rows = solver.cells()                                      # bucket of every row + the rows (a growth step is rare:
    cnt = np.bincount(np.asarray(cells)[np.asarray(cells) >= 0], minlength=len(attr))   # one read-back for the host's bookkeeping)
    mesh.update_bgmesh(attr, cnt, n)                                  # :286
    solver.set_cells(attr)
    img, _ = mesh.wall_ghost_images(rows, growth["mindom"], n, growth["delta"], attr, growth.get("len_scale", 1.0))   # :292
    solver.set_images(img)
    return solver.apply_bcond()                                       # :298


def run_loop(solver, tp, on_output=None, max_steps=None, erupt=False, growth=None):
    """The reference time loop with the stage calls on the GPU.  One host round trip per step
    (the CFL dt comes back, is clipped by TimeProps.incrtime and goes down again) exactly as
    main.cc:220-231 does; ``Solver.run`` is the variant without it.  growth (see grow_domain): the domain follows the
    plume as in an ADJUST_DOMAIN build — psph_scan_outside_layer every step (:435), the adapt block at the start of the
    step after it raised `adapt` (:256-304)."""
    cfg = solver.cfg
    nsteps = 0
    adapt = 0
    while not tp.ifend() and (max_steps is None or nsteps < max_steps):
        dt = solver.timestep()                       # main.cc:220 / :679
        dt = tp.incrtime(dt)                         # :231 / :682
        if growth is not None and adapt:
            grow_domain(solver, tp, growth)          # :256-304
        adapt = 0
        solver.set_dt(dt, tp.step, tp.iferupt())
        solver.search_neighs_consth()                # :318 (3-D; a geometry refresh in 1-D)
        if cfg.adaptive_sml and (cfg.dim == 1 or tp.step % cfg.sml_update_int == 0):
            solver.adaptive_sml()                    # :322 / :690-694
        if cfg.use_gsph:
            solver.calc_gradients()                  # :343 / :699
        solver.mom_engr_update()                     # :358 / :705
        solver.smooth_density()                      # :378 / :722
        if cfg.have_turbulence_lans:
            solver.smooth_velocity()                 # :392 / :726
        solver.update_pos()                          # :405 / :730
        if cfg.dim == 3:
            solver.update_particles_in_bucket()      # :408
            if erupt and tp.iferupt():
                solver.add_new_erupt(tp.time, dt, tp.step)   # :414-417 (after Solver.set_erupt)
            if growth is not None:
                adapt = solver.scan_outside_layer()  # :435
            solver.apply_bcond()                     # :468 (no-op until the driver hands over its Image_table)
        nsteps += 1
        if tp.ifoutput() and on_output is not None:  # :474 / :746
            on_output(solver, tp)
    return nsteps
