"""Solver — the reference's per-time-step stage API on the GPU.

Method names and order follow the free functions the reference driver calls
(/root/reference/src/plume/particler.h:22-357, call sites src/plume/main.cc:217-486 for 3-D
and :676-751 for 1-D); each forwards to one entry point of include/psph.h.  Error
behaviour mirrors the reference: a negative return (API/CUDA failure) raises, a positive
count of invariant violations is returned like mom_engr_update()'s ``ierr``.
"""
import ctypes as C
import numpy as np
from . import lib as _l

ORDER_DEVICE, ORDER_REFERENCE = 0, 1
KERNELS = ("gather_mom_engr", "gather_density", "gather_velocity", "gather_adaptive_sml", "gather_gradients",
           "gather_gsph_mom_engr", "radix_sort", "reorder", "pair_scan")
STAGES = ("timestep", "search", "adaptive_sml", "gradients", "mom_engr", "density", "velocity", "update_pos", "rebin",
          "apply_bcond", "add_new_erupt", "domain_growth", "halo_exchange")


class Solver:
    def __init__(self, cfg):
        self.L = _l.load()
        self.cfg = cfg.copy()
        h = C.c_void_p()
        rc = self.L.psph_create(C.byref(self.cfg), C.byref(h))
        if rc != 0 or not h:
            raise _l.PsphError(f"psph_create failed ({rc}): "
                               + ("no CUDA device — this library has no CPU fallback" if rc == -3 else "bad config"))
        self.h = h

    # -- life cycle ------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.L.psph_destroy(self.h)
            self.h = None
        for p in getattr(self, "_pinned", []):
            self.L.psph_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc < 0:
            raise _l.PsphError(f"{what} failed ({rc}): {self.L.psph_last_error(self.h).decode()}")
        return rc

    # -- data ------------------------------------------------------------------------
    def set_cells(self, attr):
        if attr is None:
            return self._ck(self.L.psph_set_cells(self.h, None, 0), "psph_set_cells")
        attr = np.ascontiguousarray(attr, dtype=_l.CELL_ATTR_DTYPE)
        return self._ck(self.L.psph_set_cells(self.h, attr.ctypes.data, attr.size), "psph_set_cells")

    def get_cells(self):
        """the bucket attributes as the device holds them (the domain-growth calls change has_involved, active, flags)"""
        attr = np.zeros(int(self.cfg.ncell[0]) * int(self.cfg.ncell[1]) * int(self.cfg.ncell[2]), _l.CELL_ATTR_DTYPE)
        self._ck(self.L.psph_get_cells(self.h, attr.ctypes.data, attr.size), "psph_get_cells")
        return attr

    # ---- domain growth on the device ----
    def scan_outside_layer(self):
        """scan_outside_layer (scan_outside_layer.cc:44-121) -> adapt"""
        a = C.c_int(0)
        self._ck(self.L.psph_scan_outside_layer(self.h, C.byref(a)), "psph_scan_outside_layer")
        return a.value

    def adapt_domain(self):
        """adapt_domain (adapt_domain.cc:41-194) -> rows turned real on this rank"""
        a = C.c_int(0)
        self._ck(self.L.psph_adapt_domain(self.h, C.byref(a)), "psph_adapt_domain")
        return a.value

    def set_growth(self, growth, meteo_z=None, meteo_rho=None, meteo_p=None, cell_mincrd=None):
        """growth record (lib.PsphGrowth), the sounding's columns (height [km], density, pressure [hPa]) and, optionally,
        Bucket::mincrd of every bucket as the driver stores it"""
        g = _l.PsphGrowth()
        C.memmove(C.byref(g), C.byref(growth), C.sizeof(g))
        keep = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (meteo_z, meteo_rho, meteo_p, cell_mincrd)]
        ptr = [None if a is None else a.ctypes.data for a in keep]
        return self._ck(self.L.psph_set_growth(self.h, C.byref(g), *ptr), "psph_set_growth")

    def add_pressure_ghost(self, add_step):
        """add_pressure_ghost (add_pressure_ghost.cc:36-316) on the device -> rows appended on this rank"""
        a = C.c_int(0)
        self._ck(self.L.psph_add_pressure_ghost(self.h, int(add_step), C.byref(a)), "psph_add_pressure_ghost")
        return a.value

    def add_wall_ghost(self, add_step):
        """add_wall_ghost (add_wall_ghost.cc:44-293) on the device -> rows appended on this rank"""
        a = C.c_int(0)
        self._ck(self.L.psph_add_wall_ghost(self.h, int(add_step), C.byref(a)), "psph_add_wall_ghost")
        return a.value

    def host_rows(self, n, dtype=None):
        """n rows (psph_particle, or lib.PARTICLE_SPH_DTYPE) in pinned host memory (psph_host_alloc); freed with the solver"""
        dtype = dtype or _l.PARTICLE_DTYPE
        nbytes = max(int(n), 1) * dtype.itemsize
        p = self.L.psph_host_alloc(nbytes)
        if not p:
            raise _l.PsphError("psph_host_alloc failed")
        self._pinned = getattr(self, "_pinned", []) + [p]
        buf = (C.c_char * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=dtype, count=int(n))

    @staticmethod
    def _rows(rows):
        """contiguous rows in one of the two row layouts (PARTICLE_SPH_DTYPE arrays stay as they are)"""
        if getattr(rows, "dtype", None) == _l.PARTICLE_SPH_DTYPE:
            return np.ascontiguousarray(rows)
        return np.ascontiguousarray(rows, dtype=_l.PARTICLE_DTYPE)

    def upload(self, rows):
        rows = self._rows(rows)
        return self._ck(self.L.psph_upload_rows(self.h, rows.ctypes.data, rows.size, rows.dtype.itemsize), "psph_upload_rows")

    def append(self, rows):
        rows = self._rows(rows)
        return self._ck(self.L.psph_append_rows(self.h, rows.ctypes.data, rows.size, rows.dtype.itemsize), "psph_append_rows")

    def download(self, order=ORDER_REFERENCE, out=None, dtype=None):
        """rows as psph_particle (default) or, with dtype=lib.PARTICLE_SPH_DTYPE / such an `out`, without gradients"""
        n = C.c_size_t(0)
        self._ck(self.L.psph_download(self.h, None, 0, C.byref(n), order), "psph_download")
        if out is None:
            out = np.zeros(n.value, dtype or _l.PARTICLE_DTYPE)
        self._ck(self.L.psph_download_rows(self.h, out.ctypes.data, out.size, C.byref(n), order, out.dtype.itemsize),
                 "psph_download_rows")
        return out[:n.value]

    H5PART_DOUBLE = ("x", "y", "z", "Rho", "Vx", "Vy", "Vz", "engr", "mssfrc", "prss")
    H5PART_INT = ("phase", "bctp", "guest")

    def download_h5part(self, with_ghosts=False):
        """the datasets write_h5part (h5part_writer.cc:40-269) puts into one "Step#k" group, as a dict of arrays in the
        writer's order (THashTable iteration, non-guest real particles unless with_ghosts): ID, x .. prss, phase, bctp,
        guest — selected, ordered and gathered on the device (psph_download_h5part)"""
        n = C.c_size_t(0)
        self._ck(self.L.psph_download_h5part(self.h, None, None, 0, C.byref(n), int(with_ghosts)), "psph_download_h5part")
        m = n.value
        d = {k: np.zeros(m, np.float64) for k in self.H5PART_DOUBLE}
        i = {k: np.zeros(m, np.int32) for k in self.H5PART_INT}
        dp = (C.c_void_p * 10)(*[d[k].ctypes.data for k in self.H5PART_DOUBLE])
        ip = (C.c_void_p * 3)(*[i[k].ctypes.data for k in self.H5PART_INT])
        self._ck(self.L.psph_download_h5part(self.h, dp, ip, m, C.byref(n), int(with_ghosts)), "psph_download_h5part")
        out = {"ID": np.arange(m, dtype=np.int32)}
        out.update(d)
        out.update(i)
        return out

    def upload_device_soa(self, cols, bc=None, phase=None, n=0):
        """cols: list of PSPH_COL_COUNT device pointers (ints); bc/phase: device pointers or None."""
        arr = (C.c_void_p * len(cols))(*cols)
        return self._ck(self.L.psph_upload_device_soa(self.h, arr, bc, phase, n), "psph_upload_device_soa")

    @property
    def n(self):
        return int(self.L.psph_num_particles(self.h))

    def neighbours(self):
        """(offsets, idx, rows): CSR neighbour lists over device rows + the rows in that order."""
        n = self.n
        nnz = C.c_size_t(0)
        off = np.zeros(n + 1, np.int64)
        self._ck(self.L.psph_emit_neighbours(self.h, off.ctypes.data, None, 0, C.byref(nnz)), "psph_emit_neighbours")
        n = self.n
        off = off[:n + 1]
        idx = np.zeros(max(nnz.value, 1), np.int32)
        self._ck(self.L.psph_emit_neighbours(self.h, off.ctypes.data, idx.ctypes.data, idx.size, C.byref(nnz)),
                 "psph_emit_neighbours")
        rows = self.download(ORDER_DEVICE)
        return off, idx[:nnz.value], rows

    def cells(self):
        """(cells, rows): background-cell index of every row + the rows, both in device order"""
        n = C.c_size_t(0)
        self._ck(self.L.psph_download(self.h, None, 0, C.byref(n), ORDER_DEVICE), "psph_download")
        cells = np.zeros(max(n.value, 1), np.int32)
        self._ck(self.L.psph_cell_of_rows(self.h, cells.ctypes.data, cells.size), "psph_cell_of_rows")
        return cells[:n.value], self.download(ORDER_DEVICE)

    # -- stages (reference names) ------------------------------------------------------
    def timestep(self):
        dt = C.c_double(0)
        self._ck(self.L.psph_timestep(self.h, C.byref(dt)), "psph_timestep")
        return dt.value

    def set_dt(self, dt, step=0, erupting=1):
        return self._ck(self.L.psph_set_dt(self.h, dt, step, int(erupting)), "psph_set_dt")

    def search_neighs_consth(self):
        return self._ck(self.L.psph_search_neighs(self.h), "psph_search_neighs")

    def adaptive_sml(self):
        return self._ck(self.L.psph_adaptive_sml(self.h), "psph_adaptive_sml")

    def calc_gradients(self):
        return self._ck(self.L.psph_calc_gradients(self.h), "psph_calc_gradients")

    def mom_engr_update(self):
        return self._ck(self.L.psph_mom_engr_update(self.h), "psph_mom_engr_update")

    def smooth_density(self):
        return self._ck(self.L.psph_smooth_density(self.h), "psph_smooth_density")

    def smooth_velocity(self):
        return self._ck(self.L.psph_smooth_velocity(self.h), "psph_smooth_velocity")

    def update_pos(self):
        return self._ck(self.L.psph_update_pos(self.h), "psph_update_pos")

    def set_erupt(self, erupt, boxes):
        """vent record (lib.PsphErupt) + [n][6] mincrd, maxcrd of the buckets flagged is_erupt()"""
        b = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
        e = _l.PsphErupt()
        C.memmove(C.byref(e), C.byref(erupt), C.sizeof(e))
        return self._ck(self.L.psph_set_erupt(self.h, C.byref(e), b.ctypes.data, len(b)), "psph_set_erupt")

    def add_new_erupt(self, time, dt, add_step):
        """add_new_erupt (setup_erupt.cc:390-670) on the device; returns the number of particles added"""
        added = C.c_int(0)
        self._ck(self.L.psph_add_new_erupt(self.h, float(time), float(dt), int(add_step), C.byref(added)), "psph_add_new_erupt")
        return added.value

    def set_images(self, images):
        """wall-ghost images (IMAGE_DTYPE rows: ghost key, bucket cell, image point) — the driver's Image_table"""
        img = np.ascontiguousarray(images, dtype=_l.IMAGE_DTYPE)
        return self._ck(self.L.psph_set_images(self.h, img.ctypes.data, len(img)), "psph_set_images")

    def apply_bcond(self):
        return self._ck(self.L.psph_apply_bcond(self.h), "psph_apply_bcond")

    def update_particles_in_bucket(self):
        lost = C.c_int(0)
        self._ck(self.L.psph_rebin(self.h, C.byref(lost)), "psph_rebin")
        return lost.value

    def errors(self, reset=True):
        return self._ck(self.L.psph_errors(self.h, int(reset)), "psph_errors")

    # -- fused drivers -----------------------------------------------------------------
    def step(self, dt_max=0.0, erupting=1):
        """One full time step in the reference's order; returns (dt_used, lost, ierr)."""
        dt, lost = C.c_double(0), C.c_int(0)
        rc = self._ck(self.L.psph_step(self.h, dt_max, int(erupting), C.byref(dt), C.byref(lost)), "psph_step")
        return dt.value, lost.value, rc

    def run(self, nsteps, erupting=1):
        return self._ck(self.L.psph_run(self.h, nsteps, int(erupting)), "psph_run")

    def sync(self):
        return self._ck(self.L.psph_sync(self.h), "psph_sync")

    # -- instrumentation -----------------------------------------------------------------
    def launch_count(self, reset=True):
        return int(self.L.psph_launch_count(self.h, int(reset)))

    def set_timing(self, on=True):
        self.L.psph_set_timing(self.h, int(on))

    def stage_ms(self, reset=False):
        return {name: float(self.L.psph_stage_ms(self.h, i, int(reset))) for i, name in enumerate(STAGES)}

    def kernel_ms(self, reset=False):
        """{kernel: (device ms, launches)} of the individually timed kernels (timing on only)"""
        out = {}
        for i, name in enumerate(KERNELS):
            c = C.c_int64(0)
            ms = float(self.L.psph_kernel_ms(self.h, i, int(reset), C.byref(c)))
            out[name] = (ms, int(c.value))
        return out

    def stream(self):
        return self.L.psph_stream(self.h)

    def column(self, col):
        p = C.c_void_p()
        self._ck(self.L.psph_column(self.h, col, C.byref(p)), "psph_column")
        return p.value

    # -- multi-GPU ---------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        """rank 0: the NCCL unique id (bytes) every rank passes to comm_init"""
        buf = (C.c_char * 128)()
        n = _l.load().psph_comm_unique_id(buf, 128)
        if n <= 0:
            raise _l.PsphError("psph_comm_unique_id failed (NCCL not available)")
        return bytes(buf[:n])

    def comm_init(self, unique_id):
        buf = (C.c_char * len(unique_id)).from_buffer_copy(unique_id)
        return self._ck(self.L.psph_comm_init(self.h, buf, len(unique_id)), "psph_comm_init")

    def cell_ranks(self):
        """rank along the bucket Hilbert curve of every background cell (linear index order)"""
        r = np.zeros(self.cfg.ncells, np.int32)
        self._ck(self.L.psph_cell_ranks(self.h, r.ctypes.data, r.size), "psph_cell_ranks")
        return r

    def slab_cuts(self):
        cuts = (C.c_int * 64)()
        p = self._ck(self.L.psph_slab_cuts(self.h, cuts, 64), "psph_slab_cuts")
        return [int(cuts[k]) for k in range(p + 1)]

    def comm_bytes_sent(self, reset=True):
        return int(self.L.psph_comm_bytes_sent(self.h, int(reset)))

    @property
    def num_updated(self):
        return int(self._ck(self.L.psph_num_updated(self.h), "psph_num_updated"))

    def pair_stats(self):
        """pair path of the last step: dict(pair_mode, rows_updated, rows_without_list, near_overflow_rows)"""
        o = (C.c_int32 * 4)()
        self._ck(self.L.psph_pair_stats(self.h, o), "psph_pair_stats")
        return dict(pair_mode=int(o[0]), rows_updated=int(o[1]), rows_without_list=int(o[2]), near_overflow_rows=int(o[3]))

    def exchange_ghosts(self):
        return self._ck(self.L.psph_exchange_ghosts(self.h), "psph_exchange_ghosts")

    def repartition(self):
        return self._ck(self.L.psph_repartition(self.h), "psph_repartition")
