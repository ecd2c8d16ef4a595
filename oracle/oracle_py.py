"""ctypes binding of the CPU oracle (oracle/psph_oracle.c) + helpers that turn reference
dumps (oracle/refio.py) into oracle states.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpsph_oracle.so")


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("psph_oracle.c", "psph_oracle_gsph.c")]
    srcs = [s for s in srcs if os.path.exists(s)]
    deps = srcs + [os.path.join(HERE, "psph_oracle.h"), os.path.join(HERE, "..", "include", "psph.h")]
    if (not force and os.path.exists(LIB)
            and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c99", "-o", LIB] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    return LIB


class Config(C.Structure):
    """Mirror of psph_config (include/psph.h)."""
    _fields_ = [
        ("abi_version", C.c_int), ("dim", C.c_int),
        ("use_gsph", C.c_int), ("riemann_solver", C.c_int), ("have_turbulence_lans", C.c_int),
        ("have_energy_smooth", C.c_int), ("pghost_contribute_des", C.c_int),
        ("fluid_compressibility", C.c_int), ("adaptive_sml", C.c_int), ("sml_update_int", C.c_int),
        ("cutoff", C.c_double), ("cutoff2", C.c_double), ("neigh_radius", C.c_double),
        ("msfrc_thresh", C.c_double), ("epsilon_half", C.c_double), ("prandtl_num", C.c_double),
        ("heat_trans_scale_ratio", C.c_double), ("e_smooth_ratio", C.c_double), ("tiny", C.c_double),
        ("alf", C.c_double), ("beta", C.c_double), ("ata", C.c_double), ("g", C.c_double),
        ("cfl", C.c_double), ("eta_smooth", C.c_double), ("c_smooth", C.c_double), ("thresh", C.c_double),
        ("num_loop", C.c_int), ("gsph_reconstruct", C.c_int),
        ("ng0", C.c_double), ("Cvs", C.c_double), ("Cvg", C.c_double), ("Cva", C.c_double),
        ("Rg", C.c_double), ("Ra", C.c_double), ("rhoa0", C.c_double), ("pa0", C.c_double),
        ("erupt_top", C.c_double), ("smlen_real", C.c_double),
        ("mindom", C.c_double * 3), ("delta", C.c_double), ("len_scale", C.c_double),
        ("ncell", C.c_int * 3),
        ("ptable_size", C.c_int),
        ("device", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("pad1", C.c_int),
    ]


class CellAttr(C.Structure):
    _fields_ = [("type", C.c_int32), ("has_involved", C.c_int32), ("active", C.c_int32),
                ("bucket_index", C.c_int32 * 6), ("flags", C.c_int32), ("bnd", C.c_double * 6)]


class Growth(C.Structure):
    """Mirror of ``psph_growth`` (include/psph.h): what add_pressure_ghost reads (add_pressure_ghost.cc:36-316)."""
    _fields_ = [("mass", C.c_double), ("smlen", C.c_double), ("ratio_sml_dx", C.c_double), ("gamma_air", C.c_double),
                ("sndspd", C.c_double), ("gmm", C.c_double), ("Ta0", C.c_double), ("kmin", C.c_double * 3), ("kmax", C.c_double * 3),
                ("Ll", C.c_double * 3), ("Lu", C.c_double * 3), ("atmosphere_type", C.c_int32), ("nmeteo", C.c_int32)]



class Erupt(C.Structure):
    """Mirror of ``psph_erupt`` (include/psph.h): what add_new_erupt reads (setup_erupt.cc:390-670)."""
    _fields_ = [("pos_v", C.c_double * 2), ("rv", C.c_double), ("Vv0", C.c_double), ("vel_prof", C.c_int32),
                ("pad", C.c_int32), ("bot", C.c_double), ("sml", C.c_double), ("mass", C.c_double),
                ("ratio_sml_dx", C.c_double), ("ev0", C.c_double), ("rhov", C.c_double), ("pv0", C.c_double),
                ("gamma_v", C.c_double), ("sndspd", C.c_double), ("kmin", C.c_double * 3), ("kmax", C.c_double * 3)]


# host rows of the boundary (psph_particle, include/psph.h): what psph_oracle_add_new_erupt writes
PARTICLE_DTYPE = np.dtype([("key", "<u4", 3), ("bc_type", "<i4"), ("phase_num", "<i4"), ("involved", "<i4"), ("guest", "<i4"),
                           ("myprocess", "<i4"), ("mass", "<f8"), ("smlen", "<f8"), ("mass_frac", "<f8"), ("pressure", "<f8"),
                           ("gamma", "<f8"), ("sound_speed", "<f8"), ("temperature", "<f8"), ("specific_heat_p", "<f8"),
                           ("phase_density", "<f8", 2), ("coords", "<f8", 3), ("vel", "<f8", 3), ("smoothed_v", "<f8", 3),
                           ("rho", "<f8"), ("engr", "<f8"), ("smoothed_e", "<f8"), ("grad", "<f8", 15)])
assert PARTICLE_DTYPE.itemsize == 328

CELL_ATTR_DTYPE = np.dtype([("type", "<i4"), ("has_involved", "<i4"), ("active", "<i4"),
                            ("bucket_index", "<i4", 6), ("flags", "<i4"), ("bnd", "<f8", 6)])
assert CELL_ATTR_DTYPE.itemsize == C.sizeof(CellAttr)


def config_3d(smlen=100.0, Ll=(-20800.0, -20800.0, 1666.0), Lu=(20800.0, 20800.0, 24000.0), g=9.80665):
    """Reference 3-D defaults: options.h:452-837, constant.h:99-203, parameters.h:881-1046;
    mesh as src/preproc/preprocess.cc:99-133."""
    c = Config()
    c.abi_version = 1
    c.dim = 3
    c.use_gsph = 0
    c.riemann_solver = 0
    c.gsph_reconstruct = 3
    c.have_turbulence_lans = 1
    c.have_energy_smooth = 1
    c.pghost_contribute_des = 1
    c.fluid_compressibility = 0
    c.adaptive_sml = 1
    c.sml_update_int = 50
    c.cutoff, c.cutoff2, c.neigh_radius = 3.0, 8.0, 4.25
    c.msfrc_thresh = 1.0e-6
    c.epsilon_half, c.prandtl_num = 0.4, 0.85
    c.heat_trans_scale_ratio, c.e_smooth_ratio, c.tiny = 1.0, 4.0, 1.0e-8
    c.alf, c.beta, c.ata = 0.3, 0.6, 0.01
    c.g = g
    c.cfl, c.eta_smooth, c.c_smooth, c.thresh, c.num_loop = 0.8, 1.5, 2.0, 1e-5, 2
    c.ng0, c.Cvs, c.Cvg, c.Cva, c.Rg, c.Ra = 0.026, 1100.0, 1348.0, 717.0, 462.0, 287.0
    c.rhoa0, c.pa0 = 1.10894, 0.831e5
    c.erupt_top = Ll[2]
    c.smlen_real = smlen
    d = 8 * smlen
    n = []
    for k in range(3):
        lo = Ll[k] - 1.5 * d
        hi = Lu[k] + 1.5 * d
        c.mindom[k] = lo
        n.append(int(np.ceil((hi - lo) / d)))
    c.delta = d
    c.len_scale = 1.0
    for k in range(3):
        c.ncell[k] = n[k]
    c.ptable_size = 4000000
    c.device, c.rank, c.nranks = 0, 0, 1
    return c


def config_jpue():
    """BASELINE config 2, JPUE weakly-compressible jet: 3-D options with FLUID_COMPRESSIBILITY 1 (Tait EOS,
    particle.cc:528-604) and ATMOSPHERE_TYPE 2 (g = 0), parameters.h block :404-562,
    bin/simulation_JPUE_incomp_back.data (smoothing length 8)."""
    c = config_3d(smlen=8.0, Ll=(-576.0, -576.0, 0.0), Lu=(576.0, 576.0, 6000.0), g=0.0)
    c.fluid_compressibility = 1
    c.cfl, c.eta_smooth = 0.45, 1.2
    c.ng0, c.Cvs, c.Cvg, c.Cva, c.Rg, c.Ra = 1.0, 1617.0, 4197.0, 4197.0, 25182.0, 25182.0
    c.rhoa0, c.pa0 = 1000.0, 1.013e5
    return c


def config_3d_gsph():
    """the 3-D defaults with USE_GSPH 1 (options.h:544-663): Roe solver, Riemann states from the gradients
    (GSPH_SPECIFIC_VOL_APP 3, MINI_DERIVATIVE_COND 10, GSPH_MODIFIED_MONOTONICITY 1)"""
    c = config_3d()
    c.use_gsph, c.riemann_solver, c.gsph_reconstruct = 1, 0, 3
    return c


CONFIGS_3D = {"plume3d_small": config_3d, "jpue3d": config_jpue, "plume3d_sml": config_3d, "plume3d_grow": config_3d,
              "plume3d_small_gsph": config_3d_gsph, "plume3d_gsph2": config_3d_gsph, "grow3d": config_3d}


def config_for(name, z=None):
    """configuration of a 3-D golden fixture.  The smoothing length an erupt ghost takes when it turns real
    (simprops->sml_of_phase2, computed by the driver in setup_erupt.cc:95-101 and handed to the hot path as
    psph_config.smlen_real) is read off the fixture: the rows that turned real in step 1."""
    c = CONFIGS_3D[name]()
    if z is not None and "pre" in z:
        turned = (z["pre"]["bc_type"] == 0) & (z["s1_pos_bc_type"] == 100)
        if turned.any():
            c.smlen_real = float(z["s1_pos_smlen"][turned][0])
    elif z is not None:   # sparse stage outputs (plume3d_sml): the rows update_pos turned real
        for st in (1, 2):
            rows, bc = z["s%d_pos_bc_type__rows" % st], z["s%d_pos_bc_type" % st]
            hr = dict(zip(z["s%d_pos_smlen__rows" % st].tolist(), z["s%d_pos_smlen" % st].tolist()))
            for r, b in zip(rows.tolist(), bc.tolist()):
                if b == 100 and r in hr:
                    c.smlen_real = float(hr[r])
    return c


def config_1d(gsph=False):
    """Reference 1-D defaults: options.h:18-409, constant.h:14-98, parameters.h:1228-1450."""
    c = Config()
    c.abi_version = 1
    c.dim = 1
    c.use_gsph = 1 if gsph else 0
    c.riemann_solver = 1
    c.have_turbulence_lans = 0
    c.have_energy_smooth = 0
    c.pghost_contribute_des = 2
    c.fluid_compressibility = 0
    c.adaptive_sml = 1
    c.sml_update_int = 1
    c.cutoff, c.cutoff2, c.neigh_radius = 3.0, 12.0, 4.25
    c.msfrc_thresh = 1.0e-6
    c.epsilon_half, c.prandtl_num = 0.4, 0.85
    c.heat_trans_scale_ratio, c.e_smooth_ratio, c.tiny = 1.0, 4.0, 1.0e-8
    c.alf, c.beta, c.ata = 1.0, 2.0, 0.01
    c.g = 0.0
    c.cfl, c.eta_smooth, c.c_smooth, c.thresh, c.num_loop = 0.1, 1.0, 2.0, 1e-5, 5
    c.ng0, c.Cvs, c.Cvg, c.Cva, c.Rg, c.Ra = 0.05, 1100.0, 1348.0, 717.0, 462.0, 287.0
    c.rhoa0, c.pa0 = 1.104, 1.0125e5
    c.erupt_top = 0.0
    c.smlen_real = 0.0
    for k in range(3):
        c.mindom[k] = 0.0
        c.ncell[k] = 1
    c.delta = 1.0
    c.len_scale = 1.0
    c.ptable_size = 10000
    c.device, c.rank, c.nranks = 0, 0, 1
    return c


_P = C.POINTER


class State(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("key", _P(C.c_uint32)),
        ("bc", _P(C.c_int32)), ("phase", _P(C.c_int32)), ("involved", _P(C.c_int32)), ("guest", _P(C.c_int32)),
        ("mass", _P(C.c_double)), ("h", _P(C.c_double)), ("ks", _P(C.c_double)), ("p", _P(C.c_double)),
        ("gamma", _P(C.c_double)), ("c", _P(C.c_double)), ("T", _P(C.c_double)), ("Cp", _P(C.c_double)),
        ("rho_ph", _P(C.c_double)), ("x", _P(C.c_double)), ("v", _P(C.c_double)), ("sv", _P(C.c_double)),
        ("rho", _P(C.c_double)), ("e", _P(C.c_double)), ("se", _P(C.c_double)), ("grad", _P(C.c_double)),
        ("ncells", C.c_int64), ("cell_off", _P(C.c_int64)), ("cell_rows", _P(C.c_int32)),
        ("cell_order", _P(C.c_int32)), ("cell_attr", C.c_void_p),
        ("nb_off", _P(C.c_int64)), ("nb_idx", _P(C.c_int32)), ("removed", _P(C.c_int32)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        cp, sp = _P(Config), _P(State)
        for name in ("psph_oracle_bin", "psph_oracle_search", "psph_oracle_adaptive_sml",
                     "psph_oracle_calc_gradients", "psph_oracle_smooth_density", "psph_oracle_smooth_velocity"):
            getattr(L, name).argtypes = [cp, sp]
            getattr(L, name).restype = C.c_int
        L.psph_oracle_free_lists.argtypes = [sp]
        L.psph_oracle_free_lists.restype = None
        L.psph_oracle_timestep.argtypes = [cp, sp]
        L.psph_oracle_timestep.restype = C.c_double
        L.psph_oracle_mom_engr_update.argtypes = [cp, sp, C.c_double]
        L.psph_oracle_mom_engr_update.restype = C.c_int
        L.psph_oracle_update_pos.argtypes = [cp, sp, C.c_double, C.c_int]
        L.psph_oracle_update_pos.restype = C.c_int
        L.psph_oracle_rebin.argtypes = [cp, sp, _P(C.c_int)]
        L.psph_oracle_rebin.restype = C.c_int
        L.psph_oracle_apply_bcond.argtypes = [cp, sp, C.c_int64, _P(C.c_int32), _P(C.c_int64), _P(C.c_double)]
        L.psph_oracle_apply_bcond.restype = C.c_int
        L.psph_oracle_eos.argtypes = [cp, sp, C.c_int64]
        L.psph_oracle_eos.restype = None
        L.psph_oracle_table_order.argtypes = [cp, _P(C.c_uint32), C.c_int64, _P(C.c_int64)]
        L.psph_oracle_table_order.restype = None
        L.psph_oracle_weight.argtypes = [C.c_int, _P(C.c_double), C.c_double, C.c_double]
        L.psph_oracle_weight.restype = C.c_double
        L.psph_oracle_d_weight.argtypes = [C.c_int, _P(C.c_double), C.c_double, C.c_double, C.c_int]
        L.psph_oracle_d_weight.restype = C.c_double
        L.psph_oracle_add_new_erupt.argtypes = [cp, C.POINTER(Erupt), _P(C.c_double), C.c_int, C.c_double, C.c_double,
                                                C.c_uint32, C.c_void_p, C.c_int]
        L.psph_oracle_add_new_erupt.restype = C.c_int
        L.psph_oracle_hsfc3d.argtypes = [_P(C.c_double), _P(C.c_uint32)]
        L.psph_oracle_hsfc3d.restype = None
        L.psph_oracle_scan_outside_layer.argtypes = [cp, sp, C.c_void_p]
        L.psph_oracle_scan_outside_layer.restype = C.c_int
        L.psph_oracle_adapt_domain.argtypes = [cp, sp, C.c_void_p]
        L.psph_oracle_adapt_domain.restype = C.c_int
        L.psph_oracle_add_pressure_ghost.argtypes = [cp, sp, C.c_void_p, C.POINTER(Growth), _P(C.c_double), _P(C.c_double),
                                                     _P(C.c_double), C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.psph_oracle_add_pressure_ghost.restype = C.c_int
        L.psph_oracle_add_wall_ghost.argtypes = [cp, C.c_void_p, _P(C.c_int32), C.POINTER(Growth), _P(C.c_double), _P(C.c_double),
                                                 _P(C.c_double), C.c_void_p, C.c_uint32, C.c_void_p, C.c_int]
        L.psph_oracle_add_wall_ghost.restype = C.c_int
    return _lib


FIELDS_F8 = ["mass", "h", "ks", "p", "gamma", "c", "T", "Cp", "rho", "e", "se"]
FIELDS_V3 = ["x", "v", "sv"]
FIELDS_I4 = ["bc", "phase", "involved", "guest"]

# oracle/product column name -> reference dump field (oracle/refio.py DUMP_ROW)
DUMP_MAP = dict(mass="mass", h="smlen", ks="mass_frac", p="pressure", gamma="gamma", c="sound_speed",
                T="temperature", Cp="specific_heat_p", rho="rho", e="engr", se="smoothed_e",
                x="coord", v="vel", sv="smoothed_v", bc="bc_type", phase="phase", involved="involved",
                guest="guest", rho_ph="phase_density", grad="grad", key="key")


class OracleState:
    """numpy-backed particle table + mesh occupancy, bound to a psph_oracle_state."""

    def __init__(self, cfg, n):
        self.cfg = cfg
        self.n = n
        self.a = {}
        for f in FIELDS_F8:
            self.a[f] = np.zeros(n, np.float64)
        for f in FIELDS_V3:
            self.a[f] = np.zeros((n, 3), np.float64)
        for f in FIELDS_I4:
            self.a[f] = np.zeros(n, np.int32)
        self.a["key"] = np.zeros((n, 3), np.uint32)
        self.a["rho_ph"] = np.zeros((n, 2), np.float64)
        self.a["grad"] = np.zeros((n, 15), np.float64)
        self.a["removed"] = np.zeros(n, np.int32)
        self.ncells = int(cfg.ncell[0]) * int(cfg.ncell[1]) * int(cfg.ncell[2])
        self.cell_off = np.zeros(self.ncells + 1, np.int64)
        self.cell_rows = np.zeros(max(n, 1), np.int32)
        self.cell_order = None
        self.cell_attr = None
        self.st = State()
        self._bind()

    def _bind(self):
        st = self.st
        st.n = self.n
        for f in FIELDS_F8 + FIELDS_V3 + ["rho_ph", "grad"]:
            setattr(st, f, self.a[f].ctypes.data_as(_P(C.c_double)))
        for f in FIELDS_I4 + ["removed"]:
            setattr(st, f, self.a[f].ctypes.data_as(_P(C.c_int32)))
        st.key = self.a["key"].ctypes.data_as(_P(C.c_uint32))
        st.ncells = self.ncells
        st.cell_off = self.cell_off.ctypes.data_as(_P(C.c_int64))
        st.cell_rows = self.cell_rows.ctypes.data_as(_P(C.c_int32))
        st.cell_order = (self.cell_order.ctypes.data_as(_P(C.c_int32))
                         if self.cell_order is not None else _P(C.c_int32)())
        st.cell_attr = self.cell_attr.ctypes.data if self.cell_attr is not None else None

    @classmethod
    def from_dump(cls, cfg, rows):
        s = cls(cfg, len(rows))
        for f, src in DUMP_MAP.items():
            s.a[f][...] = rows[src]
        return s

    def set_cells_from_buckets(self, buckets):
        """Mesh occupancy + attributes from a reference bucket dump (refio.read_buckets):
        plists in the reference's order, buckets in HTIterator order."""
        cfg = self.cfg
        n1, n2 = int(cfg.ncell[1]), int(cfg.ncell[2])
        attr = np.zeros(self.ncells, CELL_ATTR_DTYPE)  # type 0 = brief (no Bucket object)
        order = []
        lists = {}
        for b in buckets:
            idx = [int(round((b["mincrd"][d] * cfg.len_scale - cfg.mindom[d]) / cfg.delta)) for d in range(3)]
            lin = (idx[0] * n1 + idx[1]) * n2 + idx[2]
            attr[lin]["type"] = b["type"]
            attr[lin]["has_involved"] = b["has_involved"]
            attr[lin]["active"] = b["active"]
            attr[lin]["bucket_index"] = b["bucket_index"]
            attr[lin]["bnd"] = b["bnd"]
            order.append(lin)
            lists[lin] = b["rows"]
        # HTIterator only visits existing buckets; append the rest so every cell is listed once
        seen = set(order)
        order += [c for c in range(self.ncells) if c not in seen]
        self.cell_order = np.asarray(order, np.int32)
        self.cell_attr = attr
        self.cell_off[:] = 0
        pos = 0
        for lin in sorted(lists):
            pass
        counts = np.zeros(self.ncells, np.int64)
        for lin, r in lists.items():
            counts[lin] = len(r)
        self.cell_off[1:] = np.cumsum(counts)
        for lin, r in lists.items():
            self.cell_rows[self.cell_off[lin]:self.cell_off[lin] + len(r)] = r
        self._bind()

    def bin(self):
        return lib().psph_oracle_bin(C.byref(self.cfg), C.byref(self.st))

    def search(self):
        return lib().psph_oracle_search(C.byref(self.cfg), C.byref(self.st))

    def neighbours(self):
        off = np.ctypeslib.as_array(self.st.nb_off, shape=(self.n + 1,)).copy()
        idx = np.ctypeslib.as_array(self.st.nb_idx, shape=(max(int(off[-1]), 1),))[:int(off[-1])].copy()
        return off, idx

    def timestep(self):
        return lib().psph_oracle_timestep(C.byref(self.cfg), C.byref(self.st))

    def adaptive_sml(self):
        return lib().psph_oracle_adaptive_sml(C.byref(self.cfg), C.byref(self.st))

    def calc_gradients(self):
        return lib().psph_oracle_calc_gradients(C.byref(self.cfg), C.byref(self.st))

    def mom_engr_update(self, dt):
        return lib().psph_oracle_mom_engr_update(C.byref(self.cfg), C.byref(self.st), dt)

    def smooth_density(self):
        return lib().psph_oracle_smooth_density(C.byref(self.cfg), C.byref(self.st))

    def smooth_velocity(self):
        return lib().psph_oracle_smooth_velocity(C.byref(self.cfg), C.byref(self.st))

    def update_pos(self, dt, erupting=1):
        return lib().psph_oracle_update_pos(C.byref(self.cfg), C.byref(self.st), dt, int(erupting))

    def apply_bcond(self, img_row, img_cell, img_coord):
        """img_row: row of every wall ghost that has an image, img_cell: linear cell of the image's bucket,
        img_coord [n][3]: reflection points"""
        r = np.ascontiguousarray(img_row, np.int32)
        c = np.ascontiguousarray(img_cell, np.int64)
        x = np.ascontiguousarray(img_coord, np.float64)
        return lib().psph_oracle_apply_bcond(C.byref(self.cfg), C.byref(self.st), len(r), r.ctypes.data_as(_P(C.c_int32)),
                                             c.ctypes.data_as(_P(C.c_int64)), x.ctypes.data_as(_P(C.c_double)))

    # domain growth: attr = CELL_ATTR_DTYPE array of the whole mesh, changed in place
    def scan_outside_layer(self, attr):
        return lib().psph_oracle_scan_outside_layer(C.byref(self.cfg), C.byref(self.st), attr.ctypes.data)

    def adapt_domain(self, attr):
        return lib().psph_oracle_adapt_domain(C.byref(self.cfg), C.byref(self.st), attr.ctypes.data)

    def add_pressure_ghost(self, attr, growth, meteo_z, meteo_rho, meteo_p, add_step, cell_mincrd=None, cap=1 << 20):
        out = np.zeros(cap, PARTICLE_DTYPE)
        mc = None if cell_mincrd is None else np.ascontiguousarray(cell_mincrd, np.float64)
        z, r, p = (np.ascontiguousarray(a, np.float64) for a in (meteo_z, meteo_rho, meteo_p))
        n = lib().psph_oracle_add_pressure_ghost(C.byref(self.cfg), C.byref(self.st), attr.ctypes.data, C.byref(growth),
                                                 z.ctypes.data_as(_P(C.c_double)), r.ctypes.data_as(_P(C.c_double)),
                                                 p.ctypes.data_as(_P(C.c_double)), None if mc is None else mc.ctypes.data,
                                                 int(add_step), out.ctypes.data, cap)
        if n < 0:
            raise RuntimeError("psph_oracle_add_pressure_ghost: more than %d particles" % cap)
        return out[:n]

    def rebin(self):
        lost = C.c_int(0)
        lib().psph_oracle_rebin(C.byref(self.cfg), C.byref(self.st), C.byref(lost))
        return lost.value

    def table_order(self):
        perm = np.zeros(self.n, np.int64)
        lib().psph_oracle_table_order(C.byref(self.cfg), self.st.key, self.n, perm.ctypes.data_as(_P(C.c_int64)))
        return perm

    def free(self):
        lib().psph_oracle_free_lists(C.byref(self.st))


def table_order(cfg, keys):
    """permutation that puts rows with these keys ([n][3] uint32) into THashTable iteration order"""
    k = np.ascontiguousarray(keys, np.uint32)
    perm = np.zeros(len(k), np.int64)
    lib().psph_oracle_table_order(C.byref(cfg), k.ctypes.data_as(_P(C.c_uint32)), len(k), perm.ctypes.data_as(_P(C.c_int64)))
    return perm


def add_new_erupt(cfg, erupt, boxes, time, dt, add_step, cap=65536):
    """psph_oracle_add_new_erupt: the rows (psph_particle layout) add_new_erupt adds at this time step"""
    boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 6)
    out = np.zeros(cap, PARTICLE_DTYPE)
    n = lib().psph_oracle_add_new_erupt(C.byref(cfg), C.byref(erupt), boxes.ctypes.data_as(_P(C.c_double)), len(boxes),
                                        float(time), float(dt), int(add_step), out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("psph_oracle_add_new_erupt: more than %d particles" % cap)
    return out[:n]


def add_wall_ghost(cfg, attr, cell_count, growth, meteo_z, meteo_rho, meteo_p, add_step, cell_mincrd=None, cap=1 << 20):
    """psph_oracle_add_wall_ghost (add_wall_ghost.cc:44-293): attr changed in place, returns the new rows"""
    out = np.zeros(cap, PARTICLE_DTYPE)
    cnt = np.ascontiguousarray(cell_count, np.int32)
    z, r, p = (np.ascontiguousarray(a, np.float64) for a in (meteo_z, meteo_rho, meteo_p))
    mc = None if cell_mincrd is None else np.ascontiguousarray(cell_mincrd, np.float64)
    n = lib().psph_oracle_add_wall_ghost(C.byref(cfg), attr.ctypes.data, cnt.ctypes.data_as(_P(C.c_int32)), C.byref(growth),
                                         z.ctypes.data_as(_P(C.c_double)), r.ctypes.data_as(_P(C.c_double)),
                                         p.ctypes.data_as(_P(C.c_double)), None if mc is None else mc.ctypes.data, int(add_step),
                                         out.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("psph_oracle_add_wall_ghost: %d" % n)
    return out[:n]


def growth_from_golden(z):
    """Growth record + the sounding's columns from a domain-growth fixture (tests/golden/make_golden.py::golden_grow)"""
    g = Growth()
    g.mass, g.smlen, g.gamma_air = (float(x) for x in z["param_const"])
    g.ratio_sml_dx, g.sndspd, g.gmm, g.Ta0 = 1.0, 340.0, 1.4, 0.0
    for k in range(3):
        g.kmin[k], g.kmax[k] = float(z["param_keybox"][k]), float(z["param_keybox"][3 + k])
        g.Ll[k], g.Lu[k] = float(z["param_domain"][k]), float(z["param_domain"][3 + k])
    g.atmosphere_type = int(z["param_step"][1])
    g.nmeteo = len(z["meteo_z"])
    return g, z["meteo_z"], z["meteo"][0], z["meteo"][1]


def erupt_from_params(names, v):
    """Erupt record from the parameter dump of a golden eruption fixture (tests/golden/make_golden.py ERUPT_PARAMS)"""
    d = dict(zip([str(n) for n in names], [float(x) for x in v]))
    e = Erupt()
    e.pos_v[0], e.pos_v[1] = d["pos_v0"], d["pos_v1"]
    e.rv, e.Vv0, e.vel_prof = d["rv"], d["Vv0"], int(d["vel_prof"])
    e.bot, e.sml, e.mass, e.ratio_sml_dx = d["bot"], d["sml_of_phase2"], d["mass_of_phase2"], d["ratio_sml_dx"]
    e.ev0, e.rhov, e.pv0, e.gamma_v, e.sndspd = d["ev0"], d["rhov"], d["pv0"], d["gamma_v"], d["sndspd"]
    for k in range(3):
        e.kmin[k], e.kmax[k] = d["kmin%d" % k], d["kmax%d" % k]
    return e, d
