"""ctypes binding of libpsph.so — one Python callable per entry point of include/psph.h.

Fails loudly: no library -> ImportError-like PsphError at load(); no GPU -> psph_create
returns -3 and Solver raises.  Nothing here computes anything on the CPU.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# PSPH_LIB: another build of the same library (A/B timing of compile-time variants, profiles/ab.sh)
LIB_PATH = os.environ.get("PSPH_LIB") or os.path.join(HERE, "libpsph.so")


class PsphError(RuntimeError):
    pass


class PsphConfig(C.Structure):
    """Mirror of ``psph_config`` (include/psph.h)."""
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

    def copy(self):
        c = PsphConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(PsphConfig))
        return c

    def set_mesh(self, lo, hi, smlen, particle_density=8):
        """Background mesh as src/preproc/preprocess.cc:99-133: bucket edge = PARTICLE_DENSITY*smlen,
        mindom = lo - 1.5 buckets, ncell = ceil((hi + 1.5 buckets - mindom)/edge)."""
        d = particle_density * smlen
        for k in range(3):
            self.mindom[k] = lo[k] - 1.5 * d
            self.ncell[k] = int(np.ceil((hi[k] + 1.5 * d - self.mindom[k]) / d))
        self.delta = d
        self.len_scale = 1.0
        self.smlen_real = smlen
        return self

    @property
    def ncells(self):
        return int(self.ncell[0]) * int(self.ncell[1]) * int(self.ncell[2])


# psph_particle (include/psph.h), array-of-structs host row
PARTICLE_DTYPE = np.dtype([
    ("key", "<u4", 3), ("bc_type", "<i4"), ("phase_num", "<i4"), ("involved", "<i4"), ("guest", "<i4"),
    ("myprocess", "<i4"),
    ("mass", "<f8"), ("smlen", "<f8"), ("mass_frac", "<f8"),
    ("pressure", "<f8"), ("gamma", "<f8"), ("sound_speed", "<f8"), ("temperature", "<f8"),
    ("specific_heat_p", "<f8"),
    ("phase_density", "<f8", 2),
    ("coords", "<f8", 3), ("vel", "<f8", 3), ("smoothed_v", "<f8", 3),
    ("rho", "<f8"), ("engr", "<f8"), ("smoothed_e", "<f8"),
    ("grad", "<f8", 15),
])
assert PARTICLE_DTYPE.itemsize == 328
# PSPH_ROW_BYTES_SPH rows: psph_particle without its gradient block (what an SPH build exchanges, pack_data.h:68-75)
PARTICLE_SPH_DTYPE = np.dtype([(n, PARTICLE_DTYPE.fields[n][0]) for n in PARTICLE_DTYPE.names if n != "grad"])
assert PARTICLE_SPH_DTYPE.itemsize == 208 and all(
    PARTICLE_SPH_DTYPE.fields[n][1] == PARTICLE_DTYPE.fields[n][1] for n in PARTICLE_SPH_DTYPE.names)

CELL_ATTR_DTYPE = np.dtype([("type", "<i4"), ("has_involved", "<i4"), ("active", "<i4"),
                            ("bucket_index", "<i4", 6), ("flags", "<i4"), ("bnd", "<f8", 6)])
assert CELL_ATTR_DTYPE.itemsize == 88
# psph_image (include/psph.h): one wall-ghost image
IMAGE_DTYPE = np.dtype([("ghost_key", "<u4", 3), ("cell", "<i4"), ("coord", "<f8", 3)])
assert IMAGE_DTYPE.itemsize == 40

class PsphErupt(C.Structure):
    """Mirror of ``psph_erupt`` (include/psph.h): what add_new_erupt reads (setup_erupt.cc:390-670)."""
    _fields_ = [("pos_v", C.c_double * 2), ("rv", C.c_double), ("Vv0", C.c_double), ("vel_prof", C.c_int32),
                ("pad", C.c_int32), ("bot", C.c_double), ("sml", C.c_double), ("mass", C.c_double),
                ("ratio_sml_dx", C.c_double), ("ev0", C.c_double), ("rhov", C.c_double), ("pv0", C.c_double),
                ("gamma_v", C.c_double), ("sndspd", C.c_double), ("kmin", C.c_double * 3), ("kmax", C.c_double * 3)]


class PsphGrowth(C.Structure):
    """Mirror of ``psph_growth`` (include/psph.h): what add_pressure_ghost reads (add_pressure_ghost.cc:36-316)."""
    _fields_ = [("mass", C.c_double), ("smlen", C.c_double), ("ratio_sml_dx", C.c_double), ("gamma_air", C.c_double),
                ("sndspd", C.c_double), ("gmm", C.c_double), ("Ta0", C.c_double), ("kmin", C.c_double * 3), ("kmax", C.c_double * 3),
                ("Ll", C.c_double * 3), ("Lu", C.c_double * 3), ("atmosphere_type", C.c_int32), ("nmeteo", C.c_int32)]


# every symbol include/psph.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_CFG = C.POINTER(PsphConfig)
API = {
    "psph_config_default_3d": (None, [_CFG]),
    "psph_config_default_1d": (None, [_CFG]),
    "psph_create": (C.c_int, [_CFG, C.POINTER(_VP)]),
    "psph_destroy": (None, [_VP]),
    "psph_last_error": (C.c_char_p, [_VP]),
    "psph_set_cells": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_get_cells": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_scan_outside_layer": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "psph_adapt_domain": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "psph_set_growth": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP]),
    "psph_add_pressure_ghost": (C.c_int, [_VP, C.c_uint32, C.POINTER(C.c_int)]),
    "psph_add_wall_ghost": (C.c_int, [_VP, C.c_uint32, C.POINTER(C.c_int)]),
    "psph_upload": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_append": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_download": (C.c_int, [_VP, _VP, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
    "psph_num_particles": (C.c_size_t, [_VP]),
    "psph_num_updated": (C.c_int64, [_VP]),
    "psph_column": (C.c_int, [_VP, C.c_int, C.POINTER(_VP)]),
    "psph_upload_device_soa": (C.c_int, [_VP, C.POINTER(_VP), _VP, _VP, C.c_size_t]),
    "psph_timestep": (C.c_int, [_VP, C.POINTER(C.c_double)]),
    "psph_set_dt": (C.c_int, [_VP, C.c_double, C.c_int, C.c_int]),
    "psph_search_neighs": (C.c_int, [_VP]),
    "psph_adaptive_sml": (C.c_int, [_VP]),
    "psph_calc_gradients": (C.c_int, [_VP]),
    "psph_mom_engr_update": (C.c_int, [_VP]),
    "psph_smooth_density": (C.c_int, [_VP]),
    "psph_smooth_velocity": (C.c_int, [_VP]),
    "psph_update_pos": (C.c_int, [_VP]),
    "psph_rebin": (C.c_int, [_VP, C.POINTER(C.c_int)]),
    "psph_set_erupt": (C.c_int, [_VP, _VP, _VP, C.c_size_t]),
    "psph_add_new_erupt": (C.c_int, [_VP, C.c_double, C.c_double, C.c_uint32, C.POINTER(C.c_int)]),
    "psph_set_images": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_apply_bcond": (C.c_int, [_VP]),
    "psph_upload_rows": (C.c_int, [_VP, _VP, C.c_size_t, C.c_size_t]),
    "psph_append_rows": (C.c_int, [_VP, _VP, C.c_size_t, C.c_size_t]),
    "psph_download_rows": (C.c_int, [_VP, _VP, C.c_size_t, C.POINTER(C.c_size_t), C.c_int, C.c_size_t]),
    "psph_download_h5part": (C.c_int, [_VP, _VP, _VP, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
    "psph_pair_stats": (C.c_int, [_VP, C.POINTER(C.c_int32)]),
    "psph_step": (C.c_int, [_VP, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "psph_run": (C.c_int, [_VP, C.c_int, C.c_int]),
    "psph_comm_unique_id": (C.c_int, [_VP, C.c_size_t]),
    "psph_comm_init": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_cell_ranks": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_slab_cuts": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_comm_bytes_sent": (C.c_int64, [_VP, C.c_int]),
    "psph_exchange_ghosts": (C.c_int, [_VP]),
    "psph_repartition": (C.c_int, [_VP]),
    "psph_emit_neighbours": (C.c_int, [_VP, _VP, _VP, C.c_size_t, C.POINTER(C.c_size_t)]),
    "psph_errors": (C.c_int, [_VP, C.c_int]),
    "psph_cell_of_rows": (C.c_int, [_VP, _VP, C.c_size_t]),
    "psph_launch_count": (C.c_int64, [_VP, C.c_int]),
    "psph_set_timing": (C.c_int, [_VP, C.c_int]),
    "psph_stage_ms": (C.c_double, [_VP, C.c_int, C.c_int]),
    "psph_kernel_ms": (C.c_double, [_VP, C.c_int, C.c_int, C.POINTER(C.c_int64)]),
    "psph_host_alloc": (_VP, [C.c_size_t]),
    "psph_host_free": (None, [_VP]),
    "psph_sync": (C.c_int, [_VP]),
    "psph_stream": (_VP, [_VP]),
}

_lib = None


def load():
    """dlopen libpsph.so and bind every API symbol; raises PsphError when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PsphError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in API.items():
        try:
            f = getattr(L, name)
        except AttributeError as e:
            raise PsphError(f"{LIB_PATH} does not export {name}") from e
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def config_default_3d():
    c = PsphConfig()
    load().psph_config_default_3d(C.byref(c))
    return c


def config_default_1d():
    c = PsphConfig()
    load().psph_config_default_1d(C.byref(c))
    return c
