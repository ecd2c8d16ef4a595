"""plume_sph_b200 — B200-native (sm_100a) implementation of Plume-SPH's per-time-step particle
hot path.  The product is the C-ABI library ``libpsph.so`` (include/psph.h, sources in
``csrc/``); this package is the thin Python host layer used by the tests and bench.py:

* :mod:`plume_sph_b200.lib`     ctypes binding of every entry point of include/psph.h
* :mod:`plume_sph_b200.solver`  ``Solver``: the reference's stage API (particler.h:22-357) on top of it
* :mod:`plume_sph_b200.cases`   synthetic particle clouds of the BASELINE shapes + 1-D shock tubes
* :mod:`plume_sph_b200.slabs`   key-range slab decomposition / ghost layers for multi-GPU runs

There is no CPU fallback: importing :mod:`lib` without a built ``libpsph.so`` raises, and
``psph_create`` fails when no CUDA device is visible.
"""
from .lib import (PsphConfig, PsphErupt, PsphGrowth, PARTICLE_DTYPE, PARTICLE_SPH_DTYPE, CELL_ATTR_DTYPE, IMAGE_DTYPE, PsphError, load, config_default_3d,
                  config_default_1d, LIB_PATH)
from .solver import Solver

__all__ = ["PsphConfig", "PsphErupt", "PsphGrowth", "PARTICLE_DTYPE", "PARTICLE_SPH_DTYPE", "CELL_ATTR_DTYPE", "IMAGE_DTYPE", "PsphError", "load", "config_default_3d",
           "config_default_1d", "Solver", "LIB_PATH"]
