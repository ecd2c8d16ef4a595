"""Key-range slabs: how the particle table is split across the GPUs of one box.

Replaces the reference's Hilbert-curve bucket partition and its repartitioning
(src/preproc/preprocess.cc:219-279 initial equal-count split, src/multiproc/repartition_BSFC.cc:53-469
weighted re-split) — host-side planning only; the data movement is psph_exchange_ghosts /
psph_repartition in csrc/psph_comm.cu.
"""
import numpy as np


def setup_cube_ranks(torch, dist, psph, cfg, n, dev, rank, world, cube_columns):
    """Create this rank's Solver holding its share of the synthetic cube (bench.py).
    -> (solver, real particles owned, rows held)"""
    s = psph.Solver(cfg)
    if world == 1:
        cols, bc, phase, nreal = cube_columns(torch, cfg, n, dev)
        torch.cuda.synchronize()
        s.upload_device_soa([c.data_ptr() for c in cols], bc.data_ptr(), phase.data_ptr(), cols[0].numel())
        s.sync()
        nrows = cols[0].numel()
        del cols, bc, phase
        torch.cuda.empty_cache()
        return s, nreal, nrows
    raise psph.PsphError("multi-GPU slabs: not built yet")
