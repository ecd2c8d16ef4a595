"""Key-range slabs: how the particle table is split across the GPUs of one box.

Host-side planning that mirrors what csrc/psph_comm.cu does on the device (and is what the
world_size-2 gloo tests exercise on CPU):

* ``plan_cuts``      weighted cut of the cells in Hilbert-rank order into contiguous key ranges
                     (replaces src/preproc/preprocess.cc:219-279 and the BSFC bin refinement of
                     src/multiproc/repartition_BSFC.cc:53-469, BSFC_create_bins.cc:57-141);
* ``owners``         owner rank of every cell from the cuts;
* ``boundary_masks`` which ranks need a mirror ("guest" copy, src/multiproc/move_data.cc:236-299)
                     of each cell;
* ``particle_weights`` the reference's load weights (parameters.h:1035-1038).
"""
import numpy as np

WEIGHT_X100 = {100: 200, 99: 200, 2: 75, 0: 2, 1: 0}   # real 2.0, wall 0.75, erupt 0.02, pressure 0


def particle_weights(bc_type):
    w = np.zeros(len(bc_type), np.int64)
    for bc, v in WEIGHT_X100.items():
        w[bc_type == bc] = v
    return w


def plan_cuts(cell_weight_rank_order, nranks):
    """cuts[p] = number of cells (rank order) owned by ranks < p: the smallest prefix that carries
    >= p/nranks of the total weight (exact integer arithmetic, same rule as k_find_cuts)."""
    w = np.asarray(cell_weight_rank_order, np.int64)
    prefix = np.cumsum(w)
    total = int(prefix[-1]) if len(prefix) else 0
    cuts = [0]
    for p in range(1, nranks):
        r = int(np.searchsorted(prefix * nranks, total * p, side="left"))
        cuts.append(max(min(r + 1, len(w)), cuts[-1]))
    cuts.append(len(w))
    return cuts


def partition_cell(cell, nz):
    """the cell whose key decides where `cell` goes: a bucket of the bottom layer (k == 0, underground) is sorted by the key
    of the bucket above it (PartiHead top_key, preprocess.cc:219-239, :264-275), so the two always share a rank"""
    cell = np.asarray(cell, np.int64)
    return np.where((nz > 1) & (cell % nz == 0), cell + 1, cell)


def cell_weights(cell, bc_type, cell2rank, nz):
    """per-rank-position load (x100) from the particles' cells: what k_cell_weights accumulates"""
    w = np.zeros(len(cell2rank), np.int64)
    np.add.at(w, np.asarray(cell2rank)[partition_cell(cell, nz)], particle_weights(bc_type))
    return w


def owners(cell2rank, cuts, nz=1):
    """owner rank of every cell (linear cell index -> rank)"""
    c2r = np.asarray(cell2rank)
    r = c2r[partition_cell(np.arange(len(c2r)), nz)]
    return (np.searchsorted(np.asarray(cuts[1:-1]), r, side="right")).astype(np.int32)


def boundary_masks(owner, ncell):
    """mask[c] bit q set: rank q != owner[c] owns a cell of c's 27-neighbourhood (bucket.cc:23-26)"""
    o = np.asarray(owner).reshape(ncell)
    mask = np.zeros(o.shape, np.uint32)
    pad = np.pad(o, 1, constant_values=-1)
    for di in (0, 1, 2):
        for dj in (0, 1, 2):
            for dk in (0, 1, 2):
                nb = pad[di:di + o.shape[0], dj:dj + o.shape[1], dk:dk + o.shape[2]]
                ok = (nb >= 0) & (nb != o)
                mask |= np.where(ok, np.uint32(1) << nb.clip(0).astype(np.uint32), np.uint32(0))
    return mask.reshape(-1)


def setup_cube_ranks(torch, dist, psph, cfg, n, dev, rank, world, cube_columns):
    """Create this rank's Solver holding its share of the synthetic cube (bench.py): every rank
    generates an x-block of the lattice on its GPU, then the slabs are balanced by weight and the
    rows migrate to their owners (psph_repartition).  -> (solver, real particles owned, rows held)"""
    s = psph.Solver(cfg)
    g = 8
    lo, hi = -g + (n + 2 * g) * rank // world, -g + (n + 2 * g) * (rank + 1) // world
    cols, bc, phase, _ = cube_columns(torch, cfg, n, dev, lo=(lo, -g, -g), hi=(hi, n + g, n + g))
    torch.cuda.synchronize()
    s.upload_device_soa([c.data_ptr() for c in cols], bc.data_ptr(), phase.data_ptr(), cols[0].numel())
    s.sync()
    del cols, bc, phase
    torch.cuda.empty_cache()
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(psph.Solver.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        s.comm_init(bytes(idt.cpu().numpy().tobytes()))
        s.repartition()
    nreal = s.num_updated
    return s, nreal, s.n
