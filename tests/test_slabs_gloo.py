"""CPU, world_size 2 over gloo: the host-side slab logic (plume_sph_b200/slabs.py) that mirrors
csrc/psph_comm.cu — weighted key-range cuts from all-reduced per-cell weights, owner table, migration
of rows to their owner, ghost ("guest") mirroring of boundary cells — checked against a single-process
evaluation of the same rules.  Replaces the reference's repartition (repartition_BSFC.cc:53-469) and
move_data (move_data.cc:48-427) at the planning level; the data movement itself is NCCL on the GPU
(tests/test_gpu_multi.py)."""
import ctypes as C
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hilbert_ranks(cfg):
    """rank of every cell along the bucket Hilbert curve (key of the cell centre, preprocess.cc:196-205)"""
    import oracle_py as op
    L = op.lib()
    n0, n1, n2 = (int(cfg.ncell[k]) for k in range(3))
    keys = np.zeros(n0 * n1 * n2, np.uint64)
    nrm, key = (C.c_double * 3)(), (C.c_uint32 * 2)()
    for i in range(n0):
        for j in range(n1):
            for k in range(n2):
                for d, idx in enumerate((i, j, k)):
                    nrm[d] = (idx + 0.5) / int(cfg.ncell[d])
                L.psph_oracle_hsfc3d(nrm, key)
                keys[(i * n1 + j) * n2 + k] = (np.uint64(key[0]) << np.uint64(32)) | np.uint64(key[1])
    order = np.argsort(keys, kind="stable")
    rank = np.zeros(len(keys), np.int32)
    rank[order] = np.arange(len(keys), dtype=np.int32)
    return rank


def cell_of(cfg, pos):
    idx = np.floor((pos - np.asarray([cfg.mindom[k] for k in range(3)])) / cfg.delta).astype(np.int64)
    for k in range(3):
        idx[:, k] = idx[:, k].clip(0, int(cfg.ncell[k]) - 1)
    return (idx[:, 0] * int(cfg.ncell[1]) + idx[:, 1]) * int(cfg.ncell[2]) + idx[:, 2]


def plan_single(cfg, rows, P):
    from plume_sph_b200 import slabs
    rank = hilbert_ranks(cfg)
    cell = cell_of(cfg, rows["coords"])
    nz = int(cfg.ncell[2])
    w = slabs.cell_weights(cell, rows["bc_type"], rank, nz)
    cuts = slabs.plan_cuts(w, P)
    owner = slabs.owners(rank, cuts, nz)
    mask = slabs.boundary_masks(owner, tuple(int(cfg.ncell[k]) for k in range(3)))
    return rank, cell, cuts, owner, mask


def worker(r, P, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    from plume_sph_b200 import cases, slabs
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=r, world_size=P)
    cfg, rows = cases.cube(16, seed=3)
    order = np.argsort(rows["coords"][:, 0], kind="stable")
    mine = rows[np.array_split(order, P)[r]]                   # arbitrary initial distribution
    rank = hilbert_ranks(cfg)
    cell = cell_of(cfg, mine["coords"])
    # psph_repartition: local weights in rank order -> all-reduce -> exact prefix cut
    nz = int(cfg.ncell[2])
    w = slabs.cell_weights(cell, mine["bc_type"], rank, nz)
    wt = torch.from_numpy(w)
    dist.all_reduce(wt)
    cuts = slabs.plan_cuts(wt.numpy(), P)
    owner = slabs.owners(rank, cuts, nz)
    mask = slabs.boundary_masks(owner, tuple(int(cfg.ncell[k]) for k in range(3)))
    # migrate(): rows whose cell belongs to another rank go there
    dest = owner[cell]
    send = [torch.from_numpy(np.ascontiguousarray(mine[dest == qq]).view(np.uint8).reshape(-1)) for qq in range(P)]
    counts = torch.tensor([len(x) for x in send])
    allc = [torch.zeros(P, dtype=torch.long) for _ in range(P)]
    dist.all_gather(allc, counts)
    recv = [torch.zeros(int(allc[qq][r]), dtype=torch.uint8) for qq in range(P)]
    reqs = []
    for qq in range(P):
        if qq != r:
            reqs.append(dist.isend(send[qq], qq))
            reqs.append(dist.irecv(recv[qq], qq))
    for x in reqs:
        x.wait()
    recv[r] = send[r]
    owned = np.concatenate([x.numpy().view(rows.dtype) for x in recv])
    # rebuild_ghosts(): boundary cells mirrored on the ranks that touch them
    ocell = cell_of(cfg, owned["coords"])
    gsend = [torch.from_numpy(np.ascontiguousarray(owned[(mask[ocell] >> qq) & 1 == 1]).view(np.uint8).reshape(-1))
             for qq in range(P)]
    gcounts = torch.tensor([len(x) for x in gsend])
    dist.all_gather(allc, gcounts)
    grecv = [torch.zeros(int(allc[qq][r]), dtype=torch.uint8) for qq in range(P)]
    reqs = []
    for qq in range(P):
        if qq != r:
            reqs.append(dist.isend(gsend[qq], qq))
            reqs.append(dist.irecv(grecv[qq], qq))
    for x in reqs:
        x.wait()
    guests = np.concatenate([grecv[qq].numpy().view(rows.dtype) for qq in range(P) if qq != r])
    q.put((r, cuts, owned["key"].copy(), guests["key"].copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_plan_matches_single_process():
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    from plume_sph_b200 import cases
    P = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=worker, args=(r, P, port, q)) for r in range(P)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(P):
        r, cuts, owned, guests = q.get(timeout=240)
        got[r] = (cuts, owned, guests)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, rows = cases.cube(16, seed=3)
    rank, cell, cuts, owner, mask = plan_single(cfg, rows, P)
    key = lambda k: set(map(tuple, k.tolist()))
    for r in range(P):
        assert got[r][0] == cuts
        assert key(got[r][1]) == key(rows["key"][owner[cell] == r])                 # owned rows
        assert key(got[r][2]) == key(rows["key"][(owner[cell] != r) & ((mask[cell] >> r) & 1 == 1)])  # guests
    assert sum(len(got[r][1]) for r in range(P)) == len(rows)
    # the cut balances the reference's load weights to within one cell
    from plume_sph_b200 import slabs
    w = [int(slabs.particle_weights(rows["bc_type"][owner[cell] == r]).sum()) for r in range(P)]
    assert abs(w[0] - w[1]) <= 2 * 512 * 200



def test_underground_bucket_goes_with_the_one_above():
    """preprocess.cc:219-239, :264-275: a bucket of the mesh's bottom layer is sorted by the key of the bucket above it, so
    both land on the same rank, and its load (wall ghosts: 0.75 each) is counted there"""
    from plume_sph_b200 import slabs
    n = (4, 3, 5)
    ncells = n[0] * n[1] * n[2]
    rng = np.random.default_rng(2)
    rank = rng.permutation(ncells).astype(np.int32)
    cell = rng.integers(0, ncells, 4000)
    bc = rng.choice([100, 2, 1, 0], 4000)
    w = slabs.cell_weights(cell, bc, rank, n[2])
    bottom = np.arange(ncells) % n[2] == 0
    assert w[rank[bottom]].sum() == 0 and w.sum() == slabs.particle_weights(bc).sum()
    for P in (2, 3, 4):
        owner = slabs.owners(rank, slabs.plan_cuts(w, P), n[2]).reshape(n)
        assert np.array_equal(owner[:, :, 0], owner[:, :, 1])
