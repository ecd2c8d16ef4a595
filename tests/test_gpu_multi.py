"""-m gpu, needs >= 2 CUDA devices (skipped otherwise; run with `gpurun --gpus 2`): the key-range
slab path — psph_comm_init / psph_repartition / psph_exchange_ghosts (csrc/psph_comm.cu, replacing
move_data.cc / repartition_BSFC.cc) — against the single-GPU path on the same particles.

Two ranks run as two threads of this process, one context per device (ctypes releases the GIL
around every C call, NCCL point-to-point between the two devices)."""
import threading
import numpy as np
import pytest

import plume_sph_b200 as psph
from plume_sph_b200 import cases, slabs
from plume_sph_b200.solver import ORDER_REFERENCE, ORDER_DEVICE
from common import relerr

pytestmark = pytest.mark.gpu


def ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def run_ranks(P, fn):
    out, err = [None] * P, [None] * P

    def work(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # surface in the main thread
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    for e in err:
        if e is not None:
            raise e
    assert all(o is not None for o in out), "a rank did not finish"
    return out


@pytest.mark.skipif(ndev() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("P", [2, 4])
def test_slabs_match_single_gpu(P):
    if ndev() < P:
        pytest.skip("needs %d GPUs" % P)
    n, steps = 24, 3
    cfg, rows = cases.cube(n, seed=5)
    # single GPU
    s1 = psph.Solver(cfg)
    s1.upload(rows)
    dts = [s1.step(0.0, 0)[0] for _ in range(steps)]
    ref = s1.download(ORDER_REFERENCE)
    s1.close()
    uid = psph.Solver.comm_unique_id()
    order = np.argsort(rows["coords"][:, 0], kind="stable")   # arbitrary initial distribution: x-halves
    parts = np.array_split(order, P)

    def rank(r):
        c = cfg.copy()
        c.device, c.rank, c.nranks = r, r, P
        s = psph.Solver(c)
        s.upload(rows[parts[r]])
        s.comm_init(uid)
        s.repartition()
        cuts = s.slab_cuts()
        ranks = s.cell_ranks()
        got = [s.step(0.0, 0)[0] for _ in range(steps)]
        out = s.download(ORDER_DEVICE)
        sent = s.comm_bytes_sent()
        s.close()
        return cuts, got, out, sent, ranks

    res = run_ranks(P, rank)
    # the device cut (k_cell_weights -> ncclAllReduce -> scan -> k_find_cuts) equals the host plan (slabs.plan_cuts)
    ranks = res[0][4]
    mind = np.asarray([cfg.mindom[k] for k in range(3)])
    idx = np.floor((rows["coords"] - mind) / cfg.delta).astype(np.int64)
    cell = (idx[:, 0] * int(cfg.ncell[1]) + idx[:, 1]) * int(cfg.ncell[2]) + idx[:, 2]
    w = np.zeros(cfg.ncells, np.int64)
    np.add.at(w, ranks[cell], slabs.particle_weights(rows["bc_type"]))
    assert res[0][0] == slabs.plan_cuts(w, P)
    res = [x[:4] for x in res]
    for cuts, got, out, sent in res:
        assert cuts == res[0][0] and cuts[0] == 0 and cuts[-1] == cfg.ncells
        assert got == dts, (got, dts)                      # all-reduced dt: identical on every rank and to 1 GPU
        assert sent > 0
    owned = np.concatenate([o[o["guest"] == 0] for _, _, o, _ in res])
    # every particle is owned by exactly one rank
    k = owned["key"].astype(np.uint64)
    ko = (k[:, 0] << np.uint64(32)) | k[:, 1]
    kr = (ref["key"][:, 0].astype(np.uint64) << np.uint64(32)) | ref["key"][:, 1].astype(np.uint64)
    assert len(ko) == len(kr) and len(np.unique(ko)) == len(ko)
    a = owned[np.argsort(ko)]
    b = ref[np.argsort(kr)]
    assert np.array_equal(a["key"], b["key"])
    for f in ("coords", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure"):
        e = relerr(a[f], b[f])
        assert e <= 1e-12, (f, e)
    # guests mirror their owners' final state exactly
    for _, _, o, _ in res:
        g = o[o["guest"] != 0]
        kg = (g["key"][:, 0].astype(np.uint64) << np.uint64(32)) | g["key"][:, 1].astype(np.uint64)
        src = a[np.searchsorted(np.sort(ko), kg)]
        assert np.array_equal(src["key"], g["key"])
        for f in ("coords", "vel", "rho", "engr"):
            assert np.array_equal(src[f], g[f]), f
    # load balance: the weighted cut (reference weights) is within one cell of even
    w = [int(slabs.particle_weights(o[o["guest"] == 0]["bc_type"]).sum()) for _, _, o, _ in res]
    assert max(w) - min(w) <= 2 * 512 * 200, w
