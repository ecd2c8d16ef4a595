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
import oracle_py as op
from common import relerr, golden, to_product_cfg, dump_to_particles, cell_attrs, image_cells

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

    # daemon threads: when one rank fails, the others sit in an NCCL call that never returns — they must not keep the
    # interpreter alive after pytest has reported (profiles/r02m_*: a 15-minute hang on an 8-GPU box)
    th = [threading.Thread(target=work, args=(r,), daemon=True) for r in range(P)]
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
    w = slabs.cell_weights(cell, rows["bc_type"], ranks, int(cfg.ncell[2]))
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



def key64(rows):
    k = rows["key"].astype(np.uint64)
    return (k[:, 0] << np.uint64(32)) | k[:, 1]


@pytest.mark.skipif(ndev() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("P", [2, 4])
def test_plume_with_wall_ghosts_and_vent_across_slabs(P):
    """the reference's small plume (tests/golden/plume3d_small.npz: air, pressure ghosts, wall ghosts with their image
    table, the duct's erupt ghosts) advanced 4 steps — psph_step, psph_add_new_erupt, psph_apply_bcond — on P key-range
    slabs against one GPU: images are computed by the owner of their bucket and travel back to the ghost's owner
    (move_bnd_img.cc:32-210, send_reflections.cc:34-165), every rank injects the particles born in its own buckets
    (setup_erupt.cc:609-650).  Same particles on exactly one rank each, fields within 1e-12."""
    if ndev() < P:
        pytest.skip("needs %d GPUs" % P)
    z, ze = golden("plume3d_small"), golden("erupt3d")
    ocfg = op.config_for("plume3d_small", z)
    cfg = to_product_cfg(ocfg)
    attr, _ = cell_attrs(ocfg, z)
    rows = dump_to_particles(z["pre"])
    img = np.zeros(len(z["bc_img_row"]), psph.IMAGE_DTYPE)
    img["ghost_key"] = z["bc_pre"]["key"][z["bc_img_row"]]
    img["cell"] = image_cells(ocfg, z)
    img["coord"] = z["bc_img_coord"]
    er, _ = op.erupt_from_params(ze["param_names"], ze["s2_params"])
    boxes = np.concatenate([ze["s2_cell_min"], ze["s2_cell_max"]], axis=1)
    steps = 4

    def advance(s, multi):
        t, dts, added = 0.0, [], []
        for k in range(1, steps + 1):
            dt, lost, ierr = s.step(0.0, 1)
            assert ierr == 0
            t += dt
            dts.append(dt)
            added.append(s.add_new_erupt(t, dt, k))
            assert s.apply_bcond() == 0
        if multi:
            s.exchange_ghosts()          # the wall ghosts' new state reaches the guest copies (main.cc:471)
        return dts, added

    s1 = psph.Solver(cfg)
    s1.set_cells(attr); s1.upload(rows); s1.set_images(img); s1.set_erupt(er, boxes)
    dts, added1 = advance(s1, False)
    ref = s1.download(ORDER_REFERENCE)
    s1.close()
    assert sum(added1) > 0                                   # the vent did inject
    uid = psph.Solver.comm_unique_id()
    parts = np.array_split(np.argsort(rows["coords"][:, 0], kind="stable"), P)

    def rank(r):
        c = cfg.copy()
        c.device, c.rank, c.nranks = r, r, P
        s = psph.Solver(c)
        s.set_cells(attr); s.upload(rows[parts[r]]); s.set_images(img); s.set_erupt(er, boxes)
        s.comm_init(uid)
        s.repartition()
        got, added = advance(s, True)
        out = s.download(ORDER_DEVICE)
        s.close()
        return got, added, out

    res = run_ranks(P, rank)
    for got, added, out in res:
        assert got == dts, (got, dts)
    assert [sum(a[k] for _, a, _ in res) for k in range(steps)] == added1     # every new particle on exactly one rank
    owned = np.concatenate([o[o["guest"] == 0] for _, _, o in res])
    ko, kr = key64(owned), key64(ref)
    assert len(ko) == len(kr) and len(np.unique(ko)) == len(ko)
    a, b = owned[np.argsort(ko)], ref[np.argsort(kr)]
    assert np.array_equal(a["key"], b["key"]) and np.array_equal(a["bc_type"], b["bc_type"])
    for f in ("coords", "vel", "smoothed_v", "rho", "engr", "mass_frac", "pressure", "temperature", "smlen"):
        e = relerr(a[f], b[f])
        assert e <= 1e-12, (f, e)
    wall = a["bc_type"] == 2
    assert wall.any() and relerr(a["vel"][wall], b["vel"][wall]) <= 1e-12
    # an underground bucket (bottom layer of the mesh) sits on the rank of the bucket above it (preprocess.cc:264-275)
    mind = np.asarray([cfg.mindom[k] for k in range(3)])
    rank_of_cell = {}
    for r, (_, _, o) in enumerate(res):
        own = o[o["guest"] == 0]
        idx = np.floor((own["coords"] * cfg.len_scale - mind) / cfg.delta).astype(np.int64)
        for c in set(map(tuple, idx.tolist())):
            rank_of_cell.setdefault(c, set()).add(r)
    assert all(len(v) == 1 for v in rank_of_cell.values())
    bottom = [c for c in rank_of_cell if c[2] == 0 and (c[0], c[1], 1) in rank_of_cell]
    assert bottom and all(rank_of_cell[c] == rank_of_cell[(c[0], c[1], 1)] for c in bottom)
    for _, _, o in res:                                      # guests mirror their owners
        g = o[o["guest"] != 0]
        src = a[np.searchsorted(np.sort(ko), key64(g))]
        assert np.array_equal(src["key"], g["key"])
        for f in ("coords", "vel", "rho", "engr", "pressure"):
            assert np.array_equal(src[f], g[f]), f


@pytest.mark.skipif(ndev() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("P", [2, 4])
def test_domain_growth_across_slabs(P):
    """psph_scan_outside_layer / psph_adapt_domain / psph_add_pressure_ghost on P key-range slabs == the reference's calls
    (tests/golden/grow3d.npz; the single-GPU comparison is tests/test_gpu_grow.py): the per-bucket facts are summed over the
    slabs, every rank takes the same decision for every bucket (the attribute tables stay identical without being
    exchanged), flips the rows it holds and fills the buckets it owns — every new particle exists on exactly one rank"""
    if ndev() < P:
        pytest.skip("needs %d GPUs" % P)
    from common import grow_cell_attrs, grow_cell_mincrd
    from test_gpu_grow import growth_record
    z = golden("grow3d")
    ocfg = op.config_for("grow3d")
    cfg = to_product_cfg(ocfg)
    attr0, lin = grow_cell_attrs(ocfg, z, "scan_pre")
    table = z["table"]
    rows = dump_to_particles(table)
    g, mz, mrho, mp = growth_record(z)
    mincrd = grow_cell_mincrd(ocfg, z, lin)
    attr_w, lin_w = grow_cell_attrs(ocfg, z, "wghost_pre")
    mincrd_w = grow_cell_mincrd(ocfg, z, lin_w, "wghost_pre")
    # (everything the rank threads need is read off the fixture HERE: an NpzFile is not safe to read from several threads)
    add_step = int(z["param_step"][0])
    active_pre = grow_cell_attrs(ocfg, z, "adapt_pre")[0]["active"]
    uid = psph.Solver.comm_unique_id()
    parts = np.array_split(np.argsort(rows["coords"][:, 0], kind="stable"), P)

    def rank(r):
        c = cfg.copy()
        c.device, c.rank, c.nranks = r, r, P
        s = psph.Solver(c)
        s.set_cells(attr0); s.upload(rows[parts[r]])
        s.comm_init(uid)
        s.repartition()
        s.search_neighs_consth()
        adapt = s.scan_outside_layer()
        a = s.get_cells()
        a["active"] = active_pre
        s.set_cells(a)
        turned = s.adapt_domain()
        s.set_growth(g, mz, mrho, mp, mincrd)
        added = s.add_pressure_ghost(add_step)
        s.exchange_ghosts()
        attr = s.get_cells()
        out = s.download(ORDER_DEVICE)
        # add_wall_ghost after the driver's shift_brief_buck (the fixture's bucket set at that point)
        s.set_cells(attr_w)
        s.set_growth(g, mz, mrho, mp, mincrd_w)
        nwall = s.add_wall_ghost(add_step)
        s.exchange_ghosts()
        outw = s.download(ORDER_DEVICE)
        attrw = s.get_cells()
        dt, lost, ierr = s.step(0.0, 1)                   # the grown, re-balanced table steps
        s.close()
        return adapt, turned, added, attr, out, ierr, nwall, outw, attrw

    res = run_ranks(P, rank)
    want_w, _ = grow_cell_attrs(ocfg, z, "wghost_post")
    assert sum(r[6] for r in res) == len(z["wghost_rows"])
    ownw = np.concatenate([r[7][r[7]["guest"] == 0] for r in res])
    wk = ownw[ownw["bc_type"] == 2]
    refw = np.concatenate([table[table["bc_type"] == 2]["key"], z["wghost_rows"]["key"]])
    assert np.array_equal(wk["key"][np.lexsort(wk["key"].T[::-1])], refw[np.lexsort(refw.T[::-1])])   # every wall ghost on exactly one rank
    for r in res:
        for f in ("type", "has_involved", "flags", "active"):
            assert np.array_equal(r[8][f], want_w[f]), ("wghost_post", f)
    want, _ = grow_cell_attrs(ocfg, z, "pghost_post")
    assert all(r[0] == 1 for r in res)                                   # the global count on every rank (main.cc:226-258)
    assert sum(r[1] for r in res) == len(z["adapt_rows"])
    assert sum(r[2] for r in res) == len(z["pghost_rows"])
    for r in res:
        assert r[5] == 0
        for f in ("type", "has_involved", "flags"):
            assert np.array_equal(r[3][f], want[f]), f
    owned = np.concatenate([r[4][r[4]["guest"] == 0] for r in res])
    ko = key96(owned)
    assert len(np.unique(ko, axis=0)) == len(ko) == len(table) + len(z["pghost_rows"])
    # same rows as the reference's table after the three calls
    bc, inv = table["bc_type"].copy(), table["involved"].copy()
    bc[z["adapt_rows"]] = z["adapt_bc_type"]
    inv[z["adapt_rows"]] = z["adapt_involved"]
    ref_keys = np.concatenate([table["key"], z["pghost_rows"]["key"]])
    ref_bc = np.concatenate([bc, z["pghost_rows"]["bc_type"]])
    ref_inv = np.concatenate([inv, z["pghost_rows"]["involved"]])
    ref_x = np.concatenate([table["coord"], z["pghost_rows"]["coord"]])
    ref_rho = np.concatenate([table["rho"], z["pghost_rows"]["rho"]])
    oa = np.lexsort(owned["key"].T[::-1]); ob = np.lexsort(ref_keys.T[::-1])
    assert np.array_equal(owned["key"][oa], ref_keys[ob])
    assert np.array_equal(owned["bc_type"][oa], ref_bc[ob]) and np.array_equal(owned["involved"][oa], ref_inv[ob])
    assert np.array_equal(owned["coords"][oa], ref_x[ob]) and np.array_equal(owned["rho"][oa], ref_rho[ob])


def key96(rows):
    return rows["key"].astype(np.uint64)
