"""Background mesh without the `preprocess` round trip (second half).

The reference builds its bucket lattice in a separate program (src/preproc/preprocess.cc:99-205) that writes
funkyNNNN.h5, and Read_Grid (src/plume/dataread.cc:320-470) turns those records into Bucket / BriefBucket objects.
Everything in those records is a closed-form function of the domain box, the smoothing length and the initial
domain, restated here so that a driver can fill ``psph_config`` (mindom, del, ncell) and the ``psph_cell_attr``
array directly:

* lattice: mindom = Ll - EXT_DOM_COF * del (EXT_DOM_COF_BOT along z), n = ceil((maxdom - mindom) / del),
  del = PARTICLE_DENSITY * smlen (preprocess.cc:99-133);
* bucket type and bucket_index from the position of the bucket's faces relative to the ORIGINAL box
  (determine_bucket_type / determine_face_type, preprocess.cc:386-470);
* bnd: the original box faces a MIXED bucket straddles, 0.1 elsewhere (Bucket ctor, bucket.cc:108-143);
* non-brief ("exists as a Bucket") and has_involved from the initial domain (determine_bucket_involved_flag,
  dataread.cc:52-75, :336-340, :414, :455).
"""
import numpy as np

from . import lib as _l

UNDERGROUND, MIXED, OVERGROUND, PRESS_BC = 10, 11, 12, 13   # constant.h:144-147


def _face_type(crd, hi, lo):
    """determine_face_type (preprocess.cc:386-398): -1 below the box, +1 above, 0 inside (both compares strict)"""
    return np.where(crd < lo, -1, np.where(crd > hi, 1, 0)).astype(np.int32)


def lattice(Ll, Lu, smlen, particle_density=8, ext=1.5, ext_bot=1.5):
    """(mindom[3], ncell[3], del) of the bucket lattice, preprocess.cc:99-133"""
    d = particle_density * smlen
    Ll, Lu = np.asarray(Ll, np.float64), np.asarray(Lu, np.float64)
    e = np.asarray([ext, ext, ext_bot])
    mindom = Ll - e * d
    maxdom = Lu + e * d
    n = np.ceil((maxdom - mindom) / d).astype(np.int64)
    return mindom, n, d


def background_mesh(Ll, Lu, smlen, idom_min, idom_max, particle_density=8, ext=1.5, ext_bot=1.5, len_scale=1.0):
    """-> (mindom, ncell, del, attr) with attr a CELL_ATTR_DTYPE array over ALL cells (x-major, z-minor), type 0 for the
    cells the reference keeps as BriefBucket only; `active` is a run-time flag of the reference and is left 0."""
    mindom, n, d = lattice(Ll, Lu, smlen, particle_density, ext, ext_bot)
    Ll, Lu = np.asarray(Ll, np.float64), np.asarray(Lu, np.float64)
    idx = np.stack(np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij"), axis=-1).reshape(-1, 3)
    lo = mindom + idx * d                                    # xcrd[0], ycrd[0], zcrd[0] (preprocess.cc:169-181)
    hi = mindom + (idx + 1) * d
    bt = np.zeros((len(idx), 6), np.int32)
    for k in range(3):
        bt[:, 2 * k] = _face_type(lo[:, k], Lu[k], Ll[k])
        bt[:, 2 * k + 1] = _face_type(hi[:, k], Lu[k], Ll[k])
    flag = np.abs(bt[:, 0::2] + bt[:, 1::2])
    out2 = (flag == 2).any(axis=1)
    typ = np.where(out2, np.where((bt[:, 4] == -1) & (bt[:, 5] == -1), UNDERGROUND, PRESS_BC),
                   np.where((flag == 1).any(axis=1), MIXED, OVERGROUND))
    typ = np.where((bt[:, 4] == -1) & (bt[:, 5] == 0), MIXED, typ)      # preprocess.cc:464-467
    # Read_Grid: a Bucket object only where the (scaled) bucket touches the initial domain grown by two buckets
    lo_s, hi_s = lo / len_scale, hi / len_scale
    bsize = d / len_scale
    imin, imax = np.asarray(idom_min, np.float64), np.asarray(idom_max, np.float64)

    def involved(dmin, dmax):
        b_lo = np.stack([_face_type(lo_s[:, k], dmax[k], dmin[k]) for k in range(3)], axis=1)
        b_hi = np.stack([_face_type(hi_s[:, k], dmax[k], dmin[k]) for k in range(3)], axis=1)
        return ~(np.abs(b_lo + b_hi) == 2).any(axis=1)

    exists = involved(imin - 2.0 * bsize, imax + 2.0 * bsize)
    attr = np.zeros(len(idx), _l.CELL_ATTR_DTYPE)
    attr["type"] = np.where(exists, typ, 0)
    attr["has_involved"] = (exists & involved(imin, imax)).astype(np.int32)
    attr["bucket_index"] = np.where(exists[:, None], bt, 0)
    flat = np.stack([Ll, Lu], axis=1).reshape(-1)            # flat[2k] = Ll[k], flat[2k+1] = Lu[k] (dataread.cc:357-363)
    bnd = np.where((typ == MIXED)[:, None] & (bt != 0), flat[None, :], 0.1)
    attr["bnd"] = np.where(exists[:, None], bnd, 0.0)
    return mindom, n, d, attr


def _neigh_sum(flag, n):
    """for every cell: how many of its 26 lattice neighbours carry `flag` (cells outside the mesh do not exist)"""
    f = np.pad(np.asarray(flag, np.int32).reshape(n), 1)
    s = np.zeros(n, np.int32)
    for di in range(3):
        for dj in range(3):
            for dk in range(3):
                if (di, dj, dk) != (1, 1, 1):
                    s += f[di:di + n[0], dj:dj + n[1], dk:dk + n[2]]
    return s.reshape(-1)


def shift_brief_buck(attr, latent, ncell):
    """shift_brief_buck (src/plume/shift_brief_buck.cc:33-129): every brief bucket (type 0) next to a bucket that is not involved,
    holds pressure ghosts and lies outside the vent becomes a bucket — with the attributes `switch_brief` computes
    (sph_lib.cc:1442-1560), which are the closed-form ones of `background_mesh` (`latent` = its attr array built with an
    initial domain that covers the whole mesh).  attr is changed in place; returns the linear indices of the new buckets."""
    n = [int(x) for x in ncell]
    src = (attr["type"] != 0) & (attr["has_involved"] == 0) & ((attr["flags"] & 4) != 0) & (((attr["flags"] >> 8) & 3) == 0)
    new = np.flatnonzero((attr["type"] == 0) & (_neigh_sum(src, n) > 0))
    for f in ("type", "bucket_index", "bnd"):
        attr[f][new] = latent[f][new]
    attr["has_involved"][new] = 0
    attr["flags"][new] = 0
    attr["active"][new] = 1                                  # (Bucket ctor: active until update_bgmesh says otherwise)
    return new


def update_bgmesh(attr, rows_per_cell, ncell):
    """update_bgmesh (src/datstr/mesh_update.cc:32-245): a bucket is active when one of its neighbour buckets holds a
    particle (rows_per_cell: plist sizes over the whole lattice, all ranks).  attr["active"] is rewritten."""
    n = [int(x) for x in ncell]
    occ = (attr["type"] != 0) & (np.asarray(rows_per_cell) > 0)
    attr["active"] = ((attr["type"] != 0) & (_neigh_sum(occ, n) > 0)).astype(np.int32)
    return attr


def wall_ghost_images(rows, mindom, ncell, d, attr, len_scale=1.0):
    """search_bnd_images (src/datstr/bnd_image.cc:27-270) for a FLAT ground, in closed form: every wall ghost (bc_type 2,
    not a guest) of an on-ground MIXED bucket (bucket_index[4] == -1, [5] == 0) gets the mirror point across the ground
    plane z = bnd[4] (Bucket::calc_intersection_ground, bucket.cc:320-354; refc = 2 intsct - coord, bnd_image.cc),
    and the bucket that contains it, found from the ghost's own bucket by the reference's walk (a point ON a face
    of the current bucket stays there).  -> IMAGE_DTYPE array (ghost_key, cell, coord) in row order, and the rows.
    Ghosts in UNDERGROUND buckets get no image, as in the reference."""
    n = np.asarray(ncell, np.int64)
    x = np.asarray(rows["coords"], np.float64)
    lo0 = np.asarray(mindom, np.float64)
    ci = np.floor((x * len_scale - lo0) / d).astype(np.int64)
    ci = np.clip(ci, 0, n - 1)
    for k in range(3):                                       # membership [mincrd, maxcrd) against the stored bounds
        ci[:, k] -= x[:, k] < (lo0[k] + ci[:, k] * d) / len_scale
        ci[:, k] += x[:, k] >= (lo0[k] + (ci[:, k] + 1) * d) / len_scale
    lin = (ci[:, 0] * n[1] + ci[:, 1]) * n[2] + ci[:, 2]
    a = attr[lin]
    sel = np.flatnonzero((rows["bc_type"] == 2) & (rows["guest"] == 0) & (a["type"] == MIXED) &
                         (a["bucket_index"][:, 4] == -1) & (a["bucket_index"][:, 5] == 0))
    xs, cs = x[sel], ci[sel].copy()
    ref = np.stack([2 * xs[:, 0] - xs[:, 0], 2 * xs[:, 1] - xs[:, 1], 2 * a["bnd"][sel, 4] - xs[:, 2]], axis=1)
    for _ in range(int(n.max()) + 1):                        # the walk: left if below mincrd, right if above maxcrd
        lo = (lo0 + cs * d) / len_scale
        hi = (lo0 + (cs + 1) * d) / len_scale
        step = (ref > hi).astype(np.int64) - (ref < lo).astype(np.int64)
        if not step.any():
            break
        cs += step
    img = np.zeros(len(sel), _l.IMAGE_DTYPE)
    img["ghost_key"] = rows["key"][sel]
    img["cell"] = (cs[:, 0] * n[1] + cs[:, 1]) * n[2] + cs[:, 2]
    img["coord"] = ref
    return img, sel
