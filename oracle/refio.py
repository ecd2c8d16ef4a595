"""Readers for the files written by the hooked reference build (oracle/ref_build/ref_hooks.cc)
and by the flat-record GH5 stand-in (oracle/ref_build/gh5_flat.cc).

Test infrastructure only: imported by tests/, tools under oracle/, and bench.py's
cpu_baseline / --impl reference legs. Never by the product package.
"""
import struct
import numpy as np

DUMP_ROW = np.dtype([
    ("key", "<u4", 3), ("bc_type", "<i4"), ("phase", "<i4"), ("involved", "<i4"), ("guest", "<i4"),
    ("pad", "<i4"),
    ("mass", "<f8"), ("smlen", "<f8"), ("mass_frac", "<f8"), ("pressure", "<f8"), ("gamma", "<f8"),
    ("sound_speed", "<f8"), ("temperature", "<f8"), ("specific_heat_p", "<f8"),
    ("phase_density", "<f8", 2),
    ("coord", "<f8", 3), ("vel", "<f8", 3), ("smoothed_v", "<f8", 3),
    ("rho", "<f8"), ("engr", "<f8"), ("smoothed_e", "<f8"),
    ("grad", "<f8", 15),
])
assert DUMP_ROW.itemsize == 328


def read_dump(path):
    """-> (header dict, structured array of DUMP_ROW in THashTable iteration order)."""
    with open(path, "rb") as f:
        magic, dim, gsph = struct.unpack("<iii", f.read(12))
        assert magic == 0x50535048, path
        n, = struct.unpack("<q", f.read(8))
        time, dt = struct.unpack("<dd", f.read(16))
        rows = np.fromfile(f, dtype=DUMP_ROW, count=n)
    return dict(dim=dim, gsph=gsph, n=n, time=time, dt=dt), rows


def read_neigh(path):
    """-> (offsets int64[n+1], idx int32[...]) CSR neighbour lists (row indices of the same step)."""
    with open(path, "rb") as f:
        n, = struct.unpack("<q", f.read(8))
        off = np.fromfile(f, dtype="<i8", count=n + 1)
        idx = np.fromfile(f, dtype="<i4", count=int(off[-1]))
    return off, idx


def read_images(path):
    """-> structured array of the wall-ghost images (BndImage): procs, bucket key, ghost key, reflection point"""
    dt = np.dtype([("buckproc", "<i4"), ("partproc", "<i4"), ("bucket_key", "<u4", 2), ("ghost_key", "<u4", 3),
                   ("coord", "<f8", 3)])
    with open(path, "rb") as f:
        n, = struct.unpack("<q", f.read(8))
        return np.fromfile(f, dtype=dt, count=n)


def read_buckets(path):
    """-> list of dicts (key, mincrd, maxcrd, type, has_involved, active, guest, bucket_index, bnd, rows)."""
    out = []
    with open(path, "rb") as f:
        nb, = struct.unpack("<q", f.read(8))
        for _ in range(nb):
            key = struct.unpack("<II", f.read(8))
            mn = struct.unpack("<ddd", f.read(24))
            mx = struct.unpack("<ddd", f.read(24))
            typ, inv, act, guest, npart = struct.unpack("<iiiii", f.read(20))
            bidx = struct.unpack("<6i", f.read(24))
            bnd = struct.unpack("<6d", f.read(48))
            rows = np.fromfile(f, dtype="<i4", count=npart)
            out.append(dict(key=key, mincrd=mn, maxcrd=mx, type=typ, has_involved=inv,
                            active=act, guest=guest, bucket_index=bidx, bnd=bnd, rows=rows))
    return out


def read_bucketsx(path):
    """extended bucket file of the domain-growth hooks -> list of dicts: read_buckets' fields + particles_type (bit 0 real,
    1 wall ghost, 2 pressure ghost, 3 erupt ghost), erupt_flag, neigh_proc[27]"""
    out = []
    with open(path, "rb") as f:
        nb, = struct.unpack("<q", f.read(8))
        for _ in range(nb):
            key = struct.unpack("<II", f.read(8))
            mn = struct.unpack("<ddd", f.read(24))
            mx = struct.unpack("<ddd", f.read(24))
            typ, inv, act, guest, npart, ptype, erupt = struct.unpack("<7i", f.read(28))
            bidx = struct.unpack("<6i", f.read(24))
            bnd = struct.unpack("<6d", f.read(48))
            nproc = struct.unpack("<27i", f.read(108))
            rows = np.fromfile(f, dtype="<i4", count=npart)
            out.append(dict(key=key, mincrd=mn, maxcrd=mx, type=typ, has_involved=inv, active=act, guest=guest,
                            particles_type=ptype, erupt_flag=erupt, bucket_index=bidx, bnd=bnd, neigh_proc=nproc, rows=rows))
    return out


def read_grow_params(path):
    """what add_pressure_ghost reads besides the tables (hook dump): step, constants, key box, Meteo table"""
    with open(path, "rb") as f:
        step, nr, npp, atm = struct.unpack("<4i", f.read(16))
        mass, smlen, gamma = struct.unpack("<3d", f.read(24))
        mindom = np.fromfile(f, "<f8", 3); maxdom = np.fromfile(f, "<f8", 3)
        Ll = np.fromfile(f, "<f8", 3); Lu = np.fromfile(f, "<f8", 3)
        z = np.fromfile(f, "<f8", nr)
        data = np.fromfile(f, "<f8", nr * npp).reshape(npp, nr)
    return dict(step=step, atmosphere_type=atm, mass=mass, smlen=smlen, gamma=gamma, mindom=mindom, maxdom=maxdom, Ll=Ll, Lu=Lu,
                meteo_z=z, meteo=data)


_TYPES = {1: "<i4", 2: "<u4", 3: "<f4", 4: "<f8", 5: "S1"}


def read_records(path):
    """Flat record file (pvplotNNN.h5part / funkyNNNN.h5 stand-ins) -> list of (name, array)."""
    out = []
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if len(h) < 4:
                break
            nl, = struct.unpack("<I", h)
            name = f.read(nl).decode()
            typ, = struct.unpack("<i", f.read(4))
            count, = struct.unpack("<Q", f.read(8))
            es, = struct.unpack("<I", f.read(4))
            if typ in _TYPES:
                arr = np.fromfile(f, dtype=_TYPES[typ], count=count)
            else:
                arr = np.frombuffer(f.read(count * es), dtype=np.uint8).reshape(count, es)
            out.append((name, arr))
    return out
