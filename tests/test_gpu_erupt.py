"""-m gpu: psph_add_new_erupt (CUDA, through the C ABI) against the particles the unmodified reference's add_new_erupt
added at the first injecting time steps of the small plume and of the JPUE jet (tests/golden/erupt3d*.npz;
setup_erupt.cc:390-670) and against the
oracle restatement.  Bar: bit-exact — birth keys (Hilbert index + add_step), coordinates, velocities and every other
field of the new rows; the rows that were there before are untouched; the table stays consistent (a time step runs)."""
import numpy as np
import pytest

import oracle_py as op
from common import golden, to_product_cfg, dump_to_particles, cell_attrs, D2P
import plume_sph_b200 as psph
from plume_sph_b200.solver import ORDER_DEVICE

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,base,steps", [("erupt3d", "plume3d_small", (2, 3)), ("erupt3d_jpue", "jpue3d", (2, 3, 4))])
def test_add_new_erupt_matches_reference(name, base, steps):
    z = golden(name)
    zp = golden(base)
    ocfg = op.config_for(base, zp)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        attr, _ = cell_attrs(ocfg, zp)
        s.set_cells(attr)
        pre = dump_to_particles(zp["pre"])
        s.upload(pre)
        nrows = len(pre)
        seen = {tuple(int(x) for x in k) for k in pre["key"]}
        for st in steps:
            er, d = op.erupt_from_params(z["param_names"], z["s%d_params" % st])
            boxes = np.concatenate([z["s%d_cell_min" % st], z["s%d_cell_max" % st]], axis=1)
            s.set_erupt(er, boxes)
            added = s.add_new_erupt(d["time"], d["dt"], int(d["step"]))
            want = z["s%d_new" % st]
            ora = op.add_new_erupt(ocfg, er, boxes, d["time"], d["dt"], int(d["step"]))
            assert added == len(want) == len(ora)
            out = s.download(ORDER_DEVICE)
            assert len(out) == nrows + added
            new = np.asarray([i for i, k in enumerate(out["key"]) if tuple(int(x) for x in k) not in seen])
            assert len(new) == added
            got = out[new]
            gi = {tuple(int(x) for x in k): i for i, k in enumerate(got["key"])}
            got = got[np.asarray([gi[tuple(int(x) for x in k)] for k in want["key"]])]
            oi = {tuple(int(x) for x in k): i for i, k in enumerate(ora["key"])}
            ora = ora[np.asarray([oi[tuple(int(x) for x in k)] for k in want["key"]])]
            for f, pf in D2P.items():
                if f in want.dtype.names:
                    assert np.array_equal(got[pf], want[f]), (st, f)
                    assert np.array_equal(got[pf], ora[pf]), (st, f, "oracle")
            assert np.array_equal(got["bc_type"], want["bc_type"]) and not got["guest"].any()
            seen |= set(gi)
            nrows += added
        # the rows that were there before are untouched, and the table is consistent: a time step runs
        out = s.download(ORDER_DEVICE)
        oi = {tuple(int(x) for x in k): i for i, k in enumerate(out["key"])}
        old = out[np.asarray([oi[tuple(int(x) for x in k)] for k in pre["key"]])]
        for f in ("coords", "vel", "rho", "engr", "mass", "smlen", "bc_type"):
            assert np.array_equal(old[f], pre[f]), f
        dt, lost, ierr = s.step(0.0, 1)
        assert dt > 0 and ierr == 0
        assert s.n == nrows - lost
    finally:
        s.close()


def test_add_new_erupt_needs_the_vent_record():
    zp = golden("plume3d_small")
    ocfg = op.config_for("plume3d_small", zp)
    s = psph.Solver(to_product_cfg(ocfg))
    try:
        with pytest.raises(psph.PsphError):
            s.add_new_erupt(0.5, 0.2, 2)
    finally:
        s.close()
