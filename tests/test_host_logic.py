"""CPU (-m "not gpu"): host-side logic of the drop-in layer — TimeProps (properties.h:20-150),
simulation.data parsing (dataread.cc:109-235), synthetic case generator."""
import os
import numpy as np

from plume_sph_b200.driver import TimeProps, stof
from plume_sph_b200 import cases


def test_timeprops_clips_to_output_and_end_times():
    tp = TimeProps(max_time=0.25, max_steps=100, timeoutput=0.1)
    dts = []
    while not tp.ifend():
        dts.append(tp.incrtime(0.03))
        tp.ifoutput()
    t = np.cumsum(dts)
    assert any(abs(t - 0.1) < 1e-15) and any(abs(t - 0.2) < 1e-15) and abs(t[-1] - 0.25) < 1e-15
    assert max(dts) <= 0.03 and min(dts) > 0
    assert tp.ioutput == 3  # 0.1, 0.2 and the end time


def test_timeprops_erupt_window_and_step_limit():
    tp = TimeProps(max_time=10, max_steps=3, timeoutput=100, stat_erupt=0.05, end_erupt=0.1)
    assert not tp.iferupt()
    tp.incrtime(0.06)
    assert tp.iferupt()
    tp.incrtime(0.06)
    assert not tp.iferupt()
    tp.incrtime(0.01); tp.incrtime(0.01)
    assert tp.ifend()  # step 4 > max_steps 3 (properties.h:92-95)


def test_stof_rounds_to_float():
    assert stof("0.17") == float(np.float32(0.17)) != 0.17
    assert stof("100") == 100.0


def test_cube_case_shape_and_eos():
    cfg, rows = cases.cube(8, seed=5)
    assert len(rows) == 24 ** 3
    real = rows["bc_type"] == 100
    assert real.sum() == 512 and (rows["bc_type"][~real] == 1).all()
    assert (rows["phase_num"][real] == 2).sum() > 0
    # pressure equilibrium between air and erupted material, EOS-consistent secondaries
    p = rows["pressure"]
    assert np.allclose(p, p[0], rtol=1e-12)
    assert len(np.unique(rows["key"], axis=0)) == len(rows)
    # every particle lies inside the mesh with at least one spare bucket around it
    lo = np.array([cfg.mindom[k] for k in range(3)])
    hi = lo + np.array([cfg.ncell[k] for k in range(3)]) * cfg.delta
    assert (rows["coords"] > lo + cfg.delta - 1e-9).all() and (rows["coords"] < hi - cfg.delta + 1e-9).all()
