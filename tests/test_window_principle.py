"""CPU: the window technique tests/test_gpu_sizes.py relies on -- a sub-block
of w^3 cells plus 3 layers, cut out after timestep(), evolved on its own,
reproduces the full block's result on its w^3 cells and bounding faces bit for
bit (ghost depth 3 = the reach of one VL+CT step,
EnzoMethodMHDVlct.cpp:124-133)."""
import numpy as np
import pytest

from helpers import make_config, random_state, copy_state, passive_names, oracle

FACE_AXIS = {"bfieldi_x": 0, "bfieldi_y": 1, "bfieldi_z": 2}


@pytest.mark.parametrize("kw", [
    dict(riemann="hlld", recon="plm", mhd=True),
    dict(riemann="hllc", recon="plm", mhd=False, dual_energy=True, gamma=1.4),
    dict(riemann="hlld", recon="plm_athena", mhd=True, dual_energy=True, n_passive=2),
], ids=["mhd", "hydro_de", "mhd_de_scalars"])
def test_window_reproduces_full_block(kw):
    cfg = make_config(**kw)
    n, g, d = (24, 20, 18), (3, 3, 3), (0.1, 0.12, 0.09)
    full = random_state(cfg, n, g, seed=4)
    m = oracle.CpuMethod(cfg, g)
    blk = oracle.numpy_block(full, n, g, d, passive_names(cfg))
    dt = m.timestep(blk)
    before = copy_state(full)
    m.compute(blk, dt)
    m.close()
    w, lo = (8, 7, 6), (5, 4, 3)
    win = {}
    for k, v in before.items():
        sl = []
        for ax in (2, 1, 0):
            ext = w[ax] + 2 * g[ax] + (1 if FACE_AXIS.get(k, -1) == ax else 0)
            sl.append(slice(lo[ax], lo[ax] + ext))
        win[k] = v[tuple(sl)].copy()
    m2 = oracle.CpuMethod(cfg, g)
    m2.compute(oracle.numpy_block(win, w, g, d, passive_names(cfg)), dt)
    m2.close()
    for k in win:
        if k == "pressure":
            continue
        slw, slf = [], []
        for ax in (2, 1, 0):
            ext = w[ax] + (1 if FACE_AXIS.get(k, -1) == ax else 0)
            slw.append(slice(g[ax], g[ax] + ext))
            slf.append(slice(lo[ax] + g[ax], lo[ax] + g[ax] + ext))
        assert np.array_equal(win[k][tuple(slw)], full[k][tuple(slf)]), k
