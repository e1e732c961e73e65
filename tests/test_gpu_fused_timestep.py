"""GPU: vlct_compute_and_timestep = vlct_compute followed by vlct_timestep on
the same block, bit for bit (fields, "pressure", dt), for HOST blocks (one
shot and as a z-pass staging pipeline) and DEVICE blocks."""
import numpy as np
import pytest

from helpers import (make_config, random_state, copy_state, passive_names,
                     bit_equal, max_abs_diff, oracle)

pytestmark = pytest.mark.gpu

CASES = {
    "mhd_hlld_plm": dict(riemann="hlld", recon="plm", theta=1.5, mhd=True),
    "mhd_hlld_athena_de_scalars": dict(riemann="hlld", recon="plm_athena", mhd=True,
                                       dual_energy=True, n_passive=2),
    "hd_hllc_plm_de": dict(riemann="hllc", recon="plm", mhd=False,
                           dual_energy=True, gamma=1.4),
    "hd_hllc_euler": dict(riemann="hllc", recon="plm", mhd=False,
                          time_scheme="euler", courant=0.5),
    "mhd_hlld_floors_de": dict(riemann="hlld", recon="plm", mhd=True,
                               dual_energy=True, dfloor=0.95, pfloor=0.55),
}


def reference_run(cfg, host, n, g, d, nsteps):
    """the oracle driven the way Enzo-E's cycle drives the Method: timestep,
    compute, timestep, compute, ... and one last timestep"""
    f = copy_state(host)
    blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
    m = oracle.CpuMethod(cfg, g)
    dts = [m.timestep(blk)]
    for _ in range(nsteps):
        m.compute(blk, dts[-1])
        dts.append(m.timestep(blk))
    m.close()
    return f, dts


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("where", ["host", "host_pipelined", "device"])
def test_fused_call_equals_compute_then_timestep(name, where):
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    n, g, d = (20, 12, 14), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=9)
    nsteps = 3
    want, dts_want = reference_run(cfg, host, n, g, d, nsteps)
    method = EnzoMethodMHDVlct(config=cfg)
    if where == "host_pipelined":
        method.set_option("host_pipeline_levels", 3)
    if where == "device":
        f = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    else:
        f = copy_state(host)
    block = Block(f, n, g, d, passive=passive_names(cfg))
    dts = [method.timestep(block)]
    for _ in range(nsteps):
        dts.append(method.compute_and_timestep(block, dts[-1]))
    method.synchronize()
    h2d, d2h = method.staged_bytes()
    method.close()
    assert dts == dts_want
    got = {k: v.cpu().numpy() for k, v in f.items()} if where == "device" else f
    eq = bit_equal(want, got)          # "pressure" included
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, bad
    if where != "device":
        # one upload of compute's inputs and one download of its outputs +
        # pressure per cycle (plus the first timestep's own traffic)
        cells = np.prod([n[a] + 2 * g[a] for a in range(3)]) * 8
        assert d2h < (len(want) + 2) * cells * nsteps + 4 * cells


def test_fused_call_needs_pressure():
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    from enzo_e_b200.lib import VlctError
    cfg = make_config(riemann="hllc", recon="plm", mhd=False)
    n, g, d = (8, 8, 8), (3, 3, 3), (0.1, 0.1, 0.1)
    host = random_state(cfg, n, g, seed=1)
    host.pop("pressure")
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(host, n, g, d)
    with pytest.raises(VlctError, match="pressure"):
        method.compute_and_timestep(block, 1e-3)
    method.close()


def test_mirror_grows_with_the_field_set_and_checks_shape():
    """timestep first (no face fields), then compute on the same
    handle: the device mirror must gain the face fields instead of handing
    NULL pointers to the kernels; a differently shaped HOST block is refused."""
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    from enzo_e_b200.lib import VlctError
    cfg = make_config(riemann="hlld", recon="plm", mhd=True)
    n, g, d = (12, 10, 8), (3, 3, 3), (0.1, 0.1, 0.1)
    host = random_state(cfg, n, g, seed=2)
    want, dts_want = reference_run(cfg, host, n, g, d, 1)
    method = EnzoMethodMHDVlct(config=cfg)
    cells_only = {k: v for k, v in host.items() if not k.startswith("bfieldi")}
    dt0 = method.timestep(Block(cells_only, n, g, d))
    assert dt0 == dts_want[0]
    no_pressure = {k: v for k, v in host.items() if k != "pressure"}
    method.compute(Block(no_pressure, n, g, d), dt0)
    dt1 = method.timestep(Block(host, n, g, d))
    assert dt1 == dts_want[1]
    assert all(bit_equal(want, host).values())
    other = random_state(cfg, (10, 10, 8), g, seed=3)
    with pytest.raises(VlctError, match="share one shape"):
        method.timestep(Block(other, (10, 10, 8), g, d))
    method.close()


@pytest.mark.parametrize("name", ["mhd_hlld_plm", "mhd_hlld_athena_de_scalars",
                                  "hd_hllc_plm_de", "mhd_hlld_floors_de"])
@pytest.mark.parametrize("how", ["whole", "parts", "passes"])
def test_device_resident_fold_equals_compute_then_timestep(name, how):
    """vlct_compute_and_timestep_dev / _dev_part: dt in, next dt out, both on
    the device; the CFL work rides on the last update kernel (+ a ghost-shell
    launch). Fields, "pressure" and every dt equal the oracle's
    compute() -> timestep() sequence bit for bit -- as one step, as the three
    parts of the overlapped step, and as z passes."""
    import torch
    from enzo_e_b200 import abi
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    n, g, d = (20, 12, 26), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=10)
    nsteps = 3
    want, dts_want = reference_run(cfg, host, n, g, d, nsteps)
    method = EnzoMethodMHDVlct(config=cfg)
    if how == "passes":
        method.set_option("device_pipeline_levels", 5)
    f = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    block = Block(f, n, g, d, passive=passive_names(cfg))
    dt = method.timestep_dev(block)
    dts = [dt.clone()]
    zr = (10, 17)
    for _ in range(nsteps):
        nxt = torch.empty_like(dt)
        if how == "parts":
            for part in (abi.PART_INTERIOR, abi.PART_LOWER, abi.PART_UPPER):
                method.compute_part(block, dt, part, *zr, dt_next=nxt)
        else:
            method.compute_and_timestep_dev(block, dt, out=nxt)
        dt = nxt
        dts.append(dt.clone())
    method.synchronize()
    torch.cuda.synchronize()
    assert [float(t.item()) for t in dts] == dts_want
    got = {k: v.cpu().numpy() for k, v in f.items()}
    method.close()
    eq = bit_equal(want, got)
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, bad


def test_fold_may_write_dt_in_place():
    """dt_next_device may alias dt_device"""
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES["mhd_hlld_plm"])
    n, g, d = (16, 12, 10), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=12)
    want, dts_want = reference_run(cfg, host, n, g, d, 2)
    method = EnzoMethodMHDVlct(config=cfg)
    f = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    block = Block(f, n, g, d)
    dt = method.timestep_dev(block)
    seen = [float(dt.item())]
    for _ in range(2):
        method.compute_and_timestep_dev(block, dt, out=dt)
        method.synchronize()
        seen.append(float(dt.item()))
    assert seen == dts_want
    got = {k: v.cpu().numpy() for k, v in f.items()}
    method.close()
    assert all(bit_equal(want, got).values())
