"""GPU parity of the z-pass machinery: a step executed as several z-restricted
passes (HOST staging pipeline, DEVICE test hook, the three parts used to
overlap the ghost exchange) must leave exactly the bits of the one-shot step,
i.e. of the CPU oracle -- every field, ghost zones included, and dt."""
import numpy as np
import pytest

from helpers import (make_config, random_state, copy_state, passive_names,
                     bit_equal, max_abs_diff)
from test_gpu_parity import CASES, run_cpu

pytestmark = pytest.mark.gpu

N, G, D = (14, 10, 24), (3, 3, 3), (0.1, 0.12, 0.09)
PASS_CASES = ["mhd_hlld_plm", "mhd_hlld_plm_scalars", "mhd_hlld_athena_de",
              "mhd_hlle_nn_de_scalar", "hd_hllc_plm_de_scalars",
              "mhd_hlld_gravity_de_eta0", "mhd_hlld_nn"]


def check(want, got):
    eq = bit_equal(want, got)
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, f"fields differ from the oracle: {bad}"


@pytest.mark.parametrize("name", PASS_CASES)
@pytest.mark.parametrize("levels", [1, 3, 8, 29])
def test_device_passes_bit_exact(name, levels):
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    host = random_state(cfg, N, G, seed=21)
    want, dts_want = run_cpu(cfg, host, N, G, D, 2)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    method.set_option("device_pipeline_levels", levels)
    block = Block(dev, N, G, D, passive=passive_names(cfg))
    dts = []
    for _ in range(2):
        dt = method.timestep(block)
        method.compute(block, dt)
        dts.append(dt)
    method.synchronize()
    got = {k: v.cpu().numpy() for k, v in dev.items()}
    method.close()
    assert dts == dts_want
    check(want, got)


@pytest.mark.parametrize("name", PASS_CASES)
@pytest.mark.parametrize("levels", [1, 4, 7, 30])
def test_host_pipeline_bit_exact(name, levels):
    """HOST blocks staged as a z pipeline (H2D / kernels / D2H overlapped),
    for compute and for timestep."""
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    host = random_state(cfg, N, G, seed=22)
    want, dts_want = run_cpu(cfg, host, N, G, D, 2)
    f = copy_state(host)
    method = EnzoMethodMHDVlct(config=cfg)
    method.set_option("host_pipeline_levels", levels)
    block = Block(f, N, G, D, passive=passive_names(cfg))
    dts = []
    for _ in range(2):
        dt = method.timestep(block)
        method.compute(block, dt)
        dts.append(dt)
    h2d, d2h = method.staged_bytes()
    method.close()
    assert dts == dts_want
    check(want, f)
    # every input level goes up exactly once per call, every output level
    # comes down exactly once
    cells = np.prod(f["density"].shape) * 8
    assert h2d > 0 and d2h > 0 and h2d % 8 == 0 and d2h >= 2 * cells


def test_host_pipeline_euler_is_one_shot():
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES["hd_hllc_athena_euler"])
    host = random_state(cfg, N, G, seed=23)
    want, dts_want = run_cpu(cfg, host, N, G, D, 2)
    f = copy_state(host)
    method = EnzoMethodMHDVlct(config=cfg)
    method.set_option("host_pipeline_levels", 4)
    block = Block(f, N, G, D)
    dts = []
    for _ in range(2):
        dt = method.timestep(block)
        method.compute(block, dt)
        dts.append(dt)
    method.close()
    assert dts == dts_want
    check(want, f)


@pytest.mark.parametrize("name", PASS_CASES)
@pytest.mark.parametrize("zr", [(8, 9), (8, 21), (10, 15), (12, 13)])
@pytest.mark.parametrize("upper_first", [False, True])
def test_three_parts_bit_exact(name, zr, upper_first):
    """interior, then lower / upper in either order == the whole step"""
    import torch
    from enzo_e_b200 import abi
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    host = random_state(cfg, N, G, seed=24)
    want, dts_want = run_cpu(cfg, host, N, G, D, 2)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(dev, N, G, D, passive=passive_names(cfg))
    dts = []
    order = [abi.PART_UPPER, abi.PART_LOWER] if upper_first \
        else [abi.PART_LOWER, abi.PART_UPPER]
    for _ in range(2):
        dt = method.timestep_dev(block)
        method.compute_part(block, dt, abi.PART_INTERIOR, *zr)
        for part in order:
            method.compute_part(block, dt, part, *zr)
        dts.append(dt)
    method.synchronize()
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in dev.items()}
    method.close()
    assert [float(t.item()) for t in dts] == dts_want
    check(want, got)


def test_interior_part_does_not_read_z_ghosts():
    """Poison the z ghost levels while the interior part runs, restore them
    before the lower / upper parts: the result must not change."""
    import torch
    from enzo_e_b200 import abi
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES["mhd_hlld_plm_scalars"])
    host = random_state(cfg, N, G, seed=25)
    want, dts_want = run_cpu(cfg, host, N, G, D, 1)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(dev, N, G, D, passive=passive_names(cfg))
    dt = method.timestep_dev(block)
    gz = G[2]
    saved = {}
    for k, v in dev.items():
        if k == "pressure":
            continue
        # a face-centred-in-z field has one more level; its boundary faces
        # (index gz and mz-gz) belong to the active zone and stay
        top = v.shape[0] - gz
        saved[k] = (v[:gz].clone(), v[top:].clone())
        v[:gz] = float("nan")
        v[top:] = float("nan")
    zr = (gz + abi.PART_REACH_BELOW, N[2] + 2 * gz - gz - abi.PART_REACH_ABOVE)
    method.compute_part(block, dt, abi.PART_INTERIOR, *zr)
    method.synchronize()
    torch.cuda.synchronize()
    for k, (lo, hi) in saved.items():
        v = dev[k]
        v[:gz] = lo
        v[v.shape[0] - gz:] = hi
    method.compute_part(block, dt, abi.PART_LOWER, *zr)
    method.compute_part(block, dt, abi.PART_UPPER, *zr)
    method.synchronize()
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in dev.items()}
    method.close()
    assert float(dt.item()) == dts_want[0]
    check(want, got)


def test_part_arguments_are_validated():
    import torch
    from enzo_e_b200 import abi
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block, VlctError
    cfg = make_config(**CASES["mhd_hlld_plm"])
    host = random_state(cfg, N, G, seed=26)
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(dev, N, G, D)
    dt = method.timestep_dev(block)
    for zr in [(6, 20), (7, 22), (15, 15)]:
        with pytest.raises(VlctError):
            method.compute_part(block, dt, abi.PART_INTERIOR, *zr)
    with pytest.raises(VlctError):
        method.set_option("no_such_option", 1)
    method.close()


@pytest.mark.parametrize("name", ["mhd_hlld_plm", "mhd_hlld_athena_de",
                                  "hd_hllc_plm_de_scalars"])
@pytest.mark.parametrize("levels", [0, 4])
def test_host_mirror_reuse_bit_exact(name, levels):
    """option host_mirror_reuse: the timestep that follows a compute of the
    same HOST block reads the device copy instead of uploading again -- same
    bits, fewer bytes; another block in between falls back to uploading"""
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    host = random_state(cfg, N, G, seed=27)
    other = random_state(cfg, N, G, seed=28)
    want, dts_want = run_cpu(cfg, host, N, G, D, 3)
    want_other, dts_other = run_cpu(cfg, other, N, G, D, 1)
    runs = {}
    for reuse in (0, 1):
        f, fo = copy_state(host), copy_state(other)
        method = EnzoMethodMHDVlct(config=cfg)
        method.set_option("host_pipeline_levels", levels)
        method.set_option("host_mirror_reuse", reuse)
        block = Block(f, N, G, D, passive=passive_names(cfg))
        block_o = Block(fo, N, G, D, passive=passive_names(cfg))
        dts = []
        for step in range(3):
            dt = method.timestep(block)
            method.compute(block, dt)
            dts.append(dt)
            if step == 1:
                # a different block right after a compute: must not be served
                # from the mirror of `block`
                dto = method.timestep(block_o)
                method.compute(block_o, dto)
                assert dto == dts_other[0]
        runs[reuse] = method.staged_bytes()
        method.close()
        assert dts == dts_want
        check(want, f)
        check(want_other, fo)
    # with reuse one of the three timesteps of `block` (the one after its own
    # compute, step 0 -> 1) skips its upload
    assert runs[1][0] < runs[0][0]
    assert runs[1][1] == runs[0][1]
