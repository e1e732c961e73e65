"""GPU parity of the batched entry points: many equally shaped blocks stacked
along z and advanced by ONE set of kernel launches must each end up with the
bits the oracle gives that block alone (ghost zones included), and the batch
timestep must be the minimum of the blocks' timesteps."""
import numpy as np
import pytest

from helpers import (make_config, random_state, copy_state, passive_names,
                     bit_equal, max_abs_diff, oracle)
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu

N, G, D = (12, 10, 8), (3, 3, 3), (0.1, 0.12, 0.09)
BATCH_CASES = ["mhd_hlld_plm", "mhd_hlld_plm_scalars", "mhd_hlld_athena_de",
               "mhd_hlle_nn_de_scalar", "hd_hllc_plm_de_scalars",
               "hd_hllc_athena_euler", "mhd_hlld_gravity_de_eta0"]


def oracle_blocks(cfg, hosts, nsteps):
    """every block alone through the oracle, with the batch-wide minimum dt"""
    fs = [copy_state(h) for h in hosts]
    blks = [oracle.numpy_block(f, N, G, D, passive_names(cfg)) for f in fs]
    m = oracle.CpuMethod(cfg, G)
    dts = []
    for _ in range(nsteps):
        dt = min(m.timestep(b) for b in blks)
        for b in blks:
            m.compute(b, dt)
        dts.append(dt)
    m.close()
    return fs, dts


def check_all(want, got):
    for n, (w, g_) in enumerate(zip(want, got)):
        eq = bit_equal(w, g_)
        bad = {k: max_abs_diff(w, g_)[k] for k, ok in eq.items() if not ok}
        assert not bad, f"block {n}: fields differ from the oracle: {bad}"


@pytest.mark.parametrize("name", BATCH_CASES)
@pytest.mark.parametrize("device_resident,pair", [(True, None), (False, None), (True, 7),
                                                  (True, 0)],
                         ids=["device", "host", "device_pair7", "device_pair0"])
def test_batch_matches_blocks_alone(name, device_resident, pair):
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    nb = 5
    hosts = [random_state(cfg, N, G, seed=100 + n) for n in range(nb)]
    want, dts_want = oracle_blocks(cfg, hosts, 2)
    if device_resident:
        fs = [{k: torch.from_numpy(v.copy()).cuda() for k, v in h.items()}
              for h in hosts]
    else:
        fs = [copy_state(h) for h in hosts]
    method = EnzoMethodMHDVlct(config=cfg)
    if pair is not None:      # stacked pair kernels (option "pair_kernels")
        method.set_option("pair_kernels", pair)
    blocks = [Block(f, N, G, D, passive=passive_names(cfg)) for f in fs]
    launches0 = method.kernel_launches()
    dts = []
    for _ in range(2):
        dt = method.timestep_batch(blocks)
        method.compute_batch(blocks, dt)
        dts.append(dt)
    method.synchronize()
    launches = method.kernel_launches() - launches0
    method.close()
    if device_resident:
        fs = [{k: v.cpu().numpy() for k, v in f.items()} for f in fs]
    assert dts == dts_want
    check_all(want, fs)
    assert all(b.compute_done_count == 2 for b in blocks)
    # one launch sequence for the whole batch, not one per block
    assert launches < 2 * 60


def test_sub_batches_and_single_block_calls_share_a_handle():
    """batch_max_blocks splits a batch; the same handle then serves single
    blocks (scratch grows, never shrinks)"""
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES["mhd_hlld_plm"])
    nb = 7
    hosts = [random_state(cfg, N, G, seed=200 + n) for n in range(nb)]
    want, dts_want = oracle_blocks(cfg, hosts, 1)
    fs = [{k: torch.from_numpy(v.copy()).cuda() for k, v in h.items()} for h in hosts]
    method = EnzoMethodMHDVlct(config=cfg)
    # a single-block call first: the scratch must grow for the batch
    warm = {k: torch.from_numpy(v.copy()).cuda() for k, v in hosts[0].items()}
    wb = Block(warm, N, G, D)
    method.compute(wb, method.timestep(wb))
    method.set_option("batch_max_blocks", 3)
    blocks = [Block(f, N, G, D) for f in fs]
    dt = method.timestep_batch(blocks)
    method.compute_batch(blocks, dt)
    method.synchronize()
    got = [{k: v.cpu().numpy() for k, v in f.items()} for f in fs]
    assert dt == dts_want[0]
    check_all(want, got)
    # ... and a single block again on the same handle
    f1 = {k: torch.from_numpy(v.copy()).cuda() for k, v in hosts[1].items()}
    b1 = Block(f1, N, G, D)
    dt1 = method.timestep(b1)
    method.compute(b1, dt)
    method.synchronize()
    assert all(bit_equal(want[1], {k: v.cpu().numpy() for k, v in f1.items()}).values())
    method.close()


def test_batch_validation():
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block, VlctError
    cfg = make_config(**CASES["mhd_hlld_plm"])
    a = random_state(cfg, N, G, seed=1)
    b = random_state(cfg, (N[0] + 2, N[1], N[2]), G, seed=2)
    method = EnzoMethodMHDVlct(config=cfg)
    ba = Block({k: torch.from_numpy(v).cuda() for k, v in a.items()}, N, G, D)
    bb = Block({k: torch.from_numpy(v).cuda() for k, v in b.items()},
               (N[0] + 2, N[1], N[2]), G, D)
    with pytest.raises(VlctError):
        method.compute_batch([ba, bb], 1e-3)       # shapes differ
    bh = Block(copy_state(a), N, G, D)
    with pytest.raises(VlctError):
        method.compute_batch([ba, bh], 1e-3)       # mem_space differs
    method.close()


@pytest.mark.parametrize("sub", [1, 2, 3])
@pytest.mark.parametrize("pinned", [False, True], ids=["pageable", "pinned"])
def test_host_batch_pipeline_bit_exact(sub, pinned):
    """HOST batches run as a double-buffered pipeline over sub-batches (copies
    of one overlap the kernels of another): same bits for every block. Pinned
    host arrays are gathered / scattered in place over PCIe by one kernel per
    field, pageable ones by one cudaMemcpyAsync per (block, field)."""
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES["mhd_hlld_plm_scalars"])
    nb = 7
    hosts = [random_state(cfg, N, G, seed=300 + n) for n in range(nb)]
    want, dts_want = oracle_blocks(cfg, hosts, 2)
    if pinned:
        keep = [{k: torch.from_numpy(v.copy()).pin_memory() for k, v in h.items()}
                for h in hosts]
        fs = [{k: t.numpy() for k, t in f.items()} for f in keep]
    else:
        fs = [copy_state(h) for h in hosts]
    method = EnzoMethodMHDVlct(config=cfg)
    method.set_option("host_batch_blocks", sub)
    blocks = [Block(f, N, G, D, passive=passive_names(cfg)) for f in fs]
    dts = []
    for _ in range(2):
        dt = method.timestep_batch(blocks)
        method.compute_batch(blocks, dt)
        dts.append(dt)
    method.close()
    assert dts == dts_want
    check_all(want, fs)


def test_registered_host_memory():
    """vlct_host_register: plain numpy arrays page-locked through the C ABI take
    the zero-copy path of the batched entry points and the asynchronous path of
    the single-block pipeline; after vlct_host_unregister they are pageable
    again and everything still works"""
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block, VlctError
    cfg = make_config(**CASES["mhd_hlld_plm"])
    nb = 4
    hosts = [random_state(cfg, N, G, seed=400 + n) for n in range(nb)]
    want, dts_want = oracle_blocks(cfg, hosts, 3)
    fs = [copy_state(h) for h in hosts]
    method = EnzoMethodMHDVlct(config=cfg)
    for f in fs:
        for a in f.values():
            method.host_register(a)
    blocks = [Block(f, N, G, D) for f in fs]
    dts = []
    for step in range(3):
        if step == 2:            # back to pageable memory for the last step
            for f in fs:
                for a in f.values():
                    method.host_unregister(a)
            with pytest.raises(VlctError):
                method.host_unregister(fs[0]["density"])
        if step == 1:            # block by block through the HOST pipeline
            dt = min(method.timestep(b) for b in blocks)
            for b in blocks:
                method.compute(b, dt)
        else:
            dt = method.timestep_batch(blocks)
            method.compute_batch(blocks, dt)
        dts.append(dt)
    method.close()
    assert dts == dts_want
    check_all(want, fs)
