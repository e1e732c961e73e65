"""GPU parity of the flux-correction output (vlct_save_face_fluxes,
EnzoMethodMHDVlct::save_fluxes_for_corrections_): dt/dx * final-stage fluxes
through the six faces of a block, bit for bit against the oracle's
restatement of hydro-mhd/EnzoMethodMHDVlct.cpp:250-330."""
import numpy as np
import pytest

from helpers import (make_config, random_state, copy_state, passive_names,
                     oracle)
from test_gpu_parity import CASES

pytestmark = pytest.mark.gpu

N, G, D = (12, 9, 7), (3, 3, 3), (0.1, 0.12, 0.09)


@pytest.mark.parametrize("name", ["hd_hllc_plm", "hd_hllc_plm_de_scalars",
                                  "hd_hllc_athena_euler"])
@pytest.mark.parametrize("where", ["device", "host", "host_pipelined"])
def test_face_fluxes_match_oracle(name, where):
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(**CASES[name])
    nf = 6 + cfg.n_passive
    host = random_state(cfg, N, G, seed=51)
    f = copy_state(host)
    blk = oracle.numpy_block(f, N, G, D, passive_names(cfg))
    cpu = oracle.CpuMethod(cfg, G)
    dt = cpu.timestep(blk)
    cpu.compute(blk, dt)
    want = cpu.face_fluxes(blk, dt, N, nf)
    cpu.close()

    method = EnzoMethodMHDVlct(config=cfg)
    if cfg.n_passive:
        # the scalars' face fluxes only exist with their flux arrays
        method.set_option("scalar_flux_arrays", 1)
    if where == "device":
        fields = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    else:
        fields = copy_state(host)
        if where == "host_pipelined":
            method.set_option("host_pipeline_levels", 3)
    block = Block(fields, N, G, D, passive=passive_names(cfg))
    assert method.timestep(block) == dt
    method.compute(block, dt)
    got = method.save_face_fluxes(block, device="cuda" if where == "device" else None)
    method.synchronize()
    method.close()
    assert set(got) == set(want)
    for key in want:
        a = got[key].cpu().numpy() if where == "device" else got[key]
        assert a.shape == want[key].shape
        assert np.array_equal(a.view(np.uint64), want[key].view(np.uint64)), key
        assert np.any(want[key] != 0.0)


def test_face_fluxes_need_hydro_and_a_compute():
    import torch
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block, VlctError
    cfg = make_config(**CASES["hd_hllc_plm"])
    host = random_state(cfg, N, G, seed=52)
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block({k: torch.from_numpy(v).cuda() for k, v in host.items()}, N, G, D)
    with pytest.raises(VlctError):          # nothing computed yet
        method.save_face_fluxes(block)
    method.close()
    cfg = make_config(**CASES["mhd_hlld_plm"])
    host = random_state(cfg, N, G, seed=53)
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block({k: torch.from_numpy(v).cuda() for k, v in host.items()}, N, G, D)
    method.compute(block, method.timestep(block))
    with pytest.raises(VlctError):          # "only supported in hydro-mode"
        method.save_face_fluxes(block)
    method.close()


def test_scalar_face_fluxes_need_their_flux_arrays():
    """by default passive-scalar fluxes are never stored (the update kernel
    forms them itself); asking for them without the option is refused"""
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    from enzo_e_b200.lib import VlctError
    cfg = make_config(riemann="hllc", recon="plm", mhd=False, n_passive=2)
    n, g, d = (10, 8, 8), (3, 3, 3), (0.1, 0.1, 0.1)
    host = random_state(cfg, n, g, seed=3)
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(host, n, g, d, passive=passive_names(cfg))
    method.compute(block, 1e-3)
    with pytest.raises(VlctError, match="scalar_flux_arrays"):
        method.save_face_fluxes(block)
    hydro_only = method.save_face_fluxes(block, n_fields=6)    # fine without
    assert len(hydro_only) > 0
    method.close()
