"""The problem initialisers that generate the reference's answer-test problems
directly in device memory (enzo-e_b200/problems.py: inclined linear waves,
shock tubes, the cloud in a wind) against the oracle's C restatement of the reference's Initial
classes -- on the CPU here (torch as the array library), and on the GPU
through whole golden runs."""
import numpy as np
import pytest

import problems as P
from helpers import oracle


@pytest.mark.parametrize("name,mhd", [(n, True) for n in sorted(P.MHD_WAVES)]
                         + [(n, False) for n in sorted(P.HD_WAVES)])
def test_inclined_wave_matches_oracle_ic(name, mhd):
    from enzo_e_b200 import problems as DP
    cfg, f, blk, n, g, d, t_final = P.linear_wave_setup(name, 16, mhd)
    wave_type = (P.MHD_WAVES if mhd else P.HD_WAVES)[name][0]
    got = DP.inclined_wave(n, g, (0.0, 0.0, 0.0), d, wave_type, P.ALPHA, P.BETA,
                           device="cpu", gamma=cfg.gamma, mhd=mhd)
    assert set(got) == set(f)
    for k in f:
        # same formulas, libm vs torch trigonometry: a few 1e-23 on 1e-6 waves
        assert np.max(np.abs(got[k].numpy() - f[k])) < 1e-20, k


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_shock_tubes_match_oracle_ic(axis):
    from enzo_e_b200 import problems as DP
    cfg, f, blk, n, g, d, t_final = P.rj2a_setup(axis)
    got = DP.shock_tube(n, g, (0.0, 0.0, 0.0), d, "rj2a", axis, device="cpu",
                        gamma=cfg.gamma)
    assert all(np.array_equal(got[k].numpy(), f[k]) for k in f)
    cfg, f, blk, n, g, d, t_final = P.sod_de_setup(axis)
    got = DP.shock_tube(n, g, (0.0, 0.0, 0.0), d, "sod", axis, device="cpu",
                        gamma=cfg.gamma, axis_velocity=P.SOD_BKG_VELOCITY,
                        dual_energy=True)
    assert all(np.array_equal(got[k].numpy(), f[k]) for k in f)


@pytest.mark.parametrize("solver", ["hlld", "hllc"])
def test_cloud_matches_oracle_ic(solver):
    """EnzoInitialCloud: the answer test's cloud, bit for bit"""
    from enzo_e_b200 import problems as DP
    cfg, f, blk, n, g, d, t_stop = P.cloud_setup(solver)
    got = DP.cloud(n, g, P.CLOUD_LOWER, d, device="cpu", mhd=cfg.mhd_choice == 1,
                   **P.CLOUD)
    assert set(got) == set(f)
    assert all(np.array_equal(got[k].numpy(), f[k]) for k in f)
    inside = f["density"] == P.CLOUD["cloud_density"]
    outside = f["density"] == P.CLOUD["wind_density"]
    assert inside.sum() > 1000 and 0 < (~inside & ~outside).sum() < inside.sum()


def test_cloud_off_centre_matches_oracle_ic():
    """an off-centre sphere on cells whose width is not a power of two, 8^3
    sub-cells, without dual energy: hundreds of distinct cut-cell fractions"""
    from enzo_e_b200 import problems as DP
    cfg = P.make_config(riemann="hllc", recon="plm", mhd=False)
    n, g, d = (20, 14, 18), (3, 3, 3), (0.11, 0.11, 0.11)
    lower = (-1.0, -0.8, -0.9)
    kw = dict(P.CLOUD, center=(0.13, -0.07, 0.21), cloud_radius=0.61, subsample_n=3,
              wind_internal_energy=0.0)
    f = P.alloc_fields(cfg, n, g)
    oracle.ic_cloud(oracle.numpy_block(f, n, g, d), lower, **kw)
    got = DP.cloud(n, g, lower, d, device="cpu", mhd=False, dual_energy=False, **kw)
    assert set(got) == set(f)
    assert all(np.array_equal(got[k].numpy(), f[k]) for k in f)
    assert np.unique(f["density"]).size > 100


def test_perturbed_cloud_matches_oracle_ic():
    """the optional density perturbation of the cloud (Initial:cloud:perturb_*,
    EnzoInitialCloud.cpp:86-163, 327-390): wave parameters drawn on the host
    bit for bit like the reference's std::minstd_rand sequence, cell and
    sub-cell averages evaluated by torch -- equal to the oracle (itself
    bit-identical to the compiled reference) to the last bits of cos()"""
    from enzo_e_b200 import problems as DP
    cfg = P.make_config(riemann="hllc", recon="plm", mhd=False)
    n, g, d = (20, 14, 18), (3, 3, 3), (0.11, 0.11, 0.11)
    lower = (-1.0, -0.8, -0.9)
    perturb = (5, 20231, 0.12, 0.2, 0.61)
    kw = dict(P.CLOUD, center=(0.13, -0.07, 0.21), cloud_radius=0.61, subsample_n=3,
              wind_internal_energy=0.0)
    f = P.alloc_fields(cfg, n, g)
    oracle.ic_cloud(oracle.numpy_block(f, n, g, d), lower, perturb=perturb, **kw)
    got = DP.cloud(n, g, lower, d, device="cpu", mhd=False, dual_energy=False,
                   perturb=perturb, **kw)
    assert set(got) == set(f)
    for k in f:
        np.testing.assert_allclose(got[k].numpy(), f[k], rtol=2e-14, atol=1e-300,
                                   err_msg=k)
    # the host-side wave table IS bit-identical: same PRNG, same libm
    waves = DP.cloud_perturbation_waves(perturb[0], perturb[1], perturb[3], perturb[4])
    assert len(waves) == 5 and all(0.0 <= w[3] < np.pi for w in waves)
    lam = [2 * np.pi / np.sqrt(w[0] ** 2 + w[1] ** 2 + w[2] ** 2) for w in waves]
    assert all(perturb[3] - 1e-12 <= x <= perturb[4] + 1e-12 for x in lam)
    unperturbed = DP.cloud(n, g, lower, d, device="cpu", mhd=False, dual_energy=False, **kw)
    assert float((got["density"] - unperturbed["density"]).abs().max()) > 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("name,mhd", [("fast", True), ("alfven", True),
                                      ("sound", False)])
def test_linear_wave_golden_from_device_ic(name, mhd):
    """the golden N16 norm with the problem generated on the device: nothing
    but the final snapshot crosses PCIe"""
    import torch
    from enzo_e_b200 import problems as DP
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = P.linear_wave_config(mhd)
    N = 16
    n, g = (2 * N, N, N), (3, 3, 3)
    d = (3.0 / n[0], 1.5 / n[1], 1.5 / n[2])
    wave_type, t_final = (P.MHD_WAVES if mhd else P.HD_WAVES)[name]
    dev = DP.inclined_wave(n, g, (0.0, 0.0, 0.0), d, wave_type, P.ALPHA, P.BETA,
                           device="cuda", gamma=cfg.gamma, mhd=mhd)
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(dev, n, g, d)

    class Run:
        def timestep(self, _b):
            return method.timestep(block)

        def compute(self, _b, dt):
            method.compute(block, dt)
    s0 = P.snapshot(cfg, {k: v.cpu().numpy() for k, v in dev.items()}, g)
    P.evolve(Run(), None, t_final, lambda _b: method.refresh_periodic(block, 7),
             dump_times=(0.0, t_final))
    method.synchronize()
    s1 = P.snapshot(cfg, {k: v.cpu().numpy() for k, v in dev.items()}, g)
    method.close()
    fields = P.LINWAVE_FIELDS_MHD if mhd else P.LINWAVE_FIELDS_HD
    norm = P.l1_error_norm(s0, s1, fields, N)
    golden = (P.GOLDEN_MHD if mhd else P.GOLDEN_HD)[(name, N)]
    assert P.golden_isclose(norm, golden), (norm, golden)
