"""CPU: pin the oracle.

1. The plain-C restatement (oracle/vlct_oracle.c) and -- where it is available
   -- the reference's own compiled sources (oracle/_ref) reproduce the golden
   L1 error norms that the reference's vlct answer tests hard-code
   (input/vlct/run_MHD_linear_wave_test.py:178-188,
    input/vlct/run_HD_linear_wave_test.py:70-80), with the reference's own
   tolerance (rtol 1e-13, atol 7e-14; input/vlct/testing_utils.py:275-292).
2. The restatement is bit-identical to the compiled reference on seeded random
   states for every solver / reconstructor / dual-energy / scalar combination.
"""
import numpy as np
import pytest

import problems as P
from helpers import (make_config, random_state, copy_state, passive_names,
                     bit_equal, max_abs_diff, oracle)


@pytest.mark.parametrize("name", sorted(P.MHD_WAVES))
def test_mhd_linear_wave_n16_golden(name):
    l1, dts, _ = P.run_linear_wave(name, 16, mhd=True, kind="oracle")
    assert P.golden_isclose(l1, P.GOLDEN_MHD[(name, 16)]), (l1, len(dts))


@pytest.mark.parametrize("name", sorted(P.HD_WAVES))
def test_hd_linear_wave_n16_golden(name):
    l1, dts, _ = P.run_linear_wave(name, 16, mhd=False, kind="oracle")
    assert P.golden_isclose(l1, P.GOLDEN_HD[(name, 16)]), (l1, len(dts))


@pytest.mark.slow
@pytest.mark.parametrize("name", ["fast", "entropy"])
def test_mhd_linear_wave_n32_golden(name):
    l1, _, _ = P.run_linear_wave(name, 32, mhd=True, kind="oracle")
    assert P.golden_isclose(l1, P.GOLDEN_MHD[(name, 32)]), l1


@pytest.mark.slow
def test_hd_linear_wave_n32_golden():
    l1, _, _ = P.run_linear_wave("sound", 32, mhd=False, kind="oracle")
    assert P.golden_isclose(l1, P.GOLDEN_HD[("sound", 32)]), l1


def test_compiled_reference_reproduces_golden():
    """The compiled reference itself, driven by our cycle loop + ICs: pins the
    IC restatement, the refresh and the loop independently of vlct_oracle.c."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    l1, dts_ref, f_ref = P.run_linear_wave("fast", 16, mhd=True, kind="ref")
    assert P.golden_isclose(l1, P.GOLDEN_MHD[("fast", 16)]), l1
    l1o, dts_or, f_or = P.run_linear_wave("fast", 16, mhd=True, kind="oracle")
    assert dts_ref == dts_or
    assert all(bit_equal(f_ref, f_or).values())


RANDOM_CASES = {
    "mhd_hlld_plm": dict(riemann="hlld", recon="plm", theta=1.5, mhd=True),
    "mhd_hlld_plm_theta1": dict(riemann="hlld", recon="plm", theta=1.0, mhd=True),
    "mhd_hlld_athena_de_sc": dict(riemann="hlld", recon="plm_athena", mhd=True,
                                  dual_energy=True, n_passive=2),
    "mhd_hlld_nn": dict(riemann="hlld", recon="nn", mhd=True),
    "mhd_hlle_plm_de": dict(riemann="hlle", recon="plm", mhd=True,
                            dual_energy=True),
    "hd_hllc_euler": dict(riemann="hllc", recon="plm", mhd=False,
                          time_scheme="euler", courant=0.5),
    "hd_hllc_plm": dict(riemann="hllc", recon="plm", mhd=False),
    "hd_hllc_de_sc": dict(riemann="hllc", recon="plm", mhd=False,
                          dual_energy=True, gamma=1.4, n_passive=3),
    "hd_hllc_gravity": dict(riemann="hllc", recon="plm_athena", mhd=False,
                            accel=True),
    "mhd_hlld_gravity_de_eta0": dict(riemann="hlld", recon="plm", mhd=True,
                                     accel=True, dual_energy=True, eta=0.0),
    "mhd_hlld_floors": dict(riemann="hlld", recon="plm", mhd=True,
                            dfloor=0.95, pfloor=0.55),
}


def _run(cfg, host, n, g, d, nsteps, kind):
    f = copy_state(host)
    blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
    m = oracle.CpuMethod(cfg, g, kind=kind)
    dts = []
    for _ in range(nsteps):
        dt = m.timestep(blk)
        m.compute(blk, dt)
        dts.append(dt)
    m.close()
    return f, dts


@pytest.mark.parametrize("name", sorted(RANDOM_CASES))
def test_oracle_bit_identical_to_compiled_reference(name):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    cfg = make_config(**RANDOM_CASES[name])
    n, g, d = (14, 10, 8), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=21)
    want, dts_want = _run(cfg, host, n, g, d, 3, "ref")
    got, dts_got = _run(cfg, host, n, g, d, 3, "oracle")
    assert dts_got == dts_want
    eq = bit_equal(want, got)
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, bad
    # the update did something
    assert not np.array_equal(want["density"], host["density"])


def test_oracle_active_floors_really_trigger():
    """the 'floors' case above must actually exercise the floor branches"""
    cfg = make_config(**RANDOM_CASES["mhd_hlld_floors"])
    n, g, d = (14, 10, 8), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=21)
    got, _ = _run(cfg, host, n, g, d, 1, "oracle")
    act = got["density"][3:-3, 3:-3, 3:-3]
    assert np.min(act) >= 0.95
    assert np.any(act == 0.95)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_rj2a_shock_tube_golden(axis):
    """Ryu-Jones 2a MHD shock tube along x / y / z with outflow boundaries:
    the reference's golden L1 norms against its tabulated solution and exactly
    zero variation across the tube (run_MHD_shock_tube_test.py:62-87)."""
    cfg, f, g, dts = P.run_rj2a(axis)
    snap = P.snapshot(cfg, f, g)
    table = P.load_reference_table("rj2a_shock_tube_t0.2_res256.csv")
    norm = P.table_l1_norm(snap, table, axis, P.RJ2A_FIELDS)
    assert P.golden_isclose(norm, P.GOLDEN_RJ2A[axis]), (norm, P.GOLDEN_RJ2A[axis])
    for k, a in snap.items():       # "standard deviation of the L1 norms == 0.0"
        pencil = np.moveaxis(a, 2 - axis, 0)
        assert np.array_equal(pencil, np.broadcast_to(pencil[:, :1, :1], pencil.shape)), k


def test_boundary_conditions_follow_enzo_boundary():
    """outflow / reflecting ghost fill against a direct numpy statement of
    EnzoBoundary.cpp:164-283,352-466 (cell- and face-centred fields)"""
    cfg = make_config(riemann="hlld", recon="plm", mhd=True)
    n, g, d = (6, 5, 4), (3, 3, 3), (0.1, 0.1, 0.1)
    for kind in ("outflow", "reflecting"):
        for axis in range(3):
            for side in (0, 1):
                f = random_state(cfg, n, g, seed=31)
                want = copy_state(f)
                blk = oracle.numpy_block(f, n, g, d)
                oracle.boundary(blk, axis, side, kind)
                for name, a in want.items():
                    cen = 1 if name == "bfieldi_" + "xyz"[axis] else 0
                    sign = -1.0 if (kind == "reflecting" and name in (
                        "velocity_" + "xyz"[axis], "bfield_" + "xyz"[axis],
                        "bfieldi_" + "xyz"[axis])) else 1.0
                    v = np.moveaxis(a, 2 - axis, 0)      # axis first (a view)
                    na, ga = n[axis], g[axis]
                    for ig in range(ga):
                        if kind == "outflow":
                            src = ga if side == 0 else na + ga - 1 + cen
                            dst = ga - ig - 1 if side == 0 else src + ig + 1
                        else:
                            src = ga + cen + ig if side == 0 else na + ga - 1 - ig
                            dst = ga - ig - 1 if side == 0 else na + ga + ig + cen
                        v[dst] = sign * v[src]
                assert all(bit_equal(want, f).values()), (kind, axis, side)


@pytest.mark.parametrize("kind", ["outflow", "reflecting"])
def test_boundary_restatement_equals_compiled_reference(kind):
    """oracle/vlct_oracle_ic.c:vlct_oracle_boundary against the reference's own
    EnzoBoundary::enforce (compiled unmodified into oracle/_ref): every axis and
    side, cell- and face-centred fields, dual energy, two passive scalars"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    cfg = make_config(riemann="hlld", recon="plm", mhd=True, dual_energy=True,
                      n_passive=2)
    n, g, d = (7, 5, 6), (3, 3, 3), (0.1, 0.1, 0.1)
    names = passive_names(cfg)
    ref = oracle.CpuMethod(cfg, g, kind="ref")
    for axis in range(3):
        for side in (0, 1):
            f = random_state(cfg, n, g, seed=37)
            f_ref = copy_state(f)
            oracle.boundary(oracle.numpy_block(f, n, g, d, names), axis, side, kind,
                            n_passive=2)
            oracle.boundary(oracle.numpy_block(f_ref, n, g, d, names), axis, side,
                            kind, ref_method=ref)
            eq = bit_equal(f, f_ref)
            assert all(eq.values()), (axis, side, [k for k, v in eq.items() if not v])
    ref.close()


def test_inflow_boundary_follows_boundary_value():
    """"inflow" ghost fill against a direct numpy statement of
    BoundaryValue::enforce (Cello/problem_BoundaryValue.cpp:131-273): the g
    outermost layers of the listed fields, boundary face untouched"""
    cfg = make_config(riemann="hlld", recon="plm", mhd=True, dual_energy=True)
    n, g, d = (6, 5, 4), (3, 3, 3), (0.1, 0.1, 0.1)
    values = {"density": 0.25, "velocity_x": 1.5, "internal_energy": 7.0,
              "bfieldi_x": -2.0, "bfieldi_z": 3.0, "bfield_y": 0.0}
    for axis in range(3):
        for side in (0, 1):
            f = random_state(cfg, n, g, seed=32)
            want = copy_state(f)
            blk = oracle.numpy_block(f, n, g, d)
            oracle.boundary_inflow(blk, axis, side, values)
            for name, val in values.items():
                v = np.moveaxis(want[name], 2 - axis, 0)
                if side == 0:
                    v[:g[axis]] = val
                else:
                    v[v.shape[0] - g[axis]:] = val
            assert all(bit_equal(want, f).values()), (axis, side)


@pytest.mark.parametrize("solver", ["hlld", "hllc", "hlle"])
def test_dual_energy_cloud_symmetry(solver):
    """run_dual_energy_cloud_test.py:80-85: a cloud in a Mach-1.5 wind (32^3,
    dual energy, inflow / outflow boundaries) run to t = 0.0625 keeps the
    density symmetric about the wind axis to the reference's tolerance"""
    cfg, f, g, dts = P.run_cloud(solver)
    asym = P.cloud_asymmetries(f, g)
    assert max(asym) <= P.CLOUD_MAX_ASYM[solver], asym
    act = (slice(3, -3),) * 3
    assert f["density"][act].max() > 16.0 and f["density"][act].min() > 0.0
    assert len(dts) > 20
    if oracle.have_ref():          # the reference's own compiled sources agree
        cfg2, f2, g2, dts2 = P.run_cloud(solver, kind="ref")
        assert dts == dts2
        assert all(bit_equal(f, f2).values())


# Initial:cloud:perturb_{Nwaves, seed, amplitude, min_lambda, max_lambda}
CLOUD_PERTURB = (5, 20231, 0.12, 0.2, 0.61)


@pytest.mark.parametrize("case", ["answer_test_hlld", "answer_test_hllc",
                                  "off_centre", "off_centre_perturbed"])
def test_cloud_ic_restatement_equals_compiled_reference(case):
    """oracle/vlct_oracle_ic.c:vlct_ic_cloud against the reference's own
    EnzoInitialCloud::enforce_block (compiled unmodified into oracle/_ref),
    bit for bit, ghost zones included"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    perturb = None
    if case.startswith("off_centre"):
        if case.endswith("perturbed"):    # the optional density perturbation
            perturb = CLOUD_PERTURB       # (EnzoInitialCloud.cpp:86-163, 327-390)
        cfg = make_config(riemann="hllc", recon="plm", mhd=False)
        n, g, d = (20, 14, 18), (3, 3, 3), (0.11, 0.11, 0.11)
        lower = (-1.0, -0.8, -0.9)
        kw = dict(P.CLOUD, center=(0.13, -0.07, 0.21), cloud_radius=0.61,
                  subsample_n=3, wind_internal_energy=0.0)
    else:
        cfg = P.cloud_config(case[-4:])
        n, g, d = (32, 32, 32), (3, 3, 3), (0.125,) * 3
        lower, kw = P.CLOUD_LOWER, P.CLOUD
    f = P.alloc_fields(cfg, n, g)
    oracle.ic_cloud(oracle.numpy_block(f, n, g, d), lower, perturb=perturb, **kw)
    f_ref = P.alloc_fields(cfg, n, g)
    for k in f_ref:                 # everything the initialiser must overwrite
        if not k.startswith("bfield") and k != "pressure":
            f_ref[k][...] = np.nan
    ref = oracle.CpuMethod(cfg, g, kind="ref")
    oracle.ic_cloud(oracle.numpy_block(f_ref, n, g, d), lower, ref_method=ref,
                    perturb=perturb, **kw)
    ref.close()
    assert all(bit_equal(f, f_ref).values()), bit_equal(f, f_ref)
    assert np.unique(f["density"]).size > 10
    if perturb is not None:         # the cloud's interior is no longer uniform
        f0 = P.alloc_fields(cfg, n, g)
        oracle.ic_cloud(oracle.numpy_block(f0, n, g, d), lower, **kw)
        inside = f0["density"] == kw["cloud_density"]
        assert inside.sum() > 100
        rel = f["density"][inside] / kw["cloud_density"] - 1.0
        assert 0.01 < np.abs(rel).max() < 5 * perturb[2]
        assert np.array_equal(f["density"][f0["density"] == kw["wind_density"]],
                              f0["density"][f0["density"] == kw["wind_density"]])


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_shock_tube_ic_restatement_equals_compiled_reference(axis):
    """vlct_ic_shock_tube against the reference's own EnzoInitialShockTube (+ the
    static helpers of EnzoInitialBCenter), compiled unmodified into oracle/_ref:
    Ryu-Jones 2a, Sod in the Mach-10 frame, and its flipped version"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    for setup in ("rj2a", "sod", "sod_flipped"):
        lower = [0.0, 0.0, 0.0]
        if setup == "rj2a":
            cfg, f, blk, n, g, d, _ = P.rj2a_setup(axis)
            kw = dict(setup="rj2a", aligned_ax=axis)
        else:
            flipped = setup.endswith("flipped")
            cfg, f, blk, n, g, d, _ = P.sod_de_setup(axis, flipped)
            if flipped:
                lower[axis] = -P.SOD_OFFSET
            kw = dict(setup="sod", aligned_ax=axis,
                      axis_velocity=P.SOD_BKG_VELOCITY, flipped=flipped)
        f_ref = P.alloc_fields(cfg, n, g)
        ref = oracle.CpuMethod(cfg, g, kind="ref")
        oracle.ic_shock_tube(oracle.numpy_block(f_ref, n, g, d), tuple(lower),
                             cfg.gamma, ref_method=ref, **kw)
        ref.close()
        eq = bit_equal(f, f_ref)
        assert all(eq.values()), (setup, [k for k, v in eq.items() if not v])


@pytest.mark.parametrize("name,mhd", [(n, True) for n in sorted(P.MHD_WAVES)]
                         + [(n, False) for n in sorted(P.HD_WAVES)])
def test_inclined_wave_ic_restatement_equals_compiled_reference(name, mhd):
    """vlct_ic_inclined_wave against the reference's own EnzoInitialInclinedWave
    (rotation, eigenvectors, vector potential -> face B -> centred B, total
    energy; same libm), bit for bit, both propagation directions"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    wave_type = (P.MHD_WAVES if mhd else P.HD_WAVES)[name][0]
    for positive_vel in (True, False):
        cfg, f, blk, n, g, d, _ = P.linear_wave_setup(name, 16, mhd, positive_vel)
        f_ref = P.alloc_fields(cfg, n, g)
        ref = oracle.CpuMethod(cfg, g, kind="ref")
        oracle.ic_inclined_wave(oracle.numpy_block(f_ref, n, g, d), (0.0, 0.0, 0.0),
                                cfg.gamma, wave_type, P.ALPHA, P.BETA, 1e-6, 1.0,
                                positive_vel, ref_method=ref)
        ref.close()
        eq = bit_equal(f, f_ref)
        assert all(eq.values()), (positive_vel, [k for k, v in eq.items() if not v])


def test_compiled_reference_reproduces_rj2a_golden():
    """the reference's own compiled sources on the same shock tube"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    cfg, f, g, dts = P.run_rj2a(0, kind="ref")
    cfg2, f2, g2, dts2 = P.run_rj2a(0, kind="oracle")
    assert dts == dts2
    assert all(bit_equal(f, f2).values())
    table = P.load_reference_table("rj2a_shock_tube_t0.2_res256.csv")
    norm = P.table_l1_norm(P.snapshot(cfg, f, g), table, 0, P.RJ2A_FIELDS)
    assert P.golden_isclose(norm, P.GOLDEN_RJ2A[0])


def test_sod_dual_energy_shock_tube_golden():
    """Sod tube in a Mach-10 frame with the dual-energy formalism (MHD solver,
    B = 0): golden norm of run_dual_energy_shock_tube_test.py:64-78"""
    cfg, f, g, dts = P.run_sod_de(0)
    norm = P.sod_de_l1_norm(P.snapshot(cfg, f, g), 0)
    assert P.golden_isclose(norm, P.GOLDEN_SOD_DE), (norm, P.GOLDEN_SOD_DE)


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_passive_scalar_sound_wave_golden(axis):
    """run_passive_advect_sound_test.py:66-70"""
    cfg, f, blk, n, g, d, t_final = P.passive_sound_setup(axis)
    m = oracle.CpuMethod(cfg, g)
    s0 = P.passive_snapshot(f, g)
    P.evolve(m, blk, t_final, lambda b: oracle.refresh_periodic(b, 1),
             dump_times=(0.0, t_final))
    m.close()
    norm = P.passive_l1_norm(s0, P.passive_snapshot(f, g))
    assert P.golden_isclose(norm, P.GOLDEN_PASSIVE_SOUND), norm


@pytest.mark.parametrize("kw", [dict(riemann="hllc", recon="plm", mhd=False),
                                dict(riemann="hllc", recon="plm", mhd=False,
                                     dual_energy=True, gamma=1.4, n_passive=2),
                                dict(riemann="hllc", recon="plm_athena", mhd=False,
                                     time_scheme="euler", courant=0.5)])
def test_face_fluxes_match_compiled_reference(kw):
    """flux-correction output: the restatement of save_fluxes_for_corrections_
    against what the reference's own Method deposits in FluxData"""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libvlct_ref.so not available")
    cfg = make_config(**kw)
    n, g, d = (11, 8, 6), (3, 3, 3), (0.1, 0.12, 0.09)
    nf = 6 + cfg.n_passive
    host = random_state(cfg, n, g, seed=61)
    out = {}
    for kind in ("oracle", "ref"):
        f = copy_state(host)
        blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
        m = oracle.CpuMethod(cfg, g, kind=kind, store_fluxes=True)
        dt = m.timestep(blk)
        m.compute(blk, dt)
        out[kind] = (m.face_fluxes(blk, dt, n, nf), f)
        m.close()
    assert all(bit_equal(out["oracle"][1], out["ref"][1]).values())
    a, b = out["oracle"][0], out["ref"][0]
    assert set(a) == set(b)
    for key in a:
        assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), key
        assert np.any(a[key] != 0.0), key
