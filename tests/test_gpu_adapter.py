"""GPU: the reference-side binding. integration/EnzoMethodMHDVlctGpu.cpp is the
C++ `Method` subclass a maintainer adds to Enzo-E; here it is compiled against
the reference's own headers (+ the Cello stand-ins that also build the CPU
reference) and driven exactly like EnzoMethodMHDVlct: construct from a
ParameterGroup, timestep(block), compute(block) on host-resident Cello fields.
Results must equal the compiled reference's bit for bit."""
import pytest

from helpers import (make_config, random_state, copy_state, passive_names,
                     bit_equal, max_abs_diff, oracle)

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("kw", [
    dict(riemann="hlld", recon="plm", theta=1.5, mhd=True),
    dict(riemann="hllc", recon="plm", mhd=False, dual_energy=True, gamma=1.4,
         n_passive=2),
    dict(riemann="hlle", recon="plm_athena", mhd=True, accel=True),
], ids=["mhd_hlld", "hd_hllc_de_scalars", "mhd_hlle_gravity"])
def test_cxx_method_adapter_matches_reference(kw):
    if not (oracle.have_adapter() and oracle.have_ref()):
        pytest.skip("oracle/_ref/libvlct_adapter.so or libvlct_ref.so not "
                    "available (they are built where /root/reference exists)")
    cfg = make_config(**kw)
    n, g, d = (18, 12, 10), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=17)
    want, dts_want = _run(cfg, host, n, g, d, 2, "ref")
    got, dts_got = _run(cfg, host, n, g, d, 2, "adapter")
    assert dts_got == dts_want
    eq = bit_equal(want, got)
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, bad


def test_cxx_method_adapter_stores_fluxes_for_corrections():
    """constructed with store_fluxes_for_corrections = true, the C++ adapter
    deposits in the block's FluxData exactly what the reference's own Method
    deposits (EnzoMethodMHDVlct.cpp:250-330, 480-490)"""
    import numpy as np
    if not (oracle.have_adapter() and oracle.have_ref()):
        pytest.skip("oracle/_ref libraries not available")
    cfg = make_config(riemann="hllc", recon="plm", mhd=False, dual_energy=True,
                      gamma=1.4, n_passive=2)
    n, g, d = (14, 9, 8), (3, 3, 3), (0.1, 0.12, 0.09)
    nf = 6 + cfg.n_passive
    host = random_state(cfg, n, g, seed=19)
    out = {}
    for kind in ("ref", "adapter"):
        f = copy_state(host)
        blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
        m = oracle.CpuMethod(cfg, g, kind=kind, store_fluxes=True)
        dt = m.timestep(blk)
        m.compute(blk, dt)
        out[kind] = (m.face_fluxes(blk, dt, n, nf), f, dt)
        m.close()
    assert out["ref"][2] == out["adapter"][2]
    assert all(bit_equal(out["ref"][1], out["adapter"][1]).values())
    a, b = out["ref"][0], out["adapter"][0]
    for key in a:
        assert np.array_equal(a[key].view(np.uint64), b[key].view(np.uint64)), key


def _many_blocks(cfg, nblocks, n, g, seed0):
    return [random_state(cfg, n, g, seed=seed0 + i) for i in range(nblocks)]


@pytest.mark.parametrize("fused", [False, True], ids=["two_calls", "fused_timestep"])
@pytest.mark.parametrize("kw,nblocks,n", [
    (dict(riemann="hlld", recon="plm", theta=1.5, mhd=True), 64, (16, 16, 16)),
    (dict(riemann="hllc", recon="plm", mhd=False, dual_energy=True, gamma=1.4,
          n_passive=2), 24, (16, 8, 12)),
], ids=["mhd_64x16cube", "hd_de_scalars_24"])
def test_cxx_adapter_batches_the_blocks_of_a_process(kw, nblocks, n, fused):
    """Enzo-E's operating point: many small blocks per process. Driven the way
    Cello's compute phase drives a Method (compute(block) for one block after
    the other, src/Cello/control_compute.cpp:72-112), the adapter queues the
    blocks and advances them with ONE vlct_compute_batch when the last one
    arrives; every block must equal what the compiled reference makes of it
    alone, for two cycles, including the timesteps of the stopping phase
    (served from the fused device-side CFL kernel when gpu_fused_timestep is
    set)."""
    if not (oracle.have_adapter() and oracle.have_ref()):
        pytest.skip("oracle/_ref libraries not available")
    cfg = make_config(**kw)
    g, d = (3, 3, 3), (0.1, 0.12, 0.09)
    hosts = _many_blocks(cfg, nblocks, n, g, 100)
    # reference: every block on its own
    want, want_dts = [], []
    ref = oracle.CpuMethod(cfg, g, kind="ref")
    for h in hosts:
        f = copy_state(h)
        blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
        want.append(f)
        want_dts.append([ref.timestep(blk)])
    for cycle in range(2):
        dt = min(x[-1] for x in want_dts)
        for f, dts in zip(want, want_dts):
            blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
            ref.compute(blk, dt)
            dts.append(ref.timestep(blk))
    ref.close()
    # the adapter: all blocks of the "process" per phase
    got = [copy_state(h) for h in hosts]
    blks = [oracle.numpy_block(f, n, g, d, passive_names(cfg)) for f in got]
    m = oracle.CpuMethod(cfg, g, kind="adapter", gpu_batch_blocks=True,
                         gpu_fused_timestep=fused)
    dts = m.timestep_many(blks, cycle=0)
    assert dts == [x[0] for x in want_dts]
    for cycle in range(2):
        dt = min(dts)
        deferred = m.compute_many(blks, dt, cycle=cycle)
        assert deferred == nblocks - 1      # nobody moved on before the flush
        dts = m.timestep_many(blks, cycle=cycle + 1)
        ref_dts = [x[cycle + 1] for x in want_dts]
        if fused:   # the batch minimum serves every block
            assert dts == [min(ref_dts)] * nblocks
        else:
            assert dts == ref_dts
    m.close()
    for i, (w, gt) in enumerate(zip(want, got)):
        eq = bit_equal(w, gt)
        bad = {k: max_abs_diff(w, gt)[k] for k, ok in eq.items() if not ok}
        assert not bad, (i, bad)


def test_cxx_adapter_without_batching_reports_done_at_once():
    if not (oracle.have_adapter() and oracle.have_ref()):
        pytest.skip("oracle/_ref libraries not available")
    cfg = make_config(riemann="hlld", recon="plm", mhd=True)
    n, g, d = (12, 8, 8), (3, 3, 3), (0.1, 0.12, 0.09)
    hosts = _many_blocks(cfg, 3, n, g, 300)
    got = [copy_state(h) for h in hosts]
    blks = [oracle.numpy_block(f, n, g, d) for f in got]
    m = oracle.CpuMethod(cfg, g, kind="adapter", gpu_batch_blocks=False)
    assert m.compute_many(blks, 1e-3) == 0
    m.close()
    ref = oracle.CpuMethod(cfg, g, kind="ref")
    for h, gt in zip(hosts, got):
        f = copy_state(h)
        ref.compute(oracle.numpy_block(f, n, g, d), 1e-3)
        assert all(bit_equal(f, gt).values())
    ref.close()


@pytest.mark.parametrize("kw", [
    dict(riemann="hlld", recon="plm_athena", theta=1.5, mhd=True, dual_energy=True,
         n_passive=2, courant=0.25, dfloor=1e-6, pfloor=1e-7),
    dict(riemann="hllc", recon="plm", mhd=False, time_scheme="euler", courant=0.5),
], ids=["mhd_de_scalars", "hd_euler"])
def test_cxx_adapter_pup_roundtrip(kw):
    """Charm++ migration / checkpoint-restart of the Method
    (EnzoMethodMHDVlct.cpp:170-197): pack through a PUP::er, construct with
    the migration constructor, unpack -- the rebuilt object (new library
    handle, same configuration, passive-scalar names and options) continues
    bit-identically to the compiled reference."""
    if not (oracle.have_adapter() and oracle.have_ref()):
        pytest.skip("oracle/_ref libraries not available")
    cfg = make_config(**kw)
    n, g, d = (18, 12, 10), (3, 3, 3), (0.1, 0.12, 0.09)
    host = random_state(cfg, n, g, seed=23)
    want, dts_want = _run(cfg, host, n, g, d, 3, "ref")
    f = copy_state(host)
    blk = oracle.numpy_block(f, n, g, d, passive_names(cfg))
    m = oracle.CpuMethod(cfg, g, kind="adapter")
    dts = []
    for step in range(3):
        if step == 1:
            size = m.pup_roundtrip()
            import ctypes
            from enzo_e_b200 import abi
            # vlct_config + the passive-scalar names + three flags
            assert size >= ctypes.sizeof(abi.VlctConfig) + 8 + 3
        dt = m.timestep(blk)
        m.compute(blk, dt)
        dts.append(dt)
    m.close()
    assert dts == dts_want
    assert all(bit_equal(want, f).values())
