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
