"""GPU parity at BASELINE.json's stated sizes, for flows that vary along all
three axes.

The CPU oracle is far too slow to evolve a 512^3 block, but one VL+CT step of
a cell reads only the 3 cells around it (ghost depth 3 = 1 + 2 stale layers,
EnzoMethodMHDVlct.cpp:124-133). So a WINDOW of the block -- w^3 cells plus 3
layers around them, cut out of the GPU's own state before the step -- evolved
by the oracle with the GPU's dt must reproduce, bit for bit, the GPU's result
on the window's w^3 cells (and on the faces bounding them). Doing that before
every step pins the whole run by induction; windows are placed in the
corners (largest indices: 32-bit index arithmetic, last chunks of the marching
sweeps), across chunk boundaries and in the middle.

  configs[0]  fast-mode linear wave at N = 64 (128 x 64 x 64; the reference's
              own files stop at N = 32): first cycles bit-identical to the
              oracle, the whole run second-order convergent w.r.t. the golden
              N = 16 / 32 norms
  configs[1]  Sod, 256^3, PLM + HLLC + dual energy: the tube is invariant across
              x, so a 256 x 4 x 4 oracle tube pins every (y, z) pencil
  configs[2/3] decaying MHD turbulence (problems.turbulence): 128^3 against the
              oracle on the whole block, 512^3 through windows + the full-size
              CFL timestep
"""
import numpy as np
import pytest

import problems as P
from helpers import make_config, bit_equal, max_abs_diff, oracle

pytestmark = pytest.mark.gpu

G3 = (3, 3, 3)
FACE_AXIS = {"bfieldi_x": 0, "bfieldi_y": 1, "bfieldi_z": 2}


def _names(cfg):
    names = ["density", "velocity_x", "velocity_y", "velocity_z", "total_energy"]
    if cfg.dual_energy:
        names.append("internal_energy")
    if cfg.mhd_choice == 1:
        names += ["bfield_x", "bfield_y", "bfield_z",
                  "bfieldi_x", "bfieldi_y", "bfieldi_z"]
    return names


def cut_window(f, names, lo, w, g=G3):
    """host copies of the window whose first cell (ghosts included) has array
    index lo = (ix, iy, iz); w = active cells per axis"""
    out = {}
    for k in names:
        sl = []
        for ax in (2, 1, 0):
            ext = w[ax] + 2 * g[ax] + (1 if FACE_AXIS.get(k, -1) == ax else 0)
            sl.append(slice(lo[ax], lo[ax] + ext))
        out[k] = f[k][tuple(sl)].cpu().numpy().copy()
    shape = tuple(w[ax] + 2 * g[ax] for ax in (2, 1, 0))
    out["pressure"] = np.zeros(shape)
    return out


def check_window(cfg, before, f_after, names, lo, w, d, dt, g=G3, passive=()):
    """oracle step of the window vs the GPU's result on its active cells"""
    blk = oracle.numpy_block(before, w, g, d, passive)
    cpu = oracle.CpuMethod(cfg, g)
    cpu.compute(blk, dt)
    cpu.close()
    bad = {}
    for k in names:
        sl_w, sl_f = [], []
        for ax in (2, 1, 0):
            ext = w[ax] + (1 if FACE_AXIS.get(k, -1) == ax else 0)
            sl_w.append(slice(g[ax], g[ax] + ext))
            sl_f.append(slice(lo[ax] + g[ax], lo[ax] + g[ax] + ext))
        want = before[k][tuple(sl_w)]
        got = f_after[k][tuple(sl_f)].cpu().numpy()
        if not np.array_equal(want.view(np.uint64), got.view(np.uint64)):
            bad[k] = float(np.max(np.abs(want - got)))
    return bad


def window_origins(m, w, g=G3):
    """array indices of window origins: both corners, the middle, and windows
    straddling the 64-face chunk boundaries of the marching sweeps"""
    span = [w[ax] + 2 * g[ax] for ax in range(3)]
    last = [m[ax] - span[ax] for ax in range(3)]
    mid = [max(0, min(last[ax], m[ax] // 2 - span[ax] // 2)) for ax in range(3)]
    chunk = [max(0, min(last[ax], 2 + 64 - span[ax] // 2)) for ax in range(3)]
    chunk2 = [max(0, min(last[ax], 2 + 7 * 64 - span[ax] // 2)) for ax in range(3)]
    outs = [(0, 0, 0), tuple(last), tuple(mid), tuple(chunk), tuple(chunk2),
            (last[0], 0, mid[2]), (0, last[1], chunk[2]), (mid[0], chunk[1], last[2])]
    return sorted(set(outs))


# ---------------------------------------------------------------------------
# decaying MHD turbulence
# ---------------------------------------------------------------------------
def turbulence_config():
    return make_config(riemann="hlld", recon="plm", theta=1.5, mhd=True,
                       courant=0.3, gamma=5.0 / 3.0)


def test_turbulence_128_whole_block_bit_identical_to_oracle():
    """a genuinely 3-D MHD flow on 128^3 cells: three cycles, every field with
    its ghost zones and every dt equal to the CPU oracle's"""
    import torch
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    N, g = 128, G3
    n, d = (N, N, N), (1.0 / N,) * 3
    cfg = turbulence_config()
    f = problems.turbulence(n, g, (0.0, 0.0, 0.0), d, n, device="cuda")
    host = {k: v.cpu().numpy().copy() for k, v in f.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(f, n, g, d)
    dts = []
    for _ in range(3):
        dts.append(method.timestep(block))
        method.refresh_periodic(block, 7)
        method.compute(block, dts[-1])
    method.synchronize()
    cpu = oracle.CpuMethod(cfg, g)
    blk = oracle.numpy_block(host, n, g, d)
    dts_cpu = []
    for _ in range(3):
        dts_cpu.append(cpu.timestep(blk))
        oracle.refresh_periodic(blk, 0)
        cpu.compute(blk, dts_cpu[-1])
    cpu.close()
    assert dts == dts_cpu
    got = {k: v.cpu().numpy() for k, v in f.items()}
    got.pop("pressure")
    want = {k: host[k] for k in got}
    eq = bit_equal(want, got)
    bad = {k: max_abs_diff(want, got)[k] for k, ok in eq.items() if not ok}
    assert not bad, bad
    # the flow is three-dimensional: B_z varies, v_z is not zero
    assert float(np.std(got["bfield_z"][g[2]:-g[2], g[1]:-g[1], g[0]:-g[0]])) > 1e-4
    assert float(np.abs(got["velocity_z"]).max()) > 0.1
    method.close()


@pytest.mark.slow
def test_turbulence_512_windows_and_timestep_bit_identical_to_oracle():
    """BASELINE configs[2]/[3] at full size (512^3 cells per GPU, PLM + HLLD +
    CT) with a 3-D flow: three cycles; before each the oracle evolves eight
    40^3 windows of the GPU's state (corners, chunk boundaries, middle) with
    the GPU's dt and must reproduce the GPU's cells bit for bit, and the CFL
    timestep of the first and last cycle equals the oracle's over the whole
    518^3 arrays."""
    import torch
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    N, g, w = 512, G3, (40, 40, 40)
    n, d = (N, N, N), (1.0 / N,) * 3
    cfg = turbulence_config()
    names = _names(cfg)
    f = problems.turbulence(n, g, (0.0, 0.0, 0.0), d, n, device="cuda")
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(f, n, g, d)
    m = tuple(N + 2 * g[ax] for ax in range(3))
    origins = window_origins(m, w)
    ncycles = 3
    for cycle in range(ncycles):
        dt = method.timestep(block)
        if cycle in (0, ncycles - 1):
            host = {k: f[k].cpu().numpy() for k in names if k not in FACE_AXIS}
            host["pressure"] = np.zeros_like(host["density"])
            # (timestep does not read the face-centred fields)
            blk = oracle.numpy_block(host, n, g, d)
            cpu = oracle.CpuMethod(cfg, g)
            dt_cpu = cpu.timestep(blk)
            cpu.close()
            assert dt == dt_cpu, (cycle, dt, dt_cpu)
            del host, blk
        method.refresh_periodic(block, 7)
        method.synchronize()
        wins = [cut_window(f, names, lo, w) for lo in origins]
        method.compute(block, dt)
        method.synchronize()
        for lo, before in zip(origins, wins):
            bad = check_window(cfg, before, f, names, lo, w, d, dt)
            assert not bad, (cycle, lo, bad)
    # all HLLD regions are populated along z, unlike the z-extruded vortex
    assert float(f["bfield_x"].abs().max()) > 1e-3
    method.close()


def test_turbulence_brick_equals_global_on_device():
    """a brick of the decomposed turbulence IC holds the bits of the undivided
    domain (what block-decomposition invariance of the multi-GPU runs needs)"""
    import torch
    from enzo_e_b200 import problems
    g, N = G3, (32, 24, 16)
    d = tuple(1.0 / x for x in N)
    glob = problems.turbulence(N, g, (0.0, 0.0, 0.0), d, N, device="cuda")
    n_loc = (16, 12, 8)
    for c in [(0, 0, 0), (1, 1, 1), (1, 0, 1)]:
        lower = tuple(c[a] * n_loc[a] * d[a] for a in range(3))
        part = problems.turbulence(n_loc, g, lower, d, N, device="cuda")
        for k, v in part.items():
            face = FACE_AXIS.get(k, -1)
            sl_l, sl_g = [], []
            for ax in (2, 1, 0):
                ext = n_loc[ax] + (1 if face == ax else 0)
                sl_l.append(slice(g[ax], g[ax] + ext))
                lo = g[ax] + c[ax] * n_loc[ax]
                sl_g.append(slice(lo, lo + ext))
            assert torch.equal(v[tuple(sl_l)], glob[k][tuple(sl_g)]), (c, k)
    # ghosts hold their periodic images
    v = glob["velocity_x"]
    assert torch.equal(v[:, :, :3], v[:, :, N[0]:N[0] + 3])
    assert torch.equal(v[:3], v[N[2]:N[2] + 3])


# ---------------------------------------------------------------------------
# configs[0]: fast magnetosonic linear wave at N = 64
# ---------------------------------------------------------------------------
def test_fast_wave_n64_first_cycles_bit_identical_and_convergent():
    """input/vlct/MHD_linear_wave at twice the resolution of the reference's
    finest file (method_vlct_fastN32.in): 128 x 64 x 64 cells. The first four
    cycles are bit-identical to the oracle (all fields, ghost zones, dt); the
    whole run to t = 0.5 gives an L1 norm that continues the second-order
    convergence of the two golden values (N16 / N32 = 4.96)."""
    from test_gpu_golden import GpuRun
    N = 64
    cfg, f, blk, n, g, d, t_final = P.linear_wave_setup("fast", N, True)
    host = {k: v.copy() for k, v in f.items()}
    run = GpuRun(cfg, f, n, g, d)
    # -- first cycles against the oracle
    cpu = oracle.CpuMethod(cfg, g)
    cblk = oracle.numpy_block(host, n, g, d)
    s0 = P.snapshot(cfg, f, g)
    t, ncheck, dts = 0.0, 4, []
    for _ in range(ncheck):
        dt = run.timestep()
        dt_cpu = cpu.timestep(cblk)
        assert dt == dt_cpu
        run.refresh()
        oracle.refresh_periodic(cblk, 0)
        run.compute(None, dt)
        cpu.compute(cblk, dt_cpu)
        t += dt
        dts.append(dt)
    cpu.close()
    got = {k: np.empty_like(v) for k, v in f.items()}
    run.download(got)
    eq = bit_equal({k: host[k] for k in got if k != "pressure"},
                   {k: got[k] for k in got if k != "pressure"})
    assert all(eq.values()), {k: v for k, v in eq.items() if not v}
    # -- the rest of the run on the device
    while t < t_final:
        dt = min(run.timestep(), t_final - t)
        run.refresh()
        run.compute(None, dt)
        t += dt
    run.download(f)
    run.close()
    l1 = P.l1_error_norm(s0, P.snapshot(cfg, f, g), P.LINWAVE_FIELDS_MHD, N)
    g16, g32 = P.GOLDEN_MHD[("fast", 16)], P.GOLDEN_MHD[("fast", 32)]
    assert l1 < g32 / 3.0, (l1, g32)          # keeps converging ...
    assert l1 > g32 / 6.0, (l1, g32)          # ... at second order, not better


# ---------------------------------------------------------------------------
# configs[1]: Sod, 256^3, PLM + HLLC + dual energy
# ---------------------------------------------------------------------------
def test_sod_256_hllc_dual_energy_bit_identical_to_oracle_tube():
    """3-D Sod problem on 256^3 cells (hydro, PLM theta 1.5, HLLC, modern dual
    energy), outflow along x, periodic across: the problem is invariant across
    the tube, so a 256 x 4 x 4 oracle tube started from the GPU's own initial
    pencil pins every pencil of the block: 6 cycles, all fields of the lowest
    and highest (y, z) pencils with their ghost zones, every dt, and exact
    invariance across the tube on the device."""
    import torch
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    N, ns, g = 256, 4, G3
    n, d = (N, N, N), (1.0 / N,) * 3
    cfg = make_config(riemann="hllc", recon="plm", theta=1.5, mhd=False,
                      gamma=1.4, dual_energy=True, eta=0.001, courant=0.3)
    f = problems.hydro_sod(n, g, (0.0, 0.0, 0.0), d, device="cuda", gamma=1.4)
    ms = ns + 2 * g[1]
    host = {k: v[:ms, :ms, :].cpu().numpy().copy() for k, v in f.items()}
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(f, n, g, d)
    cpu = oracle.CpuMethod(cfg, g)
    blk = oracle.numpy_block(host, (N, ns, ns), g, d)
    ncycles = 6
    for _ in range(ncycles):
        dt = method.timestep(block)
        assert dt == cpu.timestep(blk)
        method.refresh_periodic(block, 6)
        method.boundary(block, 0, 0, "outflow")
        method.boundary(block, 0, 1, "outflow")
        oracle.refresh_periodic(blk, 0, 6)
        oracle.boundary(blk, 0, 0, "outflow")
        oracle.boundary(blk, 0, 1, "outflow")
        method.compute(block, dt)
        cpu.compute(blk, dt)
    cpu.close()
    method.synchronize()
    half = ms // 2
    for k, v in f.items():
        if k == "pressure":
            continue
        for (zs, ys, hz, hy) in [(slice(0, half), slice(0, half), slice(0, half), slice(0, half)),
                                 (slice(-half, None), slice(-half, None),
                                  slice(half, None), slice(half, None))]:
            got = v[zs, ys, :].cpu().numpy()
            want = host[k][hz, hy, :]
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), k
        inner = v[g[2]:-g[2], g[1]:-g[1], :]
        assert bool((inner == inner[0:1, 0:1, :]).all()), f"{k} varies across the tube"
    # the shock has moved: the state is no longer the two constant halves
    rho = f["density"][g[2], g[1], g[0]:-g[0]].cpu().numpy()
    assert len(np.unique(rho)) > 10
    method.close()


@pytest.mark.slow
def test_blast_256_hllc_dual_energy_windows():
    """configs[1]b, a 3-D hydro flow at 256^3: Sedov-like blast, PLM + HLLC +
    dual energy; windows around the blast and in the corners against the
    oracle for four cycles (the blast wave crosses the windows' cells)."""
    import torch
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    N, g, w = 256, G3, (32, 32, 32)
    n, d = (N, N, N), (1.0 / N,) * 3
    cfg = make_config(riemann="hllc", recon="plm", theta=1.5, mhd=False,
                      gamma=5.0 / 3.0, dual_energy=True, eta=0.001, courant=0.3,
                      dfloor=1e-10, pfloor=1e-10)
    names = _names(cfg)
    f = problems.hydro_blast(n, g, (0.0, 0.0, 0.0), d, device="cuda")
    method = EnzoMethodMHDVlct(config=cfg)
    block = Block(f, n, g, d)
    m = tuple(N + 2 * g[ax] for ax in range(3))
    c = m[0] // 2
    origins = [(c - 19, c - 19, c - 19), (c - 30, c - 8, c - 19), (0, 0, 0),
               tuple(m[ax] - w[ax] - 2 * g[ax] for ax in range(3))]
    for cycle in range(4):
        dt = method.timestep(block)
        method.refresh_periodic(block, 7)
        method.synchronize()
        wins = [cut_window(f, names, lo, w) for lo in origins]
        method.compute(block, dt)
        method.synchronize()
        for lo, before in zip(origins, wins):
            bad = check_window(cfg, before, f, names, lo, w, d, dt)
            assert not bad, (cycle, lo, bad)
    assert float(f["velocity_x"].abs().max()) > 0.1
    method.close()
