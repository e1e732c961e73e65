"""Synthetic problem set-ups generated directly in device memory (torch is used
as the array library only). Each returns {field name: tensor} for one block of
a periodic unigrid, ghost zones included, in the Cello field layout.

orszag_tang   input/vlct/orszag-tang/orszag-tang.in:22-51 of the reference
              (Value initialiser + vlct_bfield vector potential,
              src/Enzo/initial/EnzoInitialBCenter.cpp), extruded along z
mhd_blast     uniform medium with an over-pressured sphere and an oblique field
"""
import math

import torch

from . import abi


def _coords(n_local, g, lower, width, device):
    """cell-centre and face coordinates of a block, ghosts included"""
    out = []
    for ax in range(3):
        m = n_local[ax] + 2 * g[ax]
        # positions are formed from the GLOBAL cell index (the block's lower
        # corner is a whole number of cells from the origin), so that a brick
        # of a decomposed domain gets bit for bit the values the undivided
        # domain has there
        off = round(lower[ax] / width[ax])
        idx = torch.arange(m + 1, dtype=torch.float64, device=device) - g[ax] + off
        face = width[ax] * idx
        cen = width[ax] * (0.5 + idx[:-1])
        out.append((cen, face))
    return out


def _center_b(f):
    f["bfield_x"] = 0.5 * (f["bfieldi_x"][:, :, :-1] + f["bfieldi_x"][:, :, 1:])
    f["bfield_y"] = 0.5 * (f["bfieldi_y"][:, :-1, :] + f["bfieldi_y"][:, 1:, :])
    f["bfield_z"] = 0.5 * (f["bfieldi_z"][:-1, :, :] + f["bfieldi_z"][1:, :, :])


def orszag_tang(n_local, g, lower, width, device="cuda", gamma=5.0 / 3.0,
                n_passive=0):
    (xc, xf), (yc, yf), (zc, zf) = _coords(n_local, g, lower, width, device)
    mz, my, mx = zc.numel(), yc.numel(), xc.numel()
    pi = math.pi
    shape = (mz, my, mx)
    X = xc.view(1, 1, mx).expand(shape)
    Y = yc.view(1, my, 1).expand(shape)
    f = {}
    f["density"] = torch.full(shape, 25.0 / (36.0 * pi), dtype=torch.float64,
                              device=device)
    f["velocity_x"] = (-1.0 * torch.sin(2.0 * pi * Y)).contiguous()
    f["velocity_y"] = torch.sin(2.0 * pi * X).contiguous()
    f["velocity_z"] = torch.zeros(shape, dtype=torch.float64, device=device)
    sy, sx = torch.sin(2.0 * pi * Y), torch.sin(2.0 * pi * X)
    etot = 0.9 + 0.5 * (sy * sy + sx * sx)

    # A_z on the (x-face, y-face) corners; B_x = dAz/dy, B_y = -dAz/dx
    def az(x, y):
        return (1.0 / math.sqrt(4.0 * pi)) * (torch.cos(4.0 * pi * x) / (4.0 * pi)
                                             + torch.cos(2.0 * pi * y) / (2.0 * pi))
    Az = az(xf.view(1, mx + 1), yf.view(my + 1, 1))            # (my+1, mx+1)
    bx = (Az[1:, :] - Az[:-1, :]) / width[1]                   # (my, mx+1)
    by = -(Az[:, 1:] - Az[:, :-1]) / width[0]                  # (my+1, mx)
    f["bfieldi_x"] = bx.view(1, my, mx + 1).expand(mz, my, mx + 1).contiguous()
    f["bfieldi_y"] = by.view(1, my + 1, mx).expand(mz, my + 1, mx).contiguous()
    f["bfieldi_z"] = torch.zeros((mz + 1, my, mx), dtype=torch.float64,
                                 device=device)
    _center_b(f)
    mag = 0.5 * (f["bfield_x"] * f["bfield_x"] + f["bfield_y"] * f["bfield_y"]
                 + f["bfield_z"] * f["bfield_z"]) / f["density"]
    f["total_energy"] = (etot + mag).contiguous()
    f["pressure"] = torch.zeros(shape, dtype=torch.float64, device=device)
    Z = zc.view(mz, 1, 1).expand(shape)
    for k in range(n_passive):
        f[f"passive_{k}"] = (f["density"] * (0.5 + 0.4 * torch.sin(
            2.0 * pi * ((k + 1) * X + Y + Z)))).contiguous()
    return {k: v.contiguous() for k, v in f.items()}


def mhd_blast(n_local, g, lower, width, device="cuda", gamma=5.0 / 3.0,
              center=(0.5, 0.5, 0.5), radius=0.1, p_in=10.0, p_out=0.1):
    (xc, xf), (yc, yf), (zc, zf) = _coords(n_local, g, lower, width, device)
    mz, my, mx = zc.numel(), yc.numel(), xc.numel()
    shape = (mz, my, mx)
    X = xc.view(1, 1, mx).expand(shape)
    Y = yc.view(1, my, 1).expand(shape)
    Z = zc.view(mz, 1, 1).expand(shape)
    r2 = (X - center[0]) ** 2 + (Y - center[1]) ** 2 + (Z - center[2]) ** 2
    f = {}
    one = torch.ones(shape, dtype=torch.float64, device=device)
    f["density"] = one.clone()
    for k in "xyz":
        f["velocity_" + k] = torch.zeros(shape, dtype=torch.float64, device=device)
    p = torch.where(r2 < radius * radius, p_in * one, p_out * one)
    b0 = 1.0 / math.sqrt(2.0)
    f["bfieldi_x"] = torch.full((mz, my, mx + 1), b0, dtype=torch.float64, device=device)
    f["bfieldi_y"] = torch.full((mz, my + 1, mx), b0, dtype=torch.float64, device=device)
    f["bfieldi_z"] = torch.zeros((mz + 1, my, mx), dtype=torch.float64, device=device)
    _center_b(f)
    mag = 0.5 * (f["bfield_x"] ** 2 + f["bfield_y"] ** 2 + f["bfield_z"] ** 2)
    f["total_energy"] = (p / ((gamma - 1.0) * f["density"]) + mag / f["density"]).contiguous()
    f["pressure"] = torch.zeros(shape, dtype=torch.float64, device=device)
    return {k: v.contiguous() for k, v in f.items()}


def field_bytes(fields):
    return sum(v.numel() * v.element_size() for v in fields.values())


def hydro_sod(n_local, g, lower, width, device="cuda", gamma=1.4,
              dual_energy=True, split=0.5):
    """3-D Sod problem (BASELINE configs[1]a): the shock-tube states of
    src/Enzo/initial/EnzoInitialShockTube.cpp:37-80 (rho, p = 1, 1 | 0.125,
    0.1, v = 0) split at x = `split`."""
    (xc, _), (yc, _), (zc, _) = _coords(n_local, g, lower, width, device)
    mz, my, mx = zc.numel(), yc.numel(), xc.numel()
    shape = (mz, my, mx)
    left = (xc.view(1, 1, mx) < split).expand(shape)
    one = torch.ones(shape, dtype=torch.float64, device=device)
    f = {"density": torch.where(left, 1.0 * one, 0.125 * one)}
    p = torch.where(left, 1.0 * one, 0.1 * one)
    for k in "xyz":
        f["velocity_" + k] = torch.zeros(shape, dtype=torch.float64, device=device)
    eint = p / ((gamma - 1.0) * f["density"])
    f["total_energy"] = eint.clone()
    if dual_energy:
        f["internal_energy"] = eint.clone()
    f["pressure"] = torch.zeros(shape, dtype=torch.float64, device=device)
    return {k: v.contiguous() for k, v in f.items()}


def hydro_blast(n_local, g, lower, width, device="cuda", gamma=5.0 / 3.0,
                dual_energy=True, center=(0.5, 0.5, 0.5), radius_cells=3.5,
                p_in=1.0e2, p_out=1.0e-5):
    """Sedov-like blast (BASELINE configs[1]b): rho = 1, p = 1e-5 with p = 1e2
    inside 3.5 cells of the centre."""
    (xc, _), (yc, _), (zc, _) = _coords(n_local, g, lower, width, device)
    mz, my, mx = zc.numel(), yc.numel(), xc.numel()
    shape = (mz, my, mx)
    r2 = ((xc.view(1, 1, mx) - center[0]) ** 2 + (yc.view(1, my, 1) - center[1]) ** 2
          + (zc.view(mz, 1, 1) - center[2]) ** 2).expand(shape)
    one = torch.ones(shape, dtype=torch.float64, device=device)
    f = {"density": one.clone()}
    p = torch.where(r2 < (radius_cells * width[0]) ** 2, p_in * one, p_out * one)
    for k in "xyz":
        f["velocity_" + k] = torch.zeros(shape, dtype=torch.float64, device=device)
    eint = p / ((gamma - 1.0) * f["density"])
    f["total_energy"] = eint.clone()
    if dual_energy:
        f["internal_energy"] = eint.clone()
    f["pressure"] = torch.zeros(shape, dtype=torch.float64, device=device)
    return {k: v.contiguous() for k, v in f.items()}


def turbulence_modes(seed=20240517, kmax=3, gamma=5.0 / 3.0, mach=0.5,
                     rho0=1.0, p0=1.0):
    """The Fourier modes of the decaying-turbulence velocity field (BASELINE
    configs[3]): every integer wave vector of the half space with
    1 <= |k| <= kmax gets a Gaussian amplitude vector, projected perpendicular
    to k (solenoidal), and a uniform random phase, all drawn from
    numpy.random.default_rng(seed) in a fixed order. The amplitudes are scaled
    analytically (orthogonal modes: <v^2> = sum |a|^2 / 2) to an rms Mach
    number `mach`, so that the field is a pure function of position and a brick
    of a decomposed domain gets the values of the undivided one. Returns
    (k [M,3] ints, a [M,3], phase [M]) as lists.

    The reference's own generator (src/Enzo/initial/EnzoInitialTurbulence.cpp:
    54-120 -> turboinit.F) is Fortran and cannot be built here; only the
    spectrum shape (low-k solenoidal modes, fixed Mach number) is kept."""
    import numpy as np
    rng = np.random.default_rng(seed)
    ks = []
    for kz in range(-kmax, kmax + 1):
        for ky in range(-kmax, kmax + 1):
            for kx in range(-kmax, kmax + 1):
                k2 = kx * kx + ky * ky + kz * kz
                if k2 < 1 or k2 > kmax * kmax:
                    continue
                # one of every +-k pair: first non-zero of (kz, ky, kx) positive
                lead = kz if kz != 0 else (ky if ky != 0 else kx)
                if lead < 0:
                    continue
                ks.append((kx, ky, kz))
    k = np.array(ks, dtype=np.float64)
    a = rng.standard_normal(k.shape)
    phase = rng.uniform(0.0, 2.0 * math.pi, size=len(ks))
    khat = k / np.sqrt((k * k).sum(axis=1, keepdims=True))
    a = a - khat * (a * khat).sum(axis=1, keepdims=True)
    # Kolmogorov-like weighting |k|^(-11/6) of the mode amplitudes
    a = a * ((k * k).sum(axis=1, keepdims=True)) ** (-11.0 / 12.0)
    cs2 = gamma * p0 / rho0
    scale = math.sqrt((mach * mach * cs2) / (0.5 * float((a * a).sum())))
    a = a * scale
    return ks, a.tolist(), phase.tolist()


def turbulence(n_local, g, lower, width, global_n, device="cuda",
               gamma=5.0 / 3.0, seed=20240517, mach=0.5, beta=2.0, n_passive=0):
    """Decaying MHD turbulence (BASELINE configs[3]): rho = 1, p = 1, a uniform
    field B = (0, 0, B0) with plasma beta = 2 p / B0^2, and a solenoidal
    velocity field of low-k Fourier modes (turbulence_modes). Cell positions
    come from the global cell index wrapped into the periodic domain of
    `global_n` cells, so ghost cells hold exactly the values of their periodic
    images and every brick of a decomposition the values of the undivided
    domain. Genuinely three-dimensional: every sweep direction sees all HLLD
    regions."""
    shape = tuple(n_local[ax] + 2 * g[ax] for ax in (2, 1, 0))
    mz, my, mx = shape
    pos = []
    for ax in range(3):
        m = n_local[ax] + 2 * g[ax]
        off = round(lower[ax] / width[ax])
        idx = torch.arange(m, dtype=torch.float64, device=device) - g[ax] + off
        idx = torch.remainder(idx, float(global_n[ax]))
        pos.append((0.5 + idx) / float(global_n[ax]))        # in [0, 1)
    X = pos[0].view(1, 1, mx)
    Y = pos[1].view(1, my, 1)
    Z = pos[2].view(mz, 1, 1)
    ks, amp, phase = turbulence_modes(seed, gamma=gamma, mach=mach)
    v = [torch.zeros(shape, dtype=torch.float64, device=device) for _ in range(3)]
    two_pi = 2.0 * math.pi
    for (kx, ky, kz), a, ph in zip(ks, amp, phase):
        c = torch.cos(two_pi * (kx * X + ky * Y + kz * Z) + ph)
        for comp in range(3):
            v[comp] += a[comp] * c
        del c
    f = {"density": torch.ones(shape, dtype=torch.float64, device=device)}
    for comp, name in enumerate("xyz"):
        f["velocity_" + name] = v[comp]
    b0 = math.sqrt(2.0 * 1.0 / beta)
    f["bfieldi_x"] = torch.zeros((mz, my, mx + 1), dtype=torch.float64, device=device)
    f["bfieldi_y"] = torch.zeros((mz, my + 1, mx), dtype=torch.float64, device=device)
    f["bfieldi_z"] = torch.full((mz + 1, my, mx), b0, dtype=torch.float64, device=device)
    _center_b(f)
    ke = 0.5 * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    f["total_energy"] = 1.0 / (gamma - 1.0) + ke + 0.5 * b0 * b0
    f["pressure"] = torch.zeros(shape, dtype=torch.float64, device=device)
    for s in range(n_passive):
        f[f"passive_{s}"] = (f["density"] * (0.5 + 0.4 * torch.sin(
            two_pi * ((s + 1) * X + Y + Z)))).expand(shape).contiguous()
    return {k: t.contiguous() for k, t in f.items()}


# ---------------------------------------------------------------------------
# the reference's answer-test problems, generated in device memory
# ---------------------------------------------------------------------------
_CELLO_PI = 3.14159265358979324          # src/Cello/cello.hpp:623


def _rotation(alpha, beta):
    """Rotation (src/Enzo/initial/EnzoInitialInclinedWave.cpp:54-117): the
    wave travels along x0, the axis x rotated by beta about z and alpha about
    the new y"""
    ca, sa, cb, sb = math.cos(alpha), math.sin(alpha), math.cos(beta), math.sin(beta)
    return [[ca * cb, ca * sb, sa], [-1.0 * sb, cb, 0.0],
            [-1.0 * sa * cb, -1.0 * sa * sb, ca]]


def _wave_table(wave_type, gamma, positive_vel, parallel_vel):
    """background state and eigenvector of a linear wave in conserved form
    (prepare_MHD_initializers_ / prepare_HD_initializers_,
    EnzoInitialInclinedWave.cpp:941-1017, 1022-1122)"""
    sgn = 1.0 if positive_vel else -1.0
    w = {"b_back": (0.0, 0.0, 0.0), "b1_ev": 0.0, "b2_ev": 0.0, "has_a": False}
    hd = wave_type in ("sound", "hd_entropy", "hd_transv_entropy_v1",
                       "hd_transv_entropy_v2")
    if hd:
        v0 = 0.0
        if parallel_vel is not None:
            v0 = parallel_vel
        elif wave_type != "sound":
            v0 = sgn
        v2sq = v0 * v0
        w.update(rho_back=1.0, mom_back=(v0, 0.0, 0.0),
                 etot_back=(1.0 / gamma) / (gamma - 1.0) + 0.5 * v2sq)
        if wave_type == "sound":
            h_back = 1.0 / (gamma - 1.0) + 0.5 * v2sq
            cs = sgn * 1.0
            w.update(rho_ev=1.0, mom_ev=(v0 + cs, 0.0, 0.0), etot_ev=h_back + v0 * cs)
        elif wave_type == "hd_entropy":
            w.update(rho_ev=1.0, mom_ev=(v0, 0.0, 0.0), etot_ev=0.5 * v2sq)
        elif wave_type == "hd_transv_entropy_v1":
            w.update(rho_ev=0.0, mom_ev=(0.0, 1.0, 0.0), etot_ev=0.0)
        else:
            w.update(rho_ev=0.0, mom_ev=(0.0, 0.0, 1.0), etot_ev=0.0)
        return w
    w.update(rho_back=1.0, mom_back=(0.0, 0.0, 0.0), b_back=(1.0, 1.5, 0.0),
             etot_back=(1.0 / gamma) / (gamma - 1.0) + 1.625, has_a=True)
    if wave_type == "mhd_entropy":
        w["mom_back"] = (sgn, 0.0, 0.0)
        w["etot_back"] += 0.5
    c = 0.5 / math.sqrt(5.0)
    if wave_type == "fast":
        w.update(rho_ev=2.0 * c, mom_ev=(sgn * 4.0 * c, -1.0 * sgn * 2.0 * c, 0.0),
                 etot_ev=9.0 * c, b1_ev=4.0 * c)
    elif wave_type == "alfven":
        w.update(rho_ev=0.0, mom_ev=(0.0, 0.0, -1.0 * sgn), etot_ev=0.0, b2_ev=1.0)
    elif wave_type == "slow":
        w.update(rho_ev=4.0 * c, mom_ev=(sgn * 2.0 * c, sgn * 4.0 * c, 0.0),
                 etot_ev=3.0 * c, b1_ev=-2.0 * c)
    elif wave_type == "mhd_entropy":
        w.update(rho_ev=1.0, mom_ev=(sgn, 0.0, 0.0), etot_ev=0.5)
    else:
        raise ValueError(f"unknown wave type {wave_type!r}")
    return w


def inclined_wave(n_local, g, lower, width, wave_type, alpha, beta, device="cuda",
                  gamma=5.0 / 3.0, amplitude=1e-6, lam=1.0, positive_vel=True,
                  parallel_vel=None, mhd=True, dual_energy=False):
    """EnzoInitialInclinedWave (src/Enzo/initial/EnzoInitialInclinedWave.cpp:
    setup_fluid_ :543-643, setup_bfield :426-490) with the face B taken from
    the curl of the vector potential (EnzoInitialBCenter.cpp:43-127): the
    linear-wave problems of input/vlct/{MHD,HD}_linear_wave, generated where
    they are used. wave_type: fast | alfven | slow | mhd_entropy | sound |
    hd_entropy | hd_transv_entropy_v1 | hd_transv_entropy_v2."""
    (xc, xf), (yc, yf), (zc, zf) = _coords(n_local, g, lower, width, device)
    R = _rotation(alpha, beta)
    w = _wave_table(wave_type, gamma, positive_vel, parallel_vel)

    def rot_fwd(x, y, z):
        return tuple(R[r][0] * x + R[r][1] * y + R[r][2] * z for r in range(3))

    def rot_inv(r0, r1, r2):
        return tuple(R[0][c] * r0 + R[1][c] * r1 + R[2][c] * r2 for c in range(3))

    def grid(x, y, z):     # broadcast 1-D coordinates to (z, y, x)
        return x.view(1, 1, -1), y.view(1, -1, 1), z.view(-1, 1, 1)

    f = {}
    if mhd:
        def potential(x, y, z):
            x0, x1, x2 = rot_fwd(*grid(x, y, z))
            ct = torch.cos(2.0 * _CELLO_PI * x0 / lam)
            r0 = x2 * amplitude * w["b1_ev"] * ct - x1 * amplitude * w["b2_ev"] * ct
            r1 = w["b_back"][2] * x0 + 0.0 * x1
            r2 = w["b_back"][0] * x1 - w["b_back"][1] * x0
            return rot_inv(r0, r1, r2)
        if w["has_a"]:
            Ax = potential(xc, yf, zf)[0].contiguous()       # (mz+1, my+1, mx)
            Ay = potential(xf, yc, zf)[1].contiguous()       # (mz+1, my, mx+1)
            Az = potential(xf, yf, zc)[2].contiguous()       # (mz, my+1, mx+1)
            dx, dy, dz = width
            f["bfieldi_x"] = ((Az[:, 1:, :] - Az[:, :-1, :]) / dy
                              - (Ay[1:, :, :] - Ay[:-1, :, :]) / dz)
            f["bfieldi_y"] = ((Ax[1:, :, :] - Ax[:-1, :, :]) / dz
                              - (Az[:, :, 1:] - Az[:, :, :-1]) / dx)
            f["bfieldi_z"] = ((Ay[:, :, 1:] - Ay[:, :, :-1]) / dx
                              - (Ax[:, 1:, :] - Ax[:, :-1, :]) / dy)
        else:
            mz, my, mx = zc.numel(), yc.numel(), xc.numel()
            for k, shp in (("x", (mz, my, mx + 1)), ("y", (mz, my + 1, mx)),
                           ("z", (mz + 1, my, mx))):
                f["bfieldi_" + k] = torch.zeros(shp, dtype=torch.float64, device=device)
        _center_b(f)
    x0, _, _ = rot_fwd(*grid(xc, yc, zc))
    trig = torch.cos(x0 * 2.0 * _CELLO_PI / lam)
    rho = w["rho_back"] + amplitude * w["rho_ev"] * trig
    mom = rot_inv(*(w["mom_back"][c] + amplitude * w["mom_ev"][c] * trig
                    for c in range(3)))
    f["density"] = rho
    for c, k in enumerate("xyz"):
        f["velocity_" + k] = mom[c] / rho
    f["total_energy"] = (w["etot_back"] + amplitude * w["etot_ev"] * trig) / rho
    if dual_energy:                                          # setup_eint_ :495-539
        ke = 0.5 * sum(f["velocity_" + k] ** 2 for k in "xyz")
        me = 0.5 * sum(f["bfield_" + k] ** 2 for k in "xyz") / rho if mhd else 0.0
        f["internal_energy"] = f["total_energy"] - ke - me
    f["pressure"] = torch.zeros_like(rho)
    return {k: v.contiguous() for k, v in f.items()}


def shock_tube(n_local, g, lower, width, setup="rj2a", aligned_ax=0, device="cuda",
               gamma=5.0 / 3.0, axis_velocity=0.0, mhd=True, dual_energy=False):
    """EnzoInitialShockTube (src/Enzo/initial/EnzoInitialShockTube.cpp:37-330):
    "rj2a" (Ryu & Jones 1995 fig. 2a) or "sod", discontinuity at 0.5 along
    `aligned_ax`, vector components permuted onto the tube's axis."""
    if setup == "rj2a":
        L = dict(rho=1.08, p=0.95, v=(1.2, 0.01, 0.5), b=(1.0155412503859613, 0.5641895835477563))
        Rt = dict(rho=1.0, p=1.0, v=(0.0, 0.0, 0.0), b=(1.1283791670955126, 0.5641895835477563))
        b0 = 0.5641895835477563
    elif setup == "sod":
        L = dict(rho=1.0, p=1.0, v=(0.0, 0.0, 0.0), b=(0.0, 0.0))
        Rt = dict(rho=0.125, p=0.1, v=(0.0, 0.0, 0.0), b=(0.0, 0.0))
        b0 = 0.0
    else:
        raise ValueError(f"unknown shock tube {setup!r}")
    coords = _coords(n_local, g, lower, width, device)
    mz, my, mx = (coords[2][0].numel(), coords[1][0].numel(), coords[0][0].numel())
    ia, ja, ka = aligned_ax, (aligned_ax + 1) % 3, (aligned_ax + 2) % 3
    # index of the first cell right of the discontinuity (cpp:196-206)
    shock = math.ceil((0.5 - lower[ia]) / width[ia] - 0.5 + g[ia])

    def side_mask(extent_along):       # True = left state, for an array whose
        idx = torch.arange(extent_along, device=device)       # extent along
        shape = [1, 1, 1]                                       # the tube is given
        shape[2 - ia] = -1
        return (idx < shock).view(shape)

    def pick(lv, rv, shape):
        m = side_mask(shape[2 - ia]).expand(shape)
        one = torch.ones(shape, dtype=torch.float64, device=device)
        return torch.where(m, lv * one, rv * one)

    cshape = (mz, my, mx)
    f = {"density": pick(L["rho"], Rt["rho"], cshape)}
    names = "xyz"
    vl = (L["v"][0] + axis_velocity, L["v"][1], L["v"][2])
    vr = (Rt["v"][0] + axis_velocity, Rt["v"][1], Rt["v"][2])
    for c, ax in enumerate((ia, ja, ka)):
        f["velocity_" + names[ax]] = pick(vl[c], vr[c], cshape)

    def energy(s, v):
        eint = s["p"] / ((gamma - 1.0) * s["rho"])
        v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
        b2 = b0 * b0 + s["b"][0] * s["b"][0] + s["b"][1] * s["b"][1]
        return eint, eint + 0.5 * (v2 + b2 / s["rho"])
    (el, tl), (er, tr) = energy(L, vl), energy(Rt, vr)
    f["total_energy"] = pick(tl, tr, cshape)
    if dual_energy:
        f["internal_energy"] = pick(el, er, cshape)
    if mhd:
        fshape = {0: (mz, my, mx + 1), 1: (mz, my + 1, mx), 2: (mz + 1, my, mx)}
        f["bfieldi_" + names[ia]] = torch.full(fshape[ia], b0, dtype=torch.float64,
                                               device=device)
        f["bfieldi_" + names[ja]] = pick(L["b"][0], Rt["b"][0], fshape[ja])
        f["bfieldi_" + names[ka]] = pick(L["b"][1], Rt["b"][1], fshape[ka])
        _center_b(f)
    f["pressure"] = torch.zeros(cshape, dtype=torch.float64, device=device)
    return {k: v.contiguous() for k, v in f.items()}


def cloud_perturbation_waves(nwaves, seed, min_lambda, max_lambda):
    """The plane waves of EnzoInitialCloud's density perturbation
    (WavePerturbation, src/Enzo/initial/EnzoInitialCloud.cpp:86-119): drawn on
    the host from std::minstd_rand(seed) exactly as the reference draws them
    (uniform_dist_transform_ :16-44, Box-Muller :51-59, sphere points :61-84;
    Python's math functions are the platform's libm, like the reference's).
    Returns [(kx, ky, kz, phi)]."""
    state = seed % 2147483647 or 1

    def rand():
        nonlocal state
        state = (state * 48271) % 2147483647
        return float(state)

    MAX = 2147483646.0

    def uniform(include_zero, include_one):
        raw = rand()
        if include_zero and include_one:
            rng, raw = MAX - 1.0, raw - 1.0
        elif include_zero:
            rng, raw = MAX, raw - 1.0
        elif include_one:
            rng = MAX
        else:
            rng = MAX + 1.0
        return raw / rng

    def normal_pair():
        x1 = uniform(False, False)
        x2 = uniform(False, False)
        coef = math.sqrt(-2.0 * math.log(x1))
        return coef * math.cos(2.0 * _CELLO_PI * x2), coef * math.sin(2.0 * _CELLO_PI * x2)

    waves = []
    for _ in range(nwaves):
        lam = min_lambda + (max_lambda - min_lambda) * uniform(True, True)
        while True:
            x, y = normal_pair()
            z, _unused = normal_pair()
            if x != 0.0 or y != 0.0 or z != 0.0:
                break
        mag = math.sqrt(x * x + y * y + z * z)
        kx = (x / mag) * 2 * _CELLO_PI / lam
        ky = (y / mag) * 2 * _CELLO_PI / lam
        kz = (z / mag) * 2 * _CELLO_PI / lam
        waves.append((kx, ky, kz, _CELLO_PI * uniform(True, False)))
    return waves


def _cloud_wave_average(waves, amplitude, xc, yc, zc, h):
    """WavePerturbation::operator() (cpp:135-160): the perturbation averaged
    over cells of widths h centred at (xc, yc, zc) (broadcastable tensors)"""
    total = torch.zeros(torch.broadcast_shapes(xc.shape, yc.shape, zc.shape),
                        dtype=torch.float64, device=xc.device)
    alpha = 8.0 * amplitude / (h[0] * h[1] * h[2])
    for kx, ky, kz, phi in waves:
        ci = (math.sin(kx * h[0] * 0.5) * math.sin(ky * h[1] * 0.5)
              * math.sin(kz * h[2] * 0.5)) / (kx * ky * kz)
        total = total + ci * torch.cos(kx * xc + ky * yc + kz * zc + phi)
    return alpha * total


def cloud(n_local, g, lower, width, subsample_n, cloud_radius, center,
          cloud_density, wind_density, wind_velocity, wind_total_energy,
          wind_internal_energy=0.0, device="cuda", mhd=False, dual_energy=True,
          bfield=(0.0, 0.0, 0.0), perturb=None):
    """EnzoInitialCloud (src/Enzo/initial/EnzoInitialCloud.cpp:606-748): a sphere
    of cloud_density at rest in a wind along +x, in pressure equilibrium. Cells
    cut by the sphere's surface get the volume-weighted density of their 2^n per
    axis sub-cells and a mass-weighted velocity (cpp:320-386). `bfield`: the
    uniform field the reference expects to find pre-initialised (cpp:452-523).
    `perturb` = (Nwaves, seed, amplitude, min_lambda, max_lambda): the optional
    density perturbation of the cloud (Initial:cloud:perturb_*, cpp:86-163,
    327-390): the wave parameters are drawn on the host exactly like the
    reference's, the cell / sub-cell averages are evaluated on the device (so
    they agree with the reference to the last bits of cos, not bit for bit)."""
    f64 = dict(dtype=torch.float64, device=device)
    m = [n_local[a] + 2 * g[a] for a in range(3)]
    sqr_radius = cloud_radius * cloud_radius
    nsub = 2 ** subsample_n
    # prep_subcell_offsets_ (cpp:246-256)
    offs = []
    for a in range(3):
        cur, o = 1.0 / 2 ** (subsample_n + 1), []
        o.append(cur * width[a])
        for _ in range(1, nsub):
            cur += 1.0 / 2 ** subsample_n
            o.append(cur * width[a])
        offs.append(o)

    def along(a, v):                       # 1-D array along axis a -> (z, y, x)
        shape = [1, 1, 1]
        shape[2 - a] = -1
        return v.view(shape)

    near2, far2, sub2 = [], [], []
    for a in range(3):
        # Data::field_cell_faces (Cello/data_Data.cpp:91-121): xm + ix * hx
        idx = torch.arange(m[a] + 1, **f64) - g[a]
        face = lower[a] + idx * width[a]
        left, right, c = face[:-1], face[1:], center[a]
        # SphereRegion::check_intersect (cpp:204-237)
        inside_far = torch.where((c - left) > (right - c), left, right)
        nearest = torch.where(c <= left, left,
                              torch.where(c >= right, right, torch.full_like(left, c)))
        furthest = torch.where(c <= left, right,
                               torch.where(c >= right, left, inside_far))
        dn, df = nearest - c, furthest - c
        near2.append(along(a, dn * dn))
        far2.append(along(a, df * df))
        sub = []
        for o in offs[a]:
            ds = (left + o) - c
            sub.append(along(a, ds * ds))
        sub2.append(sub)
    enclosed = ((far2[0] + far2[1]) + far2[2]) <= sqr_radius
    overlap = ((near2[0] + near2[1]) + near2[2]) <= sqr_radius
    count = torch.zeros((m[2], m[1], m[0]), dtype=torch.int32, device=device)
    for sz in sub2[2]:
        for sy in sub2[1]:
            for sx in sub2[0]:
                count += (((sx + sy) + sz) <= sqr_radius).to(torch.int32)
    frac = count.to(torch.float64) / float(nsub ** 3)
    frac = torch.where(enclosed, torch.ones_like(frac),
                       torch.where(overlap, frac, torch.zeros_like(frac)))

    perturbation = torch.zeros_like(frac)
    if perturb is not None and perturb[0] > 0 and perturb[2] > 0.0:
        nw, seed, amplitude, lmin, lmax = perturb
        waves = cloud_perturbation_waves(int(nw), int(seed), lmin, lmax)
        # enclosed cells: the average over the whole cell, at its centre
        cen = []
        for a in range(3):
            idx = torch.arange(m[a] + 1, **f64) - g[a]
            face = lower[a] + idx * width[a]
            cen.append(along(a, 0.5 * (face[:-1] + face[1:])))
        whole = _cloud_wave_average(waves, amplitude, cen[0], cen[1], cen[2], width)
        perturbation = torch.where(enclosed, whole, perturbation)
        # cut cells: the mean over the enclosed sub-cells (z, y, x order)
        partial = overlap & ~enclosed
        pidx = torch.nonzero(partial, as_tuple=True)
        if pidx[0].numel() > 0:
            left = []
            for a in range(3):
                idx = torch.arange(m[a] + 1, **f64) - g[a]
                face = lower[a] + idx * width[a]
                left.append(face[:-1][pidx[2 - a]])
            sub_h = [width[a] / float(nsub) for a in range(3)]
            psum = torch.zeros_like(left[0])
            for oz in offs[2]:
                sub_zc = left[2] + oz
                for oy in offs[1]:
                    sub_yc = left[1] + oy
                    for ox in offs[0]:
                        sub_xc = left[0] + ox
                        dx, dy, dz = sub_xc - center[0], sub_yc - center[1], sub_zc - center[2]
                        inside = ((dx * dx + dy * dy) + dz * dz) <= sqr_radius
                        val = _cloud_wave_average(waves, amplitude, sub_xc, sub_yc, sub_zc,
                                                  sub_h)
                        psum = psum + torch.where(inside, val, torch.zeros_like(val))
            cnt = count[pidx].to(torch.float64)
            mean = torch.where(cnt > 0, psum / torch.clamp(cnt, min=1.0),
                               torch.zeros_like(psum))
            perturbation = perturbation.clone()
            perturbation[pidx] = mean
    perturbation = perturbation + 1.0
    avg_density = frac * cloud_density * perturbation + (1.0 - frac) * wind_density
    ratio = torch.full_like(avg_density, wind_density) / avg_density   # true division
    wind_mass_weight = (1.0 - frac) * ratio
    f = {"density": avg_density,
         "velocity_x": wind_mass_weight * wind_velocity,
         "velocity_y": torch.zeros_like(frac), "velocity_z": torch.zeros_like(frac)}
    magnetic_edens = 0.0
    if mhd:
        magnetic_edens = 0.5 * (bfield[0] * bfield[0] + bfield[1] * bfield[1]
                                + bfield[2] * bfield[2])
    if dual_energy:
        f["internal_energy"] = wind_internal_energy * ratio
        eint_density = wind_internal_energy * wind_density
    else:
        eint_density = ((wind_total_energy - 0.5 * wind_velocity * wind_velocity)
                        * wind_density - magnetic_edens)
    vx = f["velocity_x"]
    etot = (torch.full_like(frac, eint_density + magnetic_edens) / avg_density
            + 0.5 * vx * vx)
    f["total_energy"] = torch.where(frac == 0, torch.full_like(frac, wind_total_energy),
                                    etot)
    if mhd:
        names = "xyz"
        for a in range(3):
            f["bfield_" + names[a]] = torch.full_like(frac, bfield[a])
            shape = [m[2], m[1], m[0]]
            shape[2 - a] += 1
            f["bfieldi_" + names[a]] = torch.full(shape, bfield[a], **f64)
    f["pressure"] = torch.zeros_like(frac)
    return {k: v.contiguous() for k, v in f.items()}
