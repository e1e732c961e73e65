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
