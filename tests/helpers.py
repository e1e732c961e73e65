"""Shared helpers for the test-suite (host side only, numpy)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

from enzo_e_b200 import abi  # noqa: E402
import oracle  # noqa: E402


def make_config(riemann="hlld", recon="plm", theta=1.5, mhd=True,
                time_scheme="vl", courant=0.3, gamma=5.0 / 3.0,
                dual_energy=False, eta=0.001, dfloor=1e-200, pfloor=1e-200,
                n_passive=0, accel=False):
    cfg = abi.default_config()
    cfg.riemann_solver = abi.RIEMANN[riemann]
    cfg.reconstruct_method = abi.RECON[recon]
    cfg.theta_limiter = theta
    cfg.mhd_choice = 1 if mhd else 0
    cfg.time_scheme = abi.TIME_SCHEME[time_scheme]
    cfg.courant = courant
    cfg.gamma = gamma
    cfg.dual_energy = 1 if dual_energy else 0
    cfg.dual_energy_eta = eta
    cfg.density_floor = dfloor
    cfg.pressure_floor = pfloor
    cfg.n_passive = n_passive
    cfg.has_acceleration = 1 if accel else 0
    return cfg


def field_names(cfg):
    names = ["density", "velocity_x", "velocity_y", "velocity_z",
             "total_energy"]
    if cfg.dual_energy:
        names.append("internal_energy")
    if cfg.mhd_choice == 1:
        names += ["bfield_x", "bfield_y", "bfield_z",
                  "bfieldi_x", "bfieldi_y", "bfieldi_z"]
    names.append("pressure")
    if cfg.has_acceleration:
        names += ["acceleration_x", "acceleration_y", "acceleration_z"]
    return names


def passive_names(cfg):
    return [f"passive_{i}" for i in range(cfg.n_passive)]


def random_state(cfg, n, g, seed=0, amp=0.1, smooth=True):
    """A seeded, physically admissible random state (ghost zones included).

    Face-centred B is random; the cell-centred B is its face average, as the
    reference's initialisers guarantee (EnzoInitialBCenter)."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = n
    gx, gy, gz = g
    f = {}

    def noise(shape):
        a = rng.standard_normal(shape)
        if smooth:  # correlate neighbours a little so that slopes are mixed
            for ax in range(3):
                a = 0.5 * a + 0.25 * (np.roll(a, 1, ax) + np.roll(a, -1, ax))
        return a

    cshape = abi.field_shape("density", nx, ny, nz, gx, gy, gz)
    f["density"] = 1.0 + amp * noise(cshape)
    for k in "xyz":
        f["velocity_" + k] = 0.5 * amp * 10 * noise(cshape)
    p = 0.6 * (1.0 + amp * noise(cshape))
    ke = 0.5 * sum(f["velocity_" + k] ** 2 for k in "xyz")
    me = 0.0
    if cfg.mhd_choice == 1:
        for k in "xyz":
            shp = abi.field_shape("bfieldi_" + k, nx, ny, nz, gx, gy, gz)
            f["bfieldi_" + k] = {"x": 1.0, "y": 0.5, "z": -0.3}[k] \
                + amp * noise(shp)
        f["bfield_x"] = 0.5 * (f["bfieldi_x"][:, :, :-1] + f["bfieldi_x"][:, :, 1:])
        f["bfield_y"] = 0.5 * (f["bfieldi_y"][:, :-1, :] + f["bfieldi_y"][:, 1:, :])
        f["bfield_z"] = 0.5 * (f["bfieldi_z"][:-1, :, :] + f["bfieldi_z"][1:, :, :])
        me = 0.5 * sum(f["bfield_" + k] ** 2 for k in "xyz") / f["density"]
    eint = p / ((cfg.gamma - 1.0) * f["density"])
    f["total_energy"] = eint + ke + me
    if cfg.dual_energy:
        f["internal_energy"] = eint * (1.0 + 1e-3 * noise(cshape))
    f["pressure"] = np.zeros(cshape)
    if cfg.has_acceleration:
        for k in "xyz":
            f["acceleration_" + k] = 0.3 * noise(cshape)
    for i, name in enumerate(passive_names(cfg)):
        f[name] = f["density"] * (0.5 + 0.4 * np.sin(1.0 + i + noise(cshape)))
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in f.items()}


def copy_state(f):
    return {k: v.copy() for k, v in f.items()}


def max_abs_diff(a, b):
    return {k: float(np.max(np.abs(a[k] - b[k]))) for k in a}


def bit_equal(a, b):
    """dict of booleans: arrays identical bit for bit (NaN-safe)."""
    return {k: bool(np.array_equal(a[k].view(np.uint64), b[k].view(np.uint64)))
            for k in a}
