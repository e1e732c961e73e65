"""GPU: the device ghost-zone kernels (periodic wrap, halo pack/unpack) against
the numpy stand-ins that the gloo test (test_domain_gloo.py) uses."""
import numpy as np
import pytest
import torch

from test_domain_gloo import FIELDS, HostBlock, HostKernels, make_block

pytestmark = pytest.mark.gpu


def _gpu_objects(n, g, fields):
    from helpers import make_config
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(riemann="hlld", recon="plm", mhd=True)
    method = EnzoMethodMHDVlct(config=cfg)
    full = {}
    for name in ("density", "velocity_x", "velocity_y", "velocity_z",
                 "total_energy", "bfield_x", "bfield_y", "bfield_z",
                 "bfieldi_x", "bfieldi_y", "bfieldi_z", "pressure"):
        if name in fields:
            full[name] = torch.from_numpy(fields[name]).cuda()
        else:
            shp = fields["density"].shape
            full[name] = torch.zeros(shp, dtype=torch.float64, device="cuda")
    block = Block(full, n, g, (0.1, 0.1, 0.1))
    return method, block, full


def test_periodic_wrap_matches_host():
    n, g = (6, 5, 4), (3, 3, 3)
    blk = make_block((0, 0, 0), n, g, n, fill_ghosts=False)
    want = make_block((0, 0, 0), n, g, n, fill_ghosts=True)
    method, block, dev = _gpu_objects(n, g, blk.fields)
    method.refresh_periodic(block, 7)
    method.synchronize()
    for name in FIELDS:
        assert np.array_equal(dev[name].cpu().numpy(), want.fields[name]), name
    method.close()


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_halo_pack_unpack_match_host(axis):
    n, g = (7, 6, 5), (3, 3, 3)
    rng = np.random.default_rng(3)
    blk = make_block((0, 0, 0), n, g, n, fill_ghosts=True)
    for a in blk.fields.values():
        a += rng.standard_normal(a.shape)
    method, block, dev = _gpu_objects(n, g, blk.fields)
    # the device block carries 12 fields, the host stand-in only FIELDS: compare
    # field by field through single-field host packs
    k = HostKernels()
    nbytes = method.halo_bytes(block, axis)
    buf = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
    for side in (0, 1):
        method.halo_pack(block, axis, side, buf)
        method.synchronize()
        got = buf.cpu().numpy()
        # expected layout: fields in C-ABI order, each slab C-ordered
        order = ["density", "velocity_x", "velocity_y", "velocity_z",
                 "total_energy", "bfield_x", "bfield_y", "bfield_z",
                 "bfieldi_x", "bfieldi_y", "bfieldi_z", "pressure"]
        off = 0
        for name in order:
            arr = dev[name].cpu().numpy()
            cen = 1 if {"bfieldi_x": 0, "bfieldi_y": 1, "bfieldi_z": 2}.get(name, -1) == axis else 0
            lo = (g[axis] + cen) if side == 0 else n[axis]
            slab = arr[HostKernels._slab(arr, axis, lo, g[axis])]
            assert np.array_equal(got[off:off + slab.size], slab.ravel()), (name, side)
            off += slab.size
        assert off == got.size
    # unpack(pack) of the opposite side == periodic wrap along that axis
    ref = {name: dev[name].clone() for name in dev}
    method.refresh_periodic(block, 1 << axis)
    method.synchronize()
    wrapped = {name: dev[name].clone() for name in dev}
    for name in dev:
        dev[name].copy_(ref[name])
    lo_buf, hi_buf = torch.empty_like(buf), torch.empty_like(buf)
    method.halo_pack(block, axis, 0, lo_buf)   # my low active layers -> upper ghosts
    method.halo_pack(block, axis, 1, hi_buf)   # my high active layers -> lower ghosts
    method.halo_unpack(block, axis, 1, lo_buf)
    method.halo_unpack(block, axis, 0, hi_buf)
    method.synchronize()
    for name in dev:
        assert torch.equal(dev[name], wrapped[name]), name
    method.close()


@pytest.mark.parametrize("kind", ["outflow", "reflecting"])
def test_boundary_kernels_match_oracle(kind):
    """vlct_boundary against the oracle's restatement of EnzoBoundary, every
    axis and side, all fields (cell- and face-centred, passive scalars)"""
    from helpers import make_config, random_state, copy_state, bit_equal, oracle
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(riemann="hlld", recon="plm", mhd=True, dual_energy=True,
                      n_passive=2, accel=True)
    n, g, d = (9, 5, 7), (3, 3, 3), (0.1, 0.1, 0.1)
    passive = [f"passive_{i}" for i in range(2)]
    host = random_state(cfg, n, g, seed=41)
    method = EnzoMethodMHDVlct(config=cfg)
    for axis in range(3):
        for side in (0, 1):
            want = copy_state(host)
            blk = oracle.numpy_block(want, n, g, d, passive)
            oracle.boundary(blk, axis, side, kind, n_passive=2)
            dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
            block = Block(dev, n, g, d, passive=passive)
            method.boundary(block, axis, side, kind)
            method.synchronize()
            got = {k: v.cpu().numpy() for k, v in dev.items()}
            eq = bit_equal(want, got)
            assert all(eq.values()), (axis, side, [k for k, v in eq.items() if not v])
    method.close()


def test_inflow_boundary_kernel_matches_oracle():
    """vlct_boundary_inflow against the oracle's restatement of
    BoundaryValue::enforce: every axis and side, a field list that mixes cell-
    and face-centred fields and one of two passive scalars"""
    from helpers import make_config, random_state, copy_state, bit_equal, oracle
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(riemann="hlld", recon="plm", mhd=True, dual_energy=True,
                      n_passive=2)
    n, g, d = (9, 5, 7), (3, 3, 3), (0.1, 0.1, 0.1)
    passive = [f"passive_{i}" for i in range(2)]
    values = {"density": 0.25, "velocity_x": 1.5, "total_energy": 9.0,
              "internal_energy": 7.0, "bfieldi_x": -2.0, "bfieldi_z": 3.0,
              "bfield_y": 0.0}
    host = random_state(cfg, n, g, seed=43)
    method = EnzoMethodMHDVlct(config=cfg)
    for axis in range(3):
        for side in (0, 1):
            want = copy_state(host)
            blk = oracle.numpy_block(want, n, g, d, passive)
            oracle.boundary_inflow(blk, axis, side, values, passive=(None, 0.5),
                                   n_passive=2)
            dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
            block = Block(dev, n, g, d, passive=passive)
            method.boundary_inflow(block, axis, side, values, passive=(None, 0.5))
            method.synchronize()
            got = {k: v.cpu().numpy() for k, v in dev.items()}
            eq = bit_equal(want, got)
            assert all(eq.values()), (axis, side, [k for k, v in eq.items() if not v])
            changed = [k for k in host if not np.array_equal(host[k], got[k])]
            assert sorted(changed) == sorted(list(values) + ["passive_1"])
    method.close()
