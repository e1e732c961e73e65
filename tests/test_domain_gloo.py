"""CPU, world_size 2 over gloo: the unigrid driver's neighbour / message-ordering
logic (enzo-e_b200/domain.py). The CUDA pack/unpack/wrap kernels are replaced
by numpy stand-ins with the same slab convention as csrc (vlct_api.cu
halo_copy); the GPU versions are checked against these same stand-ins in
test_gpu_domain.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

FIELDS = ["density", "velocity_x", "bfield_x", "bfieldi_x", "bfieldi_y",
          "bfieldi_z"]


def face_axis(name):
    return {"bfieldi_x": 0, "bfieldi_y": 1, "bfieldi_z": 2}.get(name, -1)


class HostBlock:
    def __init__(self, fields, n, g):
        self.fields, self.n, self.g = fields, n, g


class HostKernels:
    """numpy restatement of k_slab_copy / k_wrap_axis index conventions"""

    def halo_bytes(self, blk, axis):
        total = 0
        for name in FIELDS:
            a = blk.fields[name]
            shp = list(a.shape[::-1])        # (x, y, z) extents
            shp[axis] = blk.g[axis]
            total += shp[0] * shp[1] * shp[2]
        return total * 8

    @staticmethod
    def _slab(a, axis, lo, width):
        sl = [slice(None)] * 3
        sl[2 - axis] = slice(lo, lo + width)
        return tuple(sl)

    def halo_pack(self, blk, axis, side, buf):
        off = 0
        n, g = blk.n[axis], blk.g[axis]
        for name in FIELDS:
            a = blk.fields[name]
            cen = 1 if face_axis(name) == axis else 0
            lo = (g + cen) if side == 0 else n
            s = a[self._slab(a, axis, lo, g)]
            buf[off:off + s.size] = torch.from_numpy(np.ascontiguousarray(s).ravel())
            off += s.size

    def halo_unpack(self, blk, axis, side, buf):
        off = 0
        n, g = blk.n[axis], blk.g[axis]
        for name in FIELDS:
            a = blk.fields[name]
            cen = 1 if face_axis(name) == axis else 0
            lo = 0 if side == 0 else g + n + cen
            view = a[self._slab(a, axis, lo, g)]
            view[...] = buf[off:off + view.size].numpy().reshape(view.shape)
            off += view.size

    def wrap(self, blk, axes):
        for axis in range(3):
            if not axes & (1 << axis):
                continue
            n, g = blk.n[axis], blk.g[axis]
            for name in FIELDS:
                a = blk.fields[name]
                cen = 1 if face_axis(name) == axis else 0
                a[self._slab(a, axis, 0, g)] = a[self._slab(a, axis, n, g)]
                a[self._slab(a, axis, g + n + cen, g)] = \
                    a[self._slab(a, axis, g + cen, g)]


    # -- domain boundaries (EnzoBoundary.cpp:164-283,352-466 /
    #    problem_BoundaryValue.cpp:131-273; the GPU kernels are checked against
    #    the oracle's restatement of the same rules in test_gpu_domain.py) ------
    def boundary(self, blk, axis, side, kind):
        n, g = blk.n[axis], blk.g[axis]
        for name in FIELDS:
            a = np.moveaxis(blk.fields[name], 2 - axis, 0)     # a view
            cen = 1 if face_axis(name) == axis else 0
            sign = -1.0 if (kind == "reflecting" and name in (
                "velocity_" + "xyz"[axis], "bfield_" + "xyz"[axis],
                "bfieldi_" + "xyz"[axis])) else 1.0
            for ig in range(g):
                if kind == "outflow":
                    src = g if side == 0 else n + g - 1 + cen
                    dst = g - ig - 1 if side == 0 else src + ig + 1
                else:
                    src = g + cen + ig if side == 0 else n + g - 1 - ig
                    dst = g - ig - 1 if side == 0 else n + g + ig + cen
                a[dst] = sign * a[src]

    def boundary_inflow(self, blk, axis, side, values, passive=()):
        g = blk.g[axis]
        for name, val in values.items():
            if name not in blk.fields:
                continue
            a = np.moveaxis(blk.fields[name], 2 - axis, 0)
            if side == 0:
                a[:g] = val
            else:
                a[a.shape[0] - g:] = val


def global_value(name, ix, iy, iz, N):
    """a periodic integer-valued function of the GLOBAL index (faces share
    the index of the cell on their upper side)"""
    h = {"density": 1, "velocity_x": 2, "bfield_x": 3, "bfieldi_x": 4,
         "bfieldi_y": 5, "bfieldi_z": 6}[name]
    return (h * 1000003 + (ix % N[0]) * 10007 + (iy % N[1]) * 101
            + (iz % N[2])).astype(np.float64)


def make_block(coords, n, g, N, fill_ghosts):
    from enzo_e_b200 import abi
    f = {}
    for name in FIELDS:
        shp = abi.field_shape(name, *n, *g)
        iz, iy, ix = np.meshgrid(np.arange(shp[0]), np.arange(shp[1]),
                                 np.arange(shp[2]), indexing="ij")
        gi = [ix - g[0] + coords[0] * n[0], iy - g[1] + coords[1] * n[1],
              iz - g[2] + coords[2] * n[2]]
        full = global_value(name, gi[0], gi[1], gi[2], N)
        if fill_ghosts:
            f[name] = full
        else:
            a = np.full(shp, -1.0)
            fa = face_axis(name)
            sl = tuple(slice(g[2 - k], g[2 - k] + n[2 - k] + (1 if fa == 2 - k else 0))
                       for k in range(3))
            a[sl] = full[sl]
            f[name] = a
    return HostBlock(f, n, g)


def _worker(rank, world, port, grid, result_q, defer=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from enzo_e_b200.domain import Domain
        dom = Domain(rank, world, grid=grid)
        n, g = (6, 5, 4), (3, 3, 3)
        N = tuple(n[a] * grid[a] for a in range(3))
        blk = make_block(dom.coords, n, g, N, fill_ghosts=False)
        want = make_block(dom.coords, n, g, N, fill_ghosts=True)
        k = HostKernels()
        pending = dom.refresh(k, blk, pack=k.halo_pack, unpack=k.halo_unpack,
                              wrap=k.wrap, defer_z=defer,
                              alloc=lambda nbytes: torch.empty(
                                  nbytes // 8, dtype=torch.float64))
        # a deferred z exchange is handed back iff the domain is split along z
        assert (pending is not None) == (defer and grid[2] > 1)
        if pending is not None:
            # x and y are complete, the z ghosts arrive with refresh_finish
            assert np.all(blk.fields["density"][:g[2]] == -1.0)
            dom.refresh_finish(k, blk, pending)
        bad = [name for name in FIELDS
               if not np.array_equal(blk.fields[name], want.fields[name])]
        dt = dom.global_dt(0.5 + rank, device="cpu")
        result_q.put((rank, bad, dt))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("grid,defer", [((1, 1, 2), False), ((2, 1, 1), False),
                                        ((1, 2, 1), False), ((1, 1, 2), True),
                                        ((1, 2, 1), True)])
def test_refresh_two_ranks(grid, defer):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, q, defer))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad, dt in results:
        assert not bad, f"rank {rank}: ghost zones wrong in {bad}"
        assert dt == 0.5


def test_proc_grid_and_neighbours():
    from enzo_e_b200.domain import Domain, proc_grid
    assert proc_grid(1) == (1, 1, 1)
    assert proc_grid(2) == (1, 1, 2)
    assert proc_grid(4) == (1, 2, 2)
    assert proc_grid(8) == (2, 2, 2)
    assert proc_grid(8, slabs=True) == (1, 1, 8)
    for world in (2, 4, 8, 6, 12):
        g = proc_grid(world)
        assert g[0] * g[1] * g[2] == world
        seen = set()
        for r in range(world):
            d = Domain(r, world)
            seen.add(d.coords)
            for axis in range(3):
                up = d.neighbor(axis, +1)
                assert Domain(up, world).neighbor(axis, -1) == r
        assert len(seen) == world


def test_single_rank_refresh_wraps():
    from enzo_e_b200.domain import Domain
    dom = Domain(0, 1)
    n, g = (6, 5, 4), (3, 3, 3)
    blk = make_block((0, 0, 0), n, g, n, fill_ghosts=False)
    want = make_block((0, 0, 0), n, g, n, fill_ghosts=True)
    k = HostKernels()
    dom.refresh(k, blk, pack=k.halo_pack, unpack=k.halo_unpack, wrap=k.wrap,
                alloc=lambda nbytes: torch.empty(nbytes // 8, dtype=torch.float64))
    for name in FIELDS:
        assert np.array_equal(blk.fields[name], want.fields[name]), name


# ---------------------------------------------------------------------------
# non-periodic domains: no exchange across a domain face, then every Boundary
# object in list order (Block::update_boundary_)
# ---------------------------------------------------------------------------
CLOUD_LIKE = [dict(type="outflow", axis=0, face=1),
              dict(type="inflow", axis=0, face=0,
                   values={"density": 0.5, "velocity_x": 2.0}),
              dict(type="outflow", axis=1), dict(type="reflecting", axis=2),
              dict(type="inflow", axis=0, face=0, values={"bfieldi_y": 0.0,
                                                          "bfield_x": 0.0})]
TUBE_X = [dict(type="outflow", axis=0)]          # periodic across the tube


def expected_global(boundaries, n, g, grid):
    """the same rules on ONE block that covers the whole domain"""
    from enzo_e_b200.domain import Domain
    N = tuple(n[a] * grid[a] for a in range(3))
    blk = make_block((0, 0, 0), N, g, N, fill_ghosts=False)
    k = HostKernels()
    Domain(0, 1, boundaries=boundaries).refresh(
        k, blk, pack=k.halo_pack, unpack=k.halo_unpack, wrap=k.wrap,
        boundary=k.boundary, boundary_inflow=k.boundary_inflow,
        alloc=lambda nbytes: torch.empty(nbytes // 8, dtype=torch.float64))
    return blk


def local_part(glob, coords, n, g):
    """the levels a brick (ghost zones included) occupies in the global block"""
    out = {}
    for name, a in glob.fields.items():
        fa = face_axis(name)
        sl = tuple(slice(coords[2 - k] * n[2 - k],
                         coords[2 - k] * n[2 - k] + n[2 - k] + 2 * g[2 - k]
                         + (1 if fa == 2 - k else 0)) for k in range(3))
        out[name] = a[sl]
    return out


def _worker_bc(rank, world, port, grid, which, defer, result_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from enzo_e_b200.domain import Domain
        boundaries = {"cloud": CLOUD_LIKE, "tube": TUBE_X}[which]
        dom = Domain(rank, world, grid=grid, boundaries=boundaries)
        n, g = (6, 5, 4), (3, 3, 3)
        N = tuple(n[a] * grid[a] for a in range(3))
        blk = make_block(dom.coords, n, g, N, fill_ghosts=False)
        want = local_part(expected_global(boundaries, n, g, grid), dom.coords, n, g)
        k = HostKernels()
        kw = dict(pack=k.halo_pack, unpack=k.halo_unpack, wrap=k.wrap,
                  boundary=k.boundary, boundary_inflow=k.boundary_inflow,
                  alloc=lambda nbytes: torch.empty(nbytes // 8, dtype=torch.float64))
        pending = dom.refresh(k, blk, defer_z=defer, **kw)
        assert (pending is not None) == (defer and grid[2] > 1)
        if pending is not None:
            dom.refresh_finish(k, blk, pending)
        bad = [name for name in FIELDS
               if not np.array_equal(blk.fields[name], want[name])]
        result_q.put((rank, bad))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grid,which,defer", [
    ((2, 1, 1), "cloud", False), ((1, 2, 1), "cloud", False),
    ((1, 1, 2), "cloud", False), ((1, 1, 2), "cloud", True),
    ((2, 1, 1), "tube", False), ((1, 1, 2), "tube", True)])
def test_non_periodic_domain_two_ranks(grid, which, defer):
    """two bricks of a domain with inflow / outflow / reflecting faces get the
    ghost zones one block covering the whole domain gets"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_bc,
                         args=(r, world, port, grid, which, defer, q))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad in results:
        assert not bad, f"rank {rank}: ghost zones wrong in {bad}"


def test_single_block_boundaries_follow_list_order():
    """one brick: the refresh of a non-periodic domain = wrap of the periodic
    axes, then every Boundary object in list order"""
    n, g = (6, 5, 4), (3, 3, 3)
    got = expected_global(CLOUD_LIKE, n, g, (1, 1, 1))
    want = make_block((0, 0, 0), n, g, n, fill_ghosts=False)
    k = HostKernels()
    k.boundary(want, 0, 1, "outflow")
    k.boundary_inflow(want, 0, 0, CLOUD_LIKE[1]["values"])
    for side in (0, 1):
        k.boundary(want, 1, side, "outflow")
    for side in (0, 1):
        k.boundary(want, 2, side, "reflecting")
    k.boundary_inflow(want, 0, 0, CLOUD_LIKE[4]["values"])
    for name in FIELDS:
        assert np.array_equal(got.fields[name], want.fields[name]), name
    # an axis without a boundary object stays periodic
    got = expected_global(TUBE_X, n, g, (1, 1, 1))
    a = got.fields["density"]
    assert np.array_equal(a[:3], a[4:7]) and np.array_equal(a[:, :3], a[:, 5:8])
    assert np.array_equal(a[:, :, 0], a[:, :, 3]) and np.array_equal(a[:, :, 2], a[:, :, 3])
