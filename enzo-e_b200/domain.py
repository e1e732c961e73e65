"""Unigrid driver: one block per GPU, periodic domain, ghost exchange.

Stands in for the pieces of Cello that surround Method::compute on a unigrid
(SURVEY 3.2-3.3): the refresh phase (src/Cello/control_refresh.cpp:243-359) and
the min-reduction of the timestep (src/Cello/control_stopping.cpp:96-142).

  * the domain N^3 is cut into px*py*pz bricks, one per rank/GPU
    (`Mesh:root_blocks`); every rank owns one `Block`
  * refresh = for axis in x, y, z: exchange the g outermost active layers with
    the two neighbours along that axis (slabs span the full ghost-including
    extent of the other axes, so edges and corners arrive after the third
    axis: 6 messages instead of 26). Along an axis with a single brick the
    block wraps onto itself on the device.
  * slabs are packed/unpacked by CUDA kernels (csrc: k_slab_copy) and moved
    with NCCL send/recv (torch.distributed P2P) over NVLink
  * dt = all_reduce(MIN) of the per-block timestep
  * `step()` overlaps the exchange along z (the last axis of the refresh) with
    the update: the slabs are packed and handed to NCCL, the interior part of
    the step (everything that does not read a z ghost level,
    vlct_compute_dev_part) is queued behind them on the compute stream while
    NCCL moves the slabs on its own stream, then the ghosts are unpacked and
    the lower / upper rest of the step runs. Results are bit-identical to
    refresh() + compute().
"""
import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def proc_grid(world, slabs=False):
    """(px, py, pz): split z first (contiguous slabs), then y, then x.
    slabs=True: z slabs only (two neighbours per rank, the whole exchange can
    be overlapped with the update)."""
    if slabs:
        return (1, 1, world)
    grid = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
    if world in grid:
        return grid[world]
    # generic fallback: factor into three near-equal parts, largest on z
    f = [1, 1, 1]
    n, p = world, 2
    facs = []
    while n > 1:
        while n % p == 0:
            facs.append(p)
            n //= p
        p += 1
    for q in sorted(facs, reverse=True):
        f[f.index(min(f))] *= q
    f.sort()
    return tuple(f)


class Domain:
    def __init__(self, rank=0, world=1, grid=None):
        self.rank, self.world = rank, world
        self.grid = tuple(grid) if grid else proc_grid(world)
        px, py, pz = self.grid
        assert px * py * pz == world
        # rank = (cz * py + cy) * px + cx
        self.coords = (rank % px, (rank // px) % py, rank // (px * py))

    def neighbor(self, axis, direction):
        c = list(self.coords)
        c[axis] = (c[axis] + direction) % self.grid[axis]
        px, py, _ = self.grid
        return (c[2] * py + c[1]) * px + c[0]

    def lower_corner(self, n_local, width, origin=(0.0, 0.0, 0.0)):
        return tuple(origin[a] + self.coords[a] * n_local[a] * width[a]
                     for a in range(3))

    # -- refresh ---------------------------------------------------------------
    def refresh(self, method, block, buffers=None, pack=None, unpack=None,
                wrap=None, alloc=None, defer_z=False):
        """Fill all ghost zones of `block` (periodic domain).

        pack/unpack/wrap/alloc default to the CUDA kernels behind `method`;
        the gloo CPU tests pass host stand-ins to exercise the neighbour and
        message-ordering logic without a GPU.

        defer_z=True: if the domain is split along z, the z exchange is only
        started (slabs packed, sends / receives posted); the returned handle
        goes to refresh_finish(), which waits for the messages and unpacks the
        ghosts. Returns None when nothing was deferred."""
        pack = pack or method.halo_pack
        unpack = unpack or method.halo_unpack
        wrap = wrap or (lambda blk, axes: method.refresh_periodic(blk, axes))
        if alloc is None:
            def alloc(nbytes):
                return torch.empty(nbytes // 8, dtype=torch.float64,
                                   device="cuda")
        if buffers is None:
            buffers = self.__dict__.setdefault("_buffers", {})
        for axis in range(3):
            if self.grid[axis] == 1:
                wrap(block, 1 << axis)
                continue
            key = axis
            if key not in buffers:
                nbytes = method.halo_bytes(block, axis)
                buffers[key] = [alloc(nbytes) for _ in range(4)]
            send_lo, send_hi, recv_lo, recv_hi = buffers[key]
            pack(block, axis, 0, send_lo)
            pack(block, axis, 1, send_hi)
            lo, hi = self.neighbor(axis, -1), self.neighbor(axis, +1)
            # If the block's kernels run on torch's current stream (bench.py
            # does that), packs, NCCL and unpacks are stream-ordered and no
            # host synchronisation is needed; otherwise fence explicitly.
            same_stream = getattr(block, "stream_is_current", False)
            if send_lo.is_cuda and not same_stream:
                method.synchronize()
            ops = [dist.P2POp(dist.isend, send_lo, lo),
                   dist.P2POp(dist.isend, send_hi, hi),
                   dist.P2POp(dist.irecv, recv_hi, hi),
                   dist.P2POp(dist.irecv, recv_lo, lo)]
            reqs = dist.batch_isend_irecv(ops)
            pending = (reqs, axis, recv_lo, recv_hi, unpack, same_stream)
            if defer_z and axis == 2:
                # NCCL works on its own stream, ordered after the packs; the
                # caller's stream only joins it in refresh_finish()
                return pending
            self.refresh_finish(method, block, pending)
        return None

    def refresh_finish(self, method, block, pending):
        """Wait for a deferred exchange and unpack its ghost slabs."""
        reqs, axis, recv_lo, recv_hi, unpack, same_stream = pending
        for req in reqs:
            req.wait()
        if recv_lo.is_cuda and not same_stream:
            torch.cuda.current_stream().synchronize()
        unpack(block, axis, 0, recv_lo)
        unpack(block, axis, 1, recv_hi)

    def step(self, method, block, dt, overlap=True):
        """refresh + compute of one cycle (dt: device-resident, already
        reduced over ranks). With overlap the z exchange runs under the
        interior part of the update."""
        from . import abi
        mz = block.n[2] + 2 * block.g[2]
        z_lo = block.g[2] + abi.PART_REACH_BELOW
        z_hi = mz - block.g[2] - abi.PART_REACH_ABOVE
        can_split = (overlap and self.grid[2] > 1 and z_hi > z_lo
                     and method.config.time_scheme == abi.TIME_SCHEME["vl"]
                     and hasattr(dt, "data_ptr")
                     and getattr(block, "stream_is_current", False))
        pending = self.refresh(method, block, defer_z=can_split)
        if pending is None:
            method.compute(block, dt)
            return
        method.compute_part(block, dt, abi.PART_INTERIOR, z_lo, z_hi)
        self.refresh_finish(method, block, pending)
        method.compute_part(block, dt, abi.PART_LOWER, z_lo, z_hi)
        method.compute_part(block, dt, abi.PART_UPPER, z_lo, z_hi)
        block.compute_done()

    def global_dt(self, dt, device=None):
        if hasattr(dt, "data_ptr"):
            # device-resident dt (Method.timestep_dev): reduce in place on the
            # stream, nothing comes back to the host
            if self.world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MIN)
            return dt
        if self.world == 1:
            return dt
        t = torch.tensor([dt], dtype=torch.float64,
                         device=device or ("cuda" if torch.cuda.is_available()
                                           else "cpu"))
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return float(t.item())
