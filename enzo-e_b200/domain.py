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
  * a non-periodic domain (`boundaries`: Cello's Boundary:list, in order):
    bricks at a domain face have no neighbour across it; once all exchanges
    are done every boundary object is enforced on the faces it applies to,
    like Block::update_boundary_ (src/Cello/mesh_Block.cpp:1057-1077,
    control_refresh.cpp:229-232)
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
    def __init__(self, rank=0, world=1, grid=None, boundaries=None):
        """boundaries: None = periodic domain; else the Boundary objects of the
        problem in list order, each a dict
          {"type": "outflow" | "reflecting" | "inflow",
           "axis": 0 | 1 | 2 | None (all), "face": 0 | 1 | None (both),
           "values": {field: constant}, "passive": (...)}   # inflow only
        An axis that no entry applies to stays periodic."""
        self.rank, self.world = rank, world
        self.grid = tuple(grid) if grid else proc_grid(world)
        px, py, pz = self.grid
        assert px * py * pz == world
        # rank = (cz * py + cy) * px + cx
        self.coords = (rank % px, (rank // px) % py, rank // (px * py))
        self.boundaries = [dict(b) for b in (boundaries or [])]
        self.periodic = [True, True, True]
        for b in self.boundaries:
            if b["type"] not in ("outflow", "reflecting", "inflow"):
                raise ValueError(f"unknown boundary type {b['type']!r}")
            for axis in ([0, 1, 2] if b.get("axis") is None else [b["axis"]]):
                self.periodic[axis] = False

    def on_boundary(self, axis, side):
        """does this brick touch the (non-periodic) domain face?"""
        if self.periodic[axis]:
            return False
        return self.coords[axis] == (0 if side == 0 else self.grid[axis] - 1)

    def apply_boundaries(self, method, block, boundary=None, boundary_inflow=None):
        """Block::update_boundary_: every Boundary object in list order, each
        over the faces (x lower, x upper, y lower, ...) it applies to and this
        brick has on the domain boundary."""
        boundary = boundary or method.boundary
        boundary_inflow = boundary_inflow or method.boundary_inflow
        for b in self.boundaries:
            for axis in range(3):
                if b.get("axis") is not None and b["axis"] != axis:
                    continue
                for side in (0, 1):
                    if b.get("face") is not None and b["face"] != side:
                        continue
                    if not self.on_boundary(axis, side):
                        continue
                    if b["type"] == "inflow":
                        boundary_inflow(block, axis, side, b["values"],
                                        b.get("passive", ()))
                    else:
                        boundary(block, axis, side, b["type"])

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
                wrap=None, alloc=None, defer_z=False, boundary=None,
                boundary_inflow=None):
        """Fill all ghost zones of `block`.

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
        bc = (boundary, boundary_inflow)
        for axis in range(3):
            if self.grid[axis] == 1:
                if self.periodic[axis]:
                    wrap(block, 1 << axis)
                continue
            key = axis
            if key not in buffers:
                nbytes = method.halo_bytes(block, axis)
                buffers[key] = [alloc(nbytes) for _ in range(4)]
            send_lo, send_hi, recv_lo, recv_hi = buffers[key]
            # no neighbour across a non-periodic domain face
            has_lo = not self.on_boundary(axis, 0)
            has_hi = not self.on_boundary(axis, 1)
            if has_lo:
                pack(block, axis, 0, send_lo)
            if has_hi:
                pack(block, axis, 1, send_hi)
            lo, hi = self.neighbor(axis, -1), self.neighbor(axis, +1)
            # If the block's kernels run on torch's current stream (bench.py
            # does that), packs, NCCL and unpacks are stream-ordered and no
            # host synchronisation is needed; otherwise fence explicitly.
            same_stream = getattr(block, "stream_is_current", False)
            if send_lo.is_cuda and not same_stream:
                method.synchronize()
            ops = []
            if has_lo:
                ops.append(dist.P2POp(dist.isend, send_lo, lo))
            if has_hi:
                ops.append(dist.P2POp(dist.isend, send_hi, hi))
                ops.append(dist.P2POp(dist.irecv, recv_hi, hi))
            if has_lo:
                ops.append(dist.P2POp(dist.irecv, recv_lo, lo))
            reqs = dist.batch_isend_irecv(ops) if ops else []
            pending = (reqs, axis, recv_lo if has_lo else None,
                       recv_hi if has_hi else None, unpack, same_stream, bc)
            if defer_z and axis == 2:
                # NCCL works on its own stream, ordered after the packs; the
                # caller's stream only joins it in refresh_finish(). What the
                # interior part of the update reads of the x / y domain
                # boundaries does not depend on the z ghosts: enforce them now
                # (refresh_finish enforces everything again, in order, once the
                # z ghosts are there)
                if self.boundaries:
                    self.apply_boundaries(method, block, *bc)
                return pending
            self._finish_axis(method, block, pending)
        if self.boundaries:
            self.apply_boundaries(method, block, *bc)
        return None

    def _finish_axis(self, method, block, pending):
        reqs, axis, recv_lo, recv_hi, unpack, same_stream, _ = pending
        for req in reqs:
            req.wait()
        if not same_stream and any(
                t is not None and t.is_cuda for t in (recv_lo, recv_hi)):
            torch.cuda.current_stream().synchronize()
        if recv_lo is not None:
            unpack(block, axis, 0, recv_lo)
        if recv_hi is not None:
            unpack(block, axis, 1, recv_hi)

    def refresh_finish(self, method, block, pending):
        """Wait for a deferred exchange, unpack its ghost slabs and enforce the
        domain boundaries."""
        self._finish_axis(method, block, pending)
        if self.boundaries:
            self.apply_boundaries(method, block, *pending[6])

    def step(self, method, block, dt, overlap=True, dt_next=None):
        """refresh + compute of one cycle (dt: device-resident, already
        reduced over ranks). With overlap the z exchange runs under the
        interior part of the update. dt_next (one-element fp64 CUDA tensor):
        also evaluate this block's timestep of the NEXT cycle, folded into the
        update (vlct_compute_and_timestep_dev[_part]); the caller min-reduces
        it over the ranks (global_dt)."""
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
            if dt_next is not None:
                method.compute_and_timestep_dev(block, dt, out=dt_next)
            else:
                method.compute(block, dt)
            return
        method.compute_part(block, dt, abi.PART_INTERIOR, z_lo, z_hi, dt_next)
        self.refresh_finish(method, block, pending)
        method.compute_part(block, dt, abi.PART_LOWER, z_lo, z_hi, dt_next)
        method.compute_part(block, dt, abi.PART_UPPER, z_lo, z_hi, dt_next)
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
