"""Host-side mirror of the reference's Method plugin for VL+CT.

`EnzoMethodMHDVlct` keeps the reference's surface -- construction from the
parameter-file keys, `compute(block)`, `timestep(block)`, `name()`
(src/Enzo/hydro-mhd/EnzoMethodMHDVlct.hpp:86-127) -- and forwards to the C ABI
in csrc/libvlct_b200.so. `Block` stands in for the few things the Method reads
from a Cello Block: the field arrays (Field::view), the active size and ghost
depth, EnzoBlock::CellWidth and Block::dt().

Fields may live in host memory (numpy arrays; staged through the GPU inside
each call) or in device memory (torch CUDA tensors; the performance mode).
"""
import ctypes as C

import numpy as np

from . import abi
from . import lib as _libmod
from .lib import VlctError

# parameter-file keys understood by vlct_config_set
METHOD_KEYS = ("riemann_solver", "reconstruct_method", "theta_limiter",
               "mhd_choice", "time_scheme", "courant")


def config_from_parameters(params, n_passive=0, has_acceleration=False):
    """Build a vlct_config from {parameter-file key: value}.

    Keys may be given in full ("Method:mhd_vlct:riemann_solver",
    "Physics:fluid_props:eos:gamma") or, for the Method group, by their last
    component ("riemann_solver")."""
    lib = _libmod.load()
    cfg = abi.VlctConfig()
    lib.vlct_config_init(C.byref(cfg))
    err = C.create_string_buffer(512)
    for key, value in params.items():
        full = key if ":" in key else "Method:mhd_vlct:" + key
        if isinstance(value, bool):
            value = "true" if value else "false"
        elif isinstance(value, float):
            value = repr(value)
        rc = lib.vlct_config_set(C.byref(cfg), full.encode(),
                                 str(value).encode(), err, len(err))
        if rc != abi.VLCT_OK:
            raise VlctError(rc, err.value.decode())
    cfg.n_passive = n_passive
    cfg.has_acceleration = 1 if has_acceleration else 0
    return cfg


class Block:
    """The slice of a Cello Block that the Method touches."""

    def __init__(self, fields, n, g, cell_width, passive=(), dt=0.0,
                 stream=None):
        self.fields = fields          # name -> numpy array | torch.Tensor
        self.n = tuple(n)
        self.g = tuple(g)
        self.cell_width = tuple(cell_width)
        self.passive = tuple(passive)
        self.dt = dt
        self.compute_done_count = 0
        first = next(iter(fields.values()))
        self.on_device = not isinstance(first, np.ndarray)
        self.stream_is_current = False
        if self.on_device and stream is None:
            # run on torch's current stream, so that the library's kernels are
            # ordered after whatever produced the tensors (the C ABI's own
            # default -- stream NULL -- is a private non-blocking stream, which
            # would NOT wait for torch's work). 0x1 is cudaStreamLegacy.
            import torch
            stream = torch.cuda.current_stream(first.device).cuda_stream or 0x1
            self.stream_is_current = True
        self._c = self._build(stream)

    def compute_done(self):
        """Method::compute must end by calling block->compute_done()
        (src/Cello/problem_Method.hpp:47-55)."""
        self.compute_done_count += 1

    def _ptr(self, name, arr):
        shape = abi.field_shape(name, *self.n, *self.g)
        if tuple(arr.shape) != shape:
            raise ValueError(f"{name}: shape {tuple(arr.shape)} != {shape}")
        if self.on_device:
            import torch
            if arr.dtype != torch.float64 or not arr.is_contiguous() \
                    or not arr.is_cuda:
                raise ValueError(f"{name}: need a contiguous fp64 CUDA tensor")
            return C.cast(arr.data_ptr(), C.POINTER(C.c_double))
        if arr.dtype != np.float64 or not arr.flags.c_contiguous:
            raise ValueError(f"{name}: need a C-contiguous float64 array")
        return arr.ctypes.data_as(C.POINTER(C.c_double))

    def _build(self, stream):
        nx, ny, nz = self.n
        gx, gy, gz = self.g
        blk = abi.VlctBlock(
            nx=nx, ny=ny, nz=nz, gx=gx, gy=gy, gz=gz,
            dx=self.cell_width[0], dy=self.cell_width[1],
            dz=self.cell_width[2],
            mem_space=abi.MEM_DEVICE if self.on_device else abi.MEM_HOST,
            stream=stream)
        for name in abi.CELL_FIELDS + abi.FACE_FIELDS + abi.OTHER_FIELDS:
            arr = self.fields.get(name)
            if arr is not None:
                setattr(blk, name, self._ptr(name, arr))
        for i, name in enumerate(self.passive):
            blk.passive[i] = self._ptr("density", self.fields[name])
        return blk

    @property
    def c_block(self):
        return self._c


class EnzoMethodMHDVlct:
    """`Method` plugin "mhd_vlct", computed on the GPU.

    Mirrors src/Enzo/hydro-mhd/EnzoMethodMHDVlct.{hpp,cpp}: the constructor
    validates the parameters exactly as the reference's constructors do and
    raises `VlctError` where the reference would ASSERT/ERROR."""

    def __init__(self, params=None, config=None, n_passive=0,
                 has_acceleration=False):
        self._lib = _libmod.load()
        if config is None:
            config = config_from_parameters(params or {}, n_passive,
                                            has_acceleration)
        self.config = config
        self._h = C.c_void_p()
        rc = self._lib.vlct_create(C.byref(config), C.byref(self._h))
        if rc != abi.VLCT_OK:
            msg = self._lib.vlct_last_error(self._h).decode() if self._h \
                else self._lib.vlct_status_string(rc).decode()
            if self._h:
                self._lib.vlct_destroy(self._h)
                self._h = C.c_void_p()
            raise VlctError(rc, msg)

    # -- Method interface ---------------------------------------------------
    def name(self):
        return self._lib.vlct_name().decode()

    def compute(self, block, dt=None):
        """EnzoMethodMHDVlct::compute(Block*): advance by block.dt in place.

        dt may be a one-element fp64 CUDA tensor (see timestep_dev): the step
        then reads it on the device and nothing waits for the host."""
        dt = block.dt if dt is None else dt
        if hasattr(dt, "data_ptr"):
            self._check(self._lib.vlct_compute_dev(
                self._h, C.byref(block.c_block),
                C.cast(dt.data_ptr(), C.POINTER(C.c_double))))
        else:
            self._check(self._lib.vlct_compute(self._h, C.byref(block.c_block),
                                               float(dt)))
        block.compute_done()

    def compute_and_timestep(self, block, dt=None):
        """compute(block) followed by the timestep(block) of the next cycle in
        one call (vlct_compute_and_timestep): one upload and one download per
        cycle for HOST blocks. Returns the next dt."""
        dt = block.dt if dt is None else dt
        out = C.c_double(0.0)
        self._check(self._lib.vlct_compute_and_timestep(
            self._h, C.byref(block.c_block), float(dt), C.byref(out)))
        block.compute_done()
        return out.value

    def compute_and_timestep_dev(self, block, dt, out=None):
        """compute(block) + the next cycle's timestep with both values on the
        device (vlct_compute_and_timestep_dev); dt, out: one-element fp64 CUDA
        tensors (out may be dt itself). Asynchronous."""
        import torch
        if out is None:
            out = torch.empty(1, dtype=torch.float64, device=dt.device)
        self._check(self._lib.vlct_compute_and_timestep_dev(
            self._h, C.byref(block.c_block),
            C.cast(dt.data_ptr(), C.POINTER(C.c_double)),
            C.cast(out.data_ptr(), C.POINTER(C.c_double))))
        block.compute_done()
        return out

    def compute_part(self, block, dt, part, z_lo, z_hi, dt_next=None):
        """One of the three parts of a step (vlct_compute_dev_part): the
        interior first, then the lower / upper rest once the z ghost levels
        have arrived. dt: one-element fp64 CUDA tensor. The caller calls
        block.compute_done() after the last part."""
        if dt_next is not None:     # with the next cycle's timestep folded in
            self._check(self._lib.vlct_compute_and_timestep_dev_part(
                self._h, C.byref(block.c_block),
                C.cast(dt.data_ptr(), C.POINTER(C.c_double)), part, z_lo, z_hi,
                C.cast(dt_next.data_ptr(), C.POINTER(C.c_double))))
            return
        self._check(self._lib.vlct_compute_dev_part(
            self._h, C.byref(block.c_block),
            C.cast(dt.data_ptr(), C.POINTER(C.c_double)), part, z_lo, z_hi))

    # -- flux corrections -------------------------------------------------------
    def save_face_fluxes(self, block, n_fields=None, device=None):
        """dt/dx * final-stage fluxes through the block's six faces, as the
        reference stores them for Method "flux_correct"
        (EnzoMethodMHDVlct::save_fluxes_for_corrections_). Returns
        {(dim, side, field): 2-D array}; numpy arrays, or CUDA tensors when
        `device` is given. field: 0 density, 1..3 momentum x..z, 4 total
        energy, 5 internal energy (dual energy), 6+s passive scalar s."""
        nf = 6 + self.config.n_passive if n_fields is None else n_fields
        ff = abi.VlctFaceFluxes()
        ff.mem_space = abi.MEM_HOST if device is None else abi.MEM_DEVICE
        out = {}
        for dim in range(3):
            shape = abi.face_flux_shape(dim, *block.n)
            for side in range(2):
                for f in range(nf):
                    if f == 5 and not self.config.dual_energy:
                        continue
                    if device is None:
                        a = np.zeros(shape)
                        ptr = a.ctypes.data_as(C.POINTER(C.c_double))
                    else:
                        import torch
                        a = torch.zeros(shape, dtype=torch.float64, device=device)
                        ptr = C.cast(a.data_ptr(), C.POINTER(C.c_double))
                    out[(dim, side, f)] = a
                    ff.face[dim][side][f] = ptr
        self._check(self._lib.vlct_save_face_fluxes(
            self._h, C.byref(block.c_block), C.byref(ff)))
        return out

    # -- many blocks per launch ------------------------------------------------
    @staticmethod
    def _block_array(blocks):
        arr = (abi.VlctBlock * len(blocks))()
        for i, b in enumerate(blocks):
            arr[i] = b.c_block
        return arr

    def compute_batch(self, blocks, dt):
        """compute() for a list of equally shaped blocks in one set of kernel
        launches (vlct_compute_batch)."""
        arr = self._block_array(blocks)
        self._check(self._lib.vlct_compute_batch(self._h, arr, len(blocks),
                                                 float(dt)))
        for b in blocks:
            b.compute_done()

    def compute_and_timestep_batch(self, blocks, dt):
        """compute_batch followed by timestep_batch in one call
        (vlct_compute_and_timestep_batch); returns the next dt."""
        arr = self._block_array(blocks)
        out = C.c_double(0.0)
        self._check(self._lib.vlct_compute_and_timestep_batch(
            self._h, arr, len(blocks), float(dt), C.byref(out)))
        for b in blocks:
            b.compute_done()
        return out.value

    def timestep_batch(self, blocks):
        """min over the blocks of timestep(block); fills every "pressure"."""
        arr = self._block_array(blocks)
        out = C.c_double(0.0)
        self._check(self._lib.vlct_timestep_batch(self._h, arr, len(blocks),
                                                  C.byref(out)))
        return out.value

    def host_register(self, array):
        """page-lock a numpy array for fast staging (vlct_host_register)"""
        self._check(self._lib.vlct_host_register(
            self._h, array.ctypes.data_as(C.c_void_p), array.nbytes))

    def host_unregister(self, array):
        self._check(self._lib.vlct_host_unregister(
            self._h, array.ctypes.data_as(C.c_void_p)))

    def set_option(self, key, value):
        self._check(self._lib.vlct_set_option(self._h, key.encode(), int(value)))

    def timestep_dev(self, block, out=None):
        """timestep() without the host round trip: returns a one-element fp64
        CUDA tensor holding courant * min(...), filled asynchronously."""
        import torch
        if out is None:
            dev = next(iter(block.fields.values())).device
            out = torch.empty(1, dtype=torch.float64, device=dev)
        self._check(self._lib.vlct_timestep_dev(
            self._h, C.byref(block.c_block),
            C.cast(out.data_ptr(), C.POINTER(C.c_double))))
        return out

    def timestep(self, block):
        """EnzoMethodMHDVlct::timestep(Block*) (already times courant)."""
        out = C.c_double(0.0)
        self._check(self._lib.vlct_timestep(self._h, C.byref(block.c_block),
                                            C.byref(out)))
        return out.value

    # -- refresh stand-ins ----------------------------------------------------
    def refresh_periodic(self, block, axes=7):
        self._check(self._lib.vlct_refresh_periodic(
            self._h, C.byref(block.c_block), axes))

    def boundary(self, block, axis, side, kind):
        """EnzoBoundary::enforce on one face of the domain: kind "outflow" or
        "reflecting", side 0 = lower / 1 = upper (apply after the periodic /
        neighbour refresh of the other axes, like Block::update_boundary_)."""
        self._check(self._lib.vlct_boundary(
            self._h, C.byref(block.c_block), axis, side, abi.BOUNDARY[kind]))

    def boundary_inflow(self, block, axis, side, values, passive=()):
        """BoundaryValue::enforce ("inflow") with constant values on one face of
        the domain: values = {field name: constant} is the boundary's field
        list (Cello/problem_BoundaryValue.cpp:131-273); passive = one constant
        (or None) per "color" field."""
        v = abi.inflow_values(values, passive)
        self._check(self._lib.vlct_boundary_inflow(
            self._h, C.byref(block.c_block), axis, side, C.byref(v)))

    def halo_bytes(self, block, axis):
        return self._lib.vlct_halo_bytes(self._h, C.byref(block.c_block), axis)

    def halo_pack(self, block, axis, side, buffer):
        self._check(self._lib.vlct_halo_pack(
            self._h, C.byref(block.c_block), axis, side,
            C.cast(buffer.data_ptr(), C.POINTER(C.c_double))))

    def halo_unpack(self, block, axis, side, buffer):
        self._check(self._lib.vlct_halo_unpack(
            self._h, C.byref(block.c_block), axis, side,
            C.cast(buffer.data_ptr(), C.POINTER(C.c_double))))

    # -- instrumentation --------------------------------------------------------
    def kernel_launches(self):
        return int(self._lib.vlct_kernel_launches(self._h))

    def scratch_bytes(self):
        return int(self._lib.vlct_scratch_bytes(self._h))

    def staged_bytes(self):
        """(host->device, device->host) bytes copied so far for HOST blocks"""
        return (int(self._lib.vlct_staged_bytes(self._h, 0)),
                int(self._lib.vlct_staged_bytes(self._h, 1)))

    def profile(self, on=True):
        """Switch per-kernel CUDA-event timing on/off (clears old samples)."""
        self._check(self._lib.vlct_profile_reset(self._h))
        self._check(self._lib.vlct_profile_enable(self._h, 1 if on else 0))

    def profile_report(self):
        """{kernel name: (total ms, launches)} since profiling was enabled."""
        out = {}
        name = C.create_string_buffer(64)
        ms, calls = C.c_double(0.0), C.c_longlong(0)
        for i in range(self._lib.vlct_profile_count(self._h)):
            self._check(self._lib.vlct_profile_get(
                self._h, i, name, len(name), C.byref(ms), C.byref(calls)))
            out[name.value.decode()] = (ms.value, calls.value)
        return out

    def synchronize(self):
        self._check(self._lib.vlct_synchronize(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.vlct_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != abi.VLCT_OK:
            raise VlctError(rc, self._lib.vlct_last_error(self._h).decode())
