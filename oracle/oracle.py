"""TEST INFRASTRUCTURE ONLY -- Python access to the two CPU checkers.

  kind="oracle": oracle/libvlct_oracle.so, this repo's plain-C restatement of
                 the reference's VL+CT algorithm (oracle/vlct_oracle.c)
  kind="ref":    oracle/_ref/libvlct_ref.so, the reference's own sources
                 compiled against oracle/ref_shim (built by `make -C oracle ref`
                 in the container that has /root/reference)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module. The product (enzo-e_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from enzo_e_b200 import abi  # noqa: E402  (struct layouts only)

ORACLE_LIB = os.path.join(_HERE, "libvlct_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libvlct_ref.so")
# integration/EnzoMethodMHDVlctGpu (the reference-side C++ binding of the CUDA
# library) behind the same Block/Field stand-ins -- a GPU path, listed here only
# because it is built by oracle/Makefile against the reference's headers
ADAPTER_LIB = os.path.join(_HERE, "_ref", "libvlct_adapter.so")


def build(target="all"):
    """Run the oracle Makefile (building the checker is not using it)."""
    subprocess.run(["make", "-C", _HERE, "-j8", target], check=True,
                   stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_LIB)


def have_oracle():
    return os.path.exists(ORACLE_LIB)


def have_adapter():
    return os.path.exists(ADAPTER_LIB)


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = {"oracle": ORACLE_LIB, "ref": REF_LIB, "adapter": ADAPTER_LIB}[kind]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} is missing: run `make -C oracle` (kind={kind})")
    lib = C.CDLL(path)
    pfx = {"oracle": "vlct_oracle", "ref": "vlct_ref",
           "adapter": "vlct_adapter"}[kind]
    create = getattr(lib, pfx + "_create")
    create.restype = C.c_void_p
    create.argtypes = [C.POINTER(abi.VlctConfig), C.c_int, C.c_int, C.c_int]
    destroy = getattr(lib, pfx + "_destroy")
    destroy.restype = None
    destroy.argtypes = [C.c_void_p]
    compute = getattr(lib, pfx + "_compute")
    compute.restype = C.c_int
    compute.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.c_double]
    timestep = getattr(lib, pfx + "_timestep")
    timestep.restype = C.c_int
    timestep.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock),
                         C.POINTER(C.c_double)]
    _libs[kind] = (lib, create, destroy, compute, timestep)
    return _libs[kind]


def numpy_block(fields, n, g, d, passive=()):
    """Build a vlct_block (host memory) over a dict of numpy arrays."""
    nx, ny, nz = n
    gx, gy, gz = g
    blk = abi.VlctBlock(nx=nx, ny=ny, nz=nz, gx=gx, gy=gy, gz=gz,
                        dx=d[0], dy=d[1], dz=d[2], mem_space=abi.MEM_HOST,
                        stream=None)
    for name in abi.CELL_FIELDS + abi.FACE_FIELDS + abi.OTHER_FIELDS:
        arr = fields.get(name)
        if arr is None:
            continue
        shape = abi.field_shape(name, nx, ny, nz, gx, gy, gz)
        assert arr.dtype == np.float64 and arr.flags.c_contiguous, name
        assert arr.shape == shape, (name, arr.shape, shape)
        setattr(blk, name, arr.ctypes.data_as(C.POINTER(C.c_double)))
    for i, name in enumerate(passive):
        arr = fields[name]
        assert arr.dtype == np.float64 and arr.flags.c_contiguous, name
        blk.passive[i] = arr.ctypes.data_as(C.POINTER(C.c_double))
    return blk


class CpuMethod:
    """A CPU `EnzoMethodMHDVlct` (oracle restatement or compiled reference)."""

    def __init__(self, cfg, ghost=(3, 3, 3), kind="oracle", store_fluxes=False,
                 gpu_batch_blocks=None, gpu_fused_timestep=None):
        """store_fluxes: construct the reference Method ("ref" / "adapter")
        with store_fluxes_for_corrections = true (the oracle restatement
        always keeps its flux arrays). gpu_batch_blocks / gpu_fused_timestep:
        the two parameter keys of the GPU binding ("adapter" only)."""
        self.kind = kind
        self.cfg = cfg
        (self._lib, create, self._destroy, self._compute,
         self._timestep) = _load(kind)
        if gpu_batch_blocks is not None or gpu_fused_timestep is not None:
            assert kind == "adapter" and not store_fluxes
            fn = self._lib.vlct_adapter_create_opts
            fn.restype = C.c_void_p
            fn.argtypes = [C.POINTER(abi.VlctConfig)] + [C.c_int] * 5

            def create(cfgp, gx, gy, gz):
                as_int = lambda v: -1 if v is None else int(bool(v))  # noqa: E731
                return fn(cfgp, gx, gy, gz, as_int(gpu_batch_blocks),
                          as_int(gpu_fused_timestep))
        if store_fluxes and kind != "oracle":
            pfx = {"ref": "vlct_ref", "adapter": "vlct_adapter"}[kind]
            create = getattr(self._lib, pfx + "_create_fc")
            create.restype = C.c_void_p
            create.argtypes = [C.POINTER(abi.VlctConfig), C.c_int, C.c_int, C.c_int]
        self._h = create(C.byref(cfg), *ghost)
        if not self._h:
            raise RuntimeError(f"{kind}: create failed")

    def close(self):
        if self._h:
            self._destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, blk, dt):
        rc = self._compute(self._h, C.byref(blk), float(dt))
        if rc != 0:
            raise RuntimeError(f"{self.kind}: compute failed ({rc})")

    def face_fluxes(self, blk, dt, n, n_fields):
        """save_fluxes_for_corrections_ of the last compute:
        {(dim, side, field): 2-D numpy array}. For "ref" / "adapter" (created
        with store_fluxes=True) this is what the Method itself deposited in
        the block's FluxData."""
        import numpy as np
        nslots = 6 + abi.VLCT_MAX_PASSIVE
        dp = C.POINTER(C.c_double)
        table = (((dp * nslots) * 2) * 3)()
        out = {}
        for dim in range(3):
            shape = abi.face_flux_shape(dim, *n)
            for side in range(2):
                for f in range(n_fields):
                    if f == 5 and not self.cfg.dual_energy:
                        continue
                    a = np.zeros(shape)
                    out[(dim, side, f)] = a
                    table[dim][side][f] = a.ctypes.data_as(dp)
        if self.kind == "oracle":
            fn = self._lib.vlct_oracle_face_fluxes
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.c_double,
                           C.c_void_p]
            rc = fn(self._h, C.byref(blk), float(dt), C.byref(table))
        else:
            pfx = {"ref": "vlct_ref", "adapter": "vlct_adapter"}[self.kind]
            fn = getattr(self._lib, pfx + "_face_fluxes")
            fn.restype = C.c_int
            fn.argtypes = [C.c_void_p, C.c_void_p]
            rc = fn(self._h, C.byref(table))
        if rc != 0:
            raise RuntimeError(f"vlct_oracle_face_fluxes failed ({rc})")
        return out

    # -- the GPU binding driven like a process with several blocks ---------
    def _block_array(self, blks):
        arr = (abi.VlctBlock * len(blks))()
        for i, b in enumerate(blks):
            arr[i] = b
        return arr

    def compute_many(self, blks, dt, cycle=0):
        """Method::compute on one block after the other, as Cello's compute
        phase does on a process owning all of them. Returns how many blocks
        had not yet reported compute_done() when the last call began."""
        assert self.kind == "adapter"
        fn = self._lib.vlct_adapter_compute_many
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.c_int, C.c_double,
                       C.c_int, C.POINTER(C.c_int)]
        deferred = C.c_int(0)
        rc = fn(self._h, self._block_array(blks), len(blks), float(dt), int(cycle),
                C.byref(deferred))
        if rc != 0:
            raise RuntimeError(f"adapter: compute_many failed ({rc})")
        return deferred.value

    def timestep_many(self, blks, cycle=0):
        assert self.kind == "adapter"
        fn = self._lib.vlct_adapter_timestep_many
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.c_int, C.c_int,
                       C.POINTER(C.c_double)]
        out = (C.c_double * len(blks))()
        rc = fn(self._h, self._block_array(blks), len(blks), int(cycle), out)
        if rc != 0:
            raise RuntimeError(f"adapter: timestep_many failed ({rc})")
        return list(out)

    def pup_roundtrip(self):
        """pack the Method with a PUP::er, rebuild it through the migration
        constructor + unpack; returns the packed size in bytes"""
        assert self.kind == "adapter"
        fn = self._lib.vlct_adapter_pup_roundtrip
        fn.restype = C.c_longlong
        fn.argtypes = [C.c_void_p]
        size = fn(self._h)
        if size <= 0:
            raise RuntimeError(f"adapter: pup round trip failed ({size})")
        return size

    def timestep(self, blk):
        out = C.c_double(0.0)
        rc = self._timestep(self._h, C.byref(blk), C.byref(out))
        if rc != 0:
            raise RuntimeError(f"{self.kind}: timestep failed ({rc})")
        return out.value


# ---------------------------------------------------------------------------
# problem initialisers + periodic refresh (oracle/vlct_oracle_ic.c)
# ---------------------------------------------------------------------------
_DBL_MIN = 2.2250738585072014e-308


def _ic_lib():
    lib = _load("oracle")[0]
    if not getattr(lib, "_ic_ready", False):
        dp = C.POINTER(C.c_double)
        lib.vlct_ic_inclined_wave.restype = C.c_int
        lib.vlct_ic_inclined_wave.argtypes = [
            C.POINTER(abi.VlctBlock), dp, C.c_double, C.c_char_p, C.c_double,
            C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        lib.vlct_ic_shock_tube.restype = C.c_int
        lib.vlct_ic_shock_tube.argtypes = [
            C.POINTER(abi.VlctBlock), dp, C.c_double, C.c_char_p, C.c_int,
            C.c_double, C.c_double, C.c_int]
        lib.vlct_ic_center_bfield.restype = C.c_int
        lib.vlct_ic_center_bfield.argtypes = [C.POINTER(abi.VlctBlock)]
        lib.vlct_oracle_refresh_periodic.restype = C.c_int
        lib.vlct_oracle_refresh_periodic.argtypes = [
            C.POINTER(abi.VlctBlock), C.c_int, C.c_int]
        lib.vlct_oracle_boundary.restype = C.c_int
        lib.vlct_oracle_boundary.argtypes = [
            C.POINTER(abi.VlctBlock), C.c_int, C.c_int, C.c_int, C.c_int]
        lib.vlct_oracle_boundary_inflow.restype = C.c_int
        lib.vlct_oracle_boundary_inflow.argtypes = [
            C.POINTER(abi.VlctBlock), C.c_int, C.c_int, C.c_int,
            C.POINTER(abi.VlctInflowValues)]
        lib.vlct_ic_cloud.restype = C.c_int
        lib.vlct_ic_cloud.argtypes = [C.POINTER(abi.VlctBlock), dp, C.c_int, dp]
        lib._ic_ready = True
    return lib


def ic_inclined_wave(blk, lower, gamma, wave_type, alpha, beta, amplitude=1e-6,
                     lam=1.0, positive_vel=True, parallel_vel=None,
                     ref_method=None):
    """EnzoInitialInclinedWave on a host block (fills ghost zones too).
    ref_method: a CpuMethod(kind="ref") whose field list (and gamma) the block
    matches -- then the reference's own compiled initialiser fills the block."""
    lo = (C.c_double * 3)(*lower)
    if ref_method is not None:
        assert ref_method.kind == "ref" and parallel_vel is None
        fn = ref_method._lib.vlct_ref_ic_inclined_wave
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.POINTER(C.c_double),
                       C.c_char_p, C.c_double, C.c_double, C.c_double, C.c_double,
                       C.c_int]
        rc = fn(ref_method._h, C.byref(blk), lo, wave_type.encode(), alpha, beta,
                amplitude, lam, 1 if positive_vel else 0)
        if rc != 0:
            raise RuntimeError(f"vlct_ref_ic_inclined_wave failed ({rc})")
        return
    pv = _DBL_MIN if parallel_vel is None else float(parallel_vel)
    rc = _ic_lib().vlct_ic_inclined_wave(
        C.byref(blk), lo, gamma, wave_type.encode(), alpha, beta, amplitude,
        lam, 1 if positive_vel else 0, pv)
    if rc != 0:
        raise RuntimeError(f"vlct_ic_inclined_wave failed ({rc})")


def ic_shock_tube(blk, lower, gamma, setup, aligned_ax=0, axis_velocity=0.0,
                  trans_velocity=0.0, flipped=False, ref_method=None):
    """EnzoInitialShockTube on a host block; ref_method as in ic_inclined_wave."""
    lo = (C.c_double * 3)(*lower)
    if ref_method is not None:
        assert ref_method.kind == "ref"
        fn = ref_method._lib.vlct_ref_ic_shock_tube
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.POINTER(C.c_double),
                       C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_int]
        rc = fn(ref_method._h, C.byref(blk), lo, setup.encode(), aligned_ax,
                axis_velocity, trans_velocity, 1 if flipped else 0)
        if rc != 0:
            raise RuntimeError(f"vlct_ref_ic_shock_tube failed ({rc})")
        return
    rc = _ic_lib().vlct_ic_shock_tube(
        C.byref(blk), lo, gamma, setup.encode(), aligned_ax, axis_velocity,
        trans_velocity, 1 if flipped else 0)
    if rc != 0:
        raise RuntimeError(f"vlct_ic_shock_tube failed ({rc})")


def center_bfield(blk):
    _ic_lib().vlct_ic_center_bfield(C.byref(blk))


def refresh_periodic(blk, n_passive=0, axes=7):
    """Ghost refresh of a single periodic block (host memory)."""
    _ic_lib().vlct_oracle_refresh_periodic(C.byref(blk), n_passive, axes)


BOUNDARY_TYPES = {"outflow": 0, "reflecting": 1}


def boundary(blk, axis, side, kind, n_passive=0, ref_method=None):
    """EnzoBoundary::enforce on one face of the domain (host memory);
    kind: "outflow" | "reflecting"; side 0 = lower, 1 = upper. ref_method: a
    CpuMethod(kind="ref") whose field list the block matches -- then the
    reference's own compiled EnzoBoundary does it instead of the restatement."""
    if ref_method is not None:
        assert ref_method.kind == "ref"
        fn = ref_method._lib.vlct_ref_boundary
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock), C.c_int, C.c_int,
                       C.c_int]
        rc = fn(ref_method._h, C.byref(blk), axis, side, BOUNDARY_TYPES[kind])
        if rc != 0:
            raise RuntimeError(f"vlct_ref_boundary failed ({rc})")
        return
    rc = _ic_lib().vlct_oracle_boundary(C.byref(blk), n_passive, axis, side,
                                        BOUNDARY_TYPES[kind])
    if rc != 0:
        raise RuntimeError(f"vlct_oracle_boundary failed ({rc})")


def boundary_inflow(blk, axis, side, values, passive=(), n_passive=0):
    """BoundaryValue::enforce with constant values (host memory): values =
    {field name: constant} is the boundary's field list."""
    v = abi.inflow_values(values, passive)
    rc = _ic_lib().vlct_oracle_boundary_inflow(C.byref(blk), n_passive, axis,
                                               side, C.byref(v))
    if rc != 0:
        raise RuntimeError(f"vlct_oracle_boundary_inflow failed ({rc})")


def ic_cloud(blk, lower, subsample_n, cloud_radius, center, cloud_density,
             wind_density, wind_velocity, wind_total_energy,
             wind_internal_energy, ref_method=None, perturb=None):
    """EnzoInitialCloud on a host block, ghost zones too; perturb = None or
    (Nwaves, seed, amplitude, min_lambda, max_lambda), the Initial:cloud:
    perturb_* parameters of the optional density perturbation;
    magnetic fields (if any) must already be initialised. ref_method: a
    CpuMethod(kind="ref") whose field list the block matches -- then the
    reference's own compiled EnzoInitialCloud::enforce_block fills the block
    instead of the C restatement."""
    lo = (C.c_double * 3)(*lower)
    p = (C.c_double * 9)(cloud_radius, *center, cloud_density, wind_density,
                         wind_velocity, wind_total_energy, wind_internal_energy)
    nw, seed, amp, lmin, lmax = perturb if perturb is not None else (0, 0, 0., 0., 0.)
    tail = [C.c_int, C.c_uint, C.c_double, C.c_double, C.c_double]
    if ref_method is not None:
        assert ref_method.kind == "ref"
        fn = ref_method._lib.vlct_ref_ic_cloud_perturbed
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(abi.VlctBlock),
                       C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)] + tail
        rc = fn(ref_method._h, C.byref(blk), lo, subsample_n, p, int(nw), int(seed),
                float(amp), float(lmin), float(lmax))
    else:
        fn = _ic_lib().vlct_ic_cloud_perturbed
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(abi.VlctBlock), C.POINTER(C.c_double), C.c_int,
                       C.POINTER(C.c_double)] + tail
        rc = fn(C.byref(blk), lo, subsample_n, p, int(nw), int(seed), float(amp),
                float(lmin), float(lmax))
    if rc != 0:
        raise RuntimeError(f"vlct_ic_cloud failed ({rc})")
