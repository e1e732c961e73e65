"""Loader for csrc/libvlct_b200.so -- the CUDA kernels behind the C ABI.

The product path has no fallback: if the shared library has not been built
(`python -c "import __graft_entry__ as g; g.build()"` or `make -C
enzo-e_b200/csrc`) importing this module's `load()` raises.
"""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VLCT_B200_LIB") or os.path.join(
    _HERE, "csrc", "libvlct_b200.so")   # env override: A/B builds while tuning

# every symbol include/vlct.h declares
EXPORTED_SYMBOLS = (
    "vlct_config_init", "vlct_config_set", "vlct_config_validate",
    "vlct_create", "vlct_destroy", "vlct_name", "vlct_compute",
    "vlct_timestep", "vlct_timestep_dev", "vlct_compute_dev",
    "vlct_compute_and_timestep", "vlct_compute_and_timestep_batch",
    "vlct_compute_and_timestep_dev", "vlct_compute_and_timestep_dev_part",
    "vlct_compute_dev_part", "vlct_set_option",
    "vlct_compute_batch", "vlct_timestep_batch", "vlct_save_face_fluxes",
    "vlct_host_register", "vlct_host_unregister",
    "vlct_last_error", "vlct_status_string",
    "vlct_kernel_launches", "vlct_scratch_bytes", "vlct_staged_bytes",
    "vlct_synchronize",
    "vlct_profile_enable", "vlct_profile_reset", "vlct_profile_count",
    "vlct_profile_get", "vlct_selftest_fpops",
    "vlct_refresh_periodic", "vlct_boundary", "vlct_boundary_inflow",
    "vlct_halo_bytes", "vlct_halo_pack",
    "vlct_halo_unpack",
)

_lib = None


class VlctError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"vlct status {status}: {message}")
        self.status = status
        self.message = message


def load():
    """Load the shared library (once) and declare the C signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build the CUDA extension first "
            "(__graft_entry__.build() or `make -C enzo-e_b200/csrc`). "
            "There is no CPU fallback for the VL+CT path.")
    lib = C.CDLL(LIB_PATH)
    cfgp, blkp = C.POINTER(abi.VlctConfig), C.POINTER(abi.VlctBlock)
    dp = C.POINTER(C.c_double)
    sig = {
        "vlct_config_init": (C.c_int, [cfgp]),
        "vlct_config_set": (C.c_int, [cfgp, C.c_char_p, C.c_char_p,
                                      C.c_char_p, C.c_int]),
        "vlct_config_validate": (C.c_int, [cfgp, C.c_char_p, C.c_int]),
        "vlct_create": (C.c_int, [cfgp, C.POINTER(C.c_void_p)]),
        "vlct_destroy": (None, [C.c_void_p]),
        "vlct_name": (C.c_char_p, []),
        "vlct_compute": (C.c_int, [C.c_void_p, blkp, C.c_double]),
        "vlct_timestep": (C.c_int, [C.c_void_p, blkp, dp]),
        "vlct_compute_and_timestep": (C.c_int, [C.c_void_p, blkp, C.c_double, dp]),
        "vlct_timestep_dev": (C.c_int, [C.c_void_p, blkp, dp]),
        "vlct_compute_dev": (C.c_int, [C.c_void_p, blkp, dp]),
        "vlct_compute_dev_part": (C.c_int, [C.c_void_p, blkp, dp, C.c_int,
                                            C.c_int, C.c_int]),
        "vlct_compute_and_timestep_dev": (C.c_int, [C.c_void_p, blkp, dp, dp]),
        "vlct_compute_and_timestep_dev_part": (C.c_int, [C.c_void_p, blkp, dp, C.c_int,
                                                         C.c_int, C.c_int, dp]),
        "vlct_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_longlong]),
        "vlct_compute_batch": (C.c_int, [C.c_void_p, blkp, C.c_int, C.c_double]),
        "vlct_timestep_batch": (C.c_int, [C.c_void_p, blkp, C.c_int, dp]),
        "vlct_compute_and_timestep_batch": (C.c_int, [C.c_void_p, blkp, C.c_int,
                                                      C.c_double, dp]),
        "vlct_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_ulonglong]),
        "vlct_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
        "vlct_save_face_fluxes": (C.c_int, [C.c_void_p, blkp,
                                            C.POINTER(abi.VlctFaceFluxes)]),
        "vlct_last_error": (C.c_char_p, [C.c_void_p]),
        "vlct_status_string": (C.c_char_p, [C.c_int]),
        "vlct_kernel_launches": (C.c_longlong, [C.c_void_p]),
        "vlct_scratch_bytes": (C.c_longlong, [C.c_void_p]),
        "vlct_staged_bytes": (C.c_longlong, [C.c_void_p, C.c_int]),
        "vlct_synchronize": (C.c_int, [C.c_void_p]),
        "vlct_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
        "vlct_profile_reset": (C.c_int, [C.c_void_p]),
        "vlct_profile_count": (C.c_int, [C.c_void_p]),
        "vlct_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p,
                                       C.c_int, dp, C.POINTER(C.c_longlong)]),
        "vlct_selftest_fpops": (C.c_int, [C.c_longlong, C.c_ulonglong, C.c_int,
                                          C.POINTER(C.c_longlong)]),
        "vlct_refresh_periodic": (C.c_int, [C.c_void_p, blkp, C.c_int]),
        "vlct_boundary": (C.c_int, [C.c_void_p, blkp, C.c_int, C.c_int, C.c_int]),
        "vlct_boundary_inflow": (C.c_int, [C.c_void_p, blkp, C.c_int, C.c_int,
                                           C.POINTER(abi.VlctInflowValues)]),
        "vlct_halo_bytes": (C.c_longlong, [C.c_void_p, blkp, C.c_int]),
        "vlct_halo_pack": (C.c_int, [C.c_void_p, blkp, C.c_int, C.c_int, dp]),
        "vlct_halo_unpack": (C.c_int, [C.c_void_p, blkp, C.c_int, C.c_int, dp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
