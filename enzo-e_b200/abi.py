"""ctypes mirror of include/vlct.h (the C ABI of the VL+CT block update).

Only plain-C layouts live here: `VlctConfig` <-> `vlct_config`,
`VlctBlock` <-> `vlct_block`, and the enum values. Nothing in this module
touches a GPU or loads a library.
"""
import ctypes as C

VLCT_MAX_PASSIVE = 16

# status codes
VLCT_OK = 0
VLCT_ERR_INVALID_CONFIG = 1
VLCT_ERR_INVALID_BLOCK = 2
VLCT_ERR_CUDA = 3
VLCT_ERR_NO_DEVICE = 4
VLCT_ERR_UNKNOWN_KEY = 5
VLCT_ERR_INTERNAL = 6

RIEMANN = {"hll": 0, "hlle": 1, "hllc": 2, "hlld": 3}
RECON = {"nn": 0, "plm": 1, "plm_enzo": 1, "plm_athena": 2}
MHD_CHOICE = {"unset": -1, "no_bfield": 0, "constrained_transport": 1}
TIME_SCHEME = {"vl": 0, "euler": 1}
DUAL_ENERGY = {"disabled": 0, "modern": 1, "bryan95": 2}
MEM_HOST, MEM_DEVICE = 0, 1
BOUNDARY = {"outflow": 0, "reflecting": 1}
PART_INTERIOR, PART_LOWER, PART_UPPER = 0, 1, 2
# input levels an interior part reads below z_lo / beyond z_hi (vlct.h)
PART_REACH_BELOW, PART_REACH_ABOVE = 5, 6

_DP = C.POINTER(C.c_double)


class VlctConfig(C.Structure):
    _fields_ = [
        ("riemann_solver", C.c_int),
        ("reconstruct_method", C.c_int),
        ("theta_limiter", C.c_double),
        ("mhd_choice", C.c_int),
        ("time_scheme", C.c_int),
        ("courant", C.c_double),
        ("gamma", C.c_double),
        ("dual_energy", C.c_int),
        ("dual_energy_eta", C.c_double),
        ("density_floor", C.c_double),
        ("pressure_floor", C.c_double),
        ("n_passive", C.c_int),
        ("has_acceleration", C.c_int),
    ]


CELL_FIELDS = ("density", "velocity_x", "velocity_y", "velocity_z",
               "total_energy", "internal_energy",
               "bfield_x", "bfield_y", "bfield_z")
FACE_FIELDS = ("bfieldi_x", "bfieldi_y", "bfieldi_z")
OTHER_FIELDS = ("pressure", "acceleration_x", "acceleration_y",
                "acceleration_z")


class VlctBlock(C.Structure):
    _fields_ = (
        [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
         ("gx", C.c_int), ("gy", C.c_int), ("gz", C.c_int),
         ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double)]
        + [(name, _DP) for name in CELL_FIELDS]
        + [(name, _DP) for name in FACE_FIELDS]
        + [(name, _DP) for name in OTHER_FIELDS]
        + [("passive", _DP * VLCT_MAX_PASSIVE),
           ("mem_space", C.c_int),
           ("stream", C.c_void_p)]
    )


INFLOW_FIELDS = ("density", "velocity_x", "velocity_y", "velocity_z",
                 "total_energy", "internal_energy",
                 "bfield_x", "bfield_y", "bfield_z",
                 "bfieldi_x", "bfieldi_y", "bfieldi_z", "pressure")


class VlctInflowValues(C.Structure):
    """vlct_inflow_values: NaN = the field is not in the boundary's list"""
    _fields_ = ([(name, C.c_double) for name in INFLOW_FIELDS]
                + [("passive", C.c_double * VLCT_MAX_PASSIVE)])


def inflow_values(values, passive=()):
    """{field name: constant} (+ one entry per passive scalar, None = not
    listed) -> VlctInflowValues"""
    nan = float("nan")
    unknown = set(values) - set(INFLOW_FIELDS)
    if unknown:
        raise KeyError(f"not fields of vlct_block: {sorted(unknown)}")
    v = VlctInflowValues(**{k: float(values.get(k, nan)) for k in INFLOW_FIELDS})
    for i in range(VLCT_MAX_PASSIVE):
        v.passive[i] = nan
    for i, x in enumerate(passive):
        if x is not None:
            v.passive[i] = float(x)
    return v


VLCT_FLUX_FIELDS = 6 + VLCT_MAX_PASSIVE


class VlctFaceFluxes(C.Structure):
    _fields_ = [("face", ((_DP * VLCT_FLUX_FIELDS) * 2) * 3),
                ("mem_space", C.c_int)]


def face_flux_shape(dim, nx, ny, nz):
    """(slower, faster) shape of one face array of vlct_face_fluxes"""
    return {0: (nz, ny), 1: (nz, nx), 2: (ny, nx)}[dim]


def default_config():
    """The defaults vlct_config_init() produces (reference defaults)."""
    return VlctConfig(
        riemann_solver=RIEMANN["hlld"], reconstruct_method=RECON["plm"],
        theta_limiter=1.5, mhd_choice=MHD_CHOICE["unset"],
        time_scheme=TIME_SCHEME["vl"], courant=-1.0, gamma=5.0 / 3.0,
        dual_energy=DUAL_ENERGY["disabled"], dual_energy_eta=0.001,
        density_floor=0.0, pressure_floor=0.0, n_passive=0,
        has_acceleration=0)


def field_shape(name, nx, ny, nz, gx, gy, gz):
    """C-order (z, y, x) shape of a field including ghost zones."""
    mx, my, mz = nx + 2 * gx, ny + 2 * gy, nz + 2 * gz
    if name == "bfieldi_x":
        return (mz, my, mx + 1)
    if name == "bfieldi_y":
        return (mz, my + 1, mx)
    if name == "bfieldi_z":
        return (mz + 1, my, mx)
    return (mz, my, mx)
