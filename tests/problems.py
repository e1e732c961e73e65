"""The reference's vlct answer-test problems restated on raw arrays.

Mirrors input/vlct/*.in + input/vlct/run_*_test.py + tools/l1_error_norm.py of
the reference: same parameter values, same cycle order (stopping/timestep ->
refresh -> compute, src/Cello/control_stopping.cpp:44-142,
control_compute.cpp:42-157), same L1 error norm.
"""
import numpy as np

from helpers import abi, make_config, oracle

# input/vlct/MHD_linear_wave/initial_*.in, HD_linear_wave/initial_*.in
ALPHA = 0.7297276562269663   # sin(alpha) = 2/3
BETA = 1.1071487177940904    # sin(beta)  = 2/sqrt(5)

# name -> (wave_type parameter, final time)
MHD_WAVES = {"fast": ("fast", 0.5), "alfven": ("alfven", 1.0),
             "slow": ("slow", 2.0), "entropy": ("mhd_entropy", 1.0)}
HD_WAVES = {"sound": ("sound", 1.0), "hd_entropy": ("hd_entropy", 1.0),
            "hd_transv_entropy_v1": ("hd_transv_entropy_v1", 1.0),
            "hd_transv_entropy_v2": ("hd_transv_entropy_v2", 1.0)}

# golden L1 norms: input/vlct/run_MHD_linear_wave_test.py:178-188,
#                  input/vlct/run_HD_linear_wave_test.py:70-80
GOLDEN_MHD = {("fast", 16): 1.6388526155394664e-07,
              ("fast", 32): 3.302538226654406e-08,
              ("alfven", 16): 1.927245356389947e-07,
              ("alfven", 32): 3.005870212811956e-08,
              ("slow", 16): 2.2373810027584788e-07,
              ("slow", 32): 4.43702e-08,
              ("entropy", 16): 1.0021263485338544e-07,
              ("entropy", 32): 2.9194839706868883e-08}
GOLDEN_HD = {("sound", 16): 1.3704437791196907e-07,
             ("sound", 32): 2.3484946798755385e-08,
             ("hd_entropy", 16): 8.736217559091042e-08,
             ("hd_entropy", 32): 2.2383944940640457e-08,
             ("hd_transv_entropy_v1", 16): 7.671004076321944e-08,
             ("hd_transv_entropy_v1", 32): 1.6725791096617187e-08,
             ("hd_transv_entropy_v2", 16): 8.640730176761958e-08,
             ("hd_transv_entropy_v2", 32): 1.6960029746447077e-08}


def golden_isclose(a, b):
    """input/vlct/testing_utils.py:275-292 with abs_tol=True"""
    return bool(np.isclose(a, b, rtol=1e-13, atol=7e-14))


def linear_wave_config(mhd):
    """input/vlct/vl.incl (+ vlct.incl for MHD)"""
    return make_config(riemann="hlld" if mhd else "hllc", recon="plm",
                       theta=2.0, mhd=mhd, courant=0.4,
                       gamma=1.6666666666666667, dfloor=1e-200, pfloor=1e-200)


def alloc_fields(cfg, n, g, passive=()):
    names = ["density", "velocity_x", "velocity_y", "velocity_z",
             "total_energy", "pressure"]
    if cfg.dual_energy:
        names.append("internal_energy")
    if cfg.mhd_choice == 1:
        names += ["bfield_x", "bfield_y", "bfield_z",
                  "bfieldi_x", "bfieldi_y", "bfieldi_z"]
    f = {k: np.zeros(abi.field_shape(k, *n, *g)) for k in names}
    for k in passive:
        f[k] = np.zeros(abi.field_shape("density", *n, *g))
    return f


def linear_wave_setup(name, N, mhd=True, positive_vel=True):
    """Domain [0,3]x[0,1.5]^2 with (2N,N,N) cells, one periodic block."""
    cfg = linear_wave_config(mhd)
    n, g = (2 * N, N, N), (3, 3, 3)
    d = (3.0 / n[0], 1.5 / n[1], 1.5 / n[2])
    f = alloc_fields(cfg, n, g)
    blk = oracle.numpy_block(f, n, g, d)
    wave_type, t_final = (MHD_WAVES if mhd else HD_WAVES)[name]
    oracle.ic_inclined_wave(blk, (0.0, 0.0, 0.0), cfg.gamma, wave_type, ALPHA,
                            BETA, 1e-6, 1.0, positive_vel)
    return cfg, f, blk, n, g, d, t_final


def pressure_of(cfg, f):
    """EnzoComputePressure formula (what an Output of "pressure" holds)."""
    gm1 = cfg.gamma - 1.0
    if cfg.dual_energy:
        return gm1 * f["density"] * f["internal_energy"]
    ke = 0.5 * (f["velocity_x"] * f["velocity_x"] +
                f["velocity_y"] * f["velocity_y"] +
                f["velocity_z"] * f["velocity_z"])
    me = 0.0
    if cfg.mhd_choice == 1:
        me = 0.5 * (f["bfield_x"] * f["bfield_x"] + f["bfield_y"] * f["bfield_y"]
                    + f["bfield_z"] * f["bfield_z"])
    return gm1 * (f["density"] * (f["total_energy"] - ke) - me)


def snapshot(cfg, f, g):
    """The active-zone fields an Output would dump."""
    gx, gy, gz = g
    out = {}
    names = ["density", "velocity_x", "velocity_y", "velocity_z",
             "total_energy"]
    if cfg.mhd_choice == 1:
        names += ["bfield_x", "bfield_y", "bfield_z"]
    p = pressure_of(cfg, f)
    for k in names:
        out[k] = f[k][gz:-gz, gy:-gy, gx:-gx].copy()
    out["pressure"] = p[gz:-gz, gy:-gy, gx:-gx].copy()
    return out


def l1_error_norm(a, b, fields, res):
    """tools/l1_error_norm.py:195-234 in "sim" mode with -n res"""
    resid = [np.sum(np.abs(a[k] - b[k])) / float(res ** 3) for k in fields]
    return float(np.sqrt(np.sum(np.square(np.array(resid)))))


def evolve(method, blk, t_stop, refresh, dump_times=(), max_cycles=100000):
    """Cycle loop: timestep -> (clip) -> refresh -> compute -> advance."""
    t, cycle, dts = 0.0, 0, []
    while t < t_stop and cycle < max_cycles:
        dt = method.timestep(blk)
        for td in dump_times:                  # io_ScheduleList.cpp:133-159
            if t < td < t + dt:
                dt = td - t
                break
        dt = min(dt, t_stop - t)               # control_stopping.cpp:86
        refresh(blk)
        method.compute(blk, dt)
        t += dt
        cycle += 1
        dts.append(dt)
    return dts


LINWAVE_FIELDS_MHD = ["density", "velocity_x", "velocity_y", "velocity_z",
                      "pressure", "bfield_x", "bfield_y", "bfield_z"]
LINWAVE_FIELDS_HD = ["density", "velocity_x", "velocity_y", "velocity_z",
                     "pressure"]


def run_linear_wave(name, N, mhd=True, kind="oracle", positive_vel=True,
                    method=None, refresh=None):
    cfg, f, blk, n, g, d, t_final = linear_wave_setup(name, N, mhd,
                                                      positive_vel)
    own = method is None
    if own:
        method = oracle.CpuMethod(cfg, g, kind=kind)
    if refresh is None:
        def refresh(b):
            oracle.refresh_periodic(b, 0)
    s0 = snapshot(cfg, f, g)
    dts = evolve(method, blk, t_final, refresh, dump_times=(0.0, t_final))
    s1 = snapshot(cfg, f, g)
    if own:
        method.close()
    fields = LINWAVE_FIELDS_MHD if mhd else LINWAVE_FIELDS_HD
    return l1_error_norm(s0, s1, fields, N), dts, f


# ---------------------------------------------------------------------------
# axis-aligned shock tubes (input/vlct/MHD_shock_tube/*.in,
# input/vlct/run_MHD_shock_tube_test.py, tools/l1_error_norm.py "table" mode)
# ---------------------------------------------------------------------------
import os as _os

GOLDEN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
# input/vlct/run_MHD_shock_tube_test.py:82-87 (x, y, z)
GOLDEN_RJ2A = (0.012523489882320429, 0.012523489882320308, 0.012523489882320315)
RJ2A_FIELDS = ["density", "velocity_x", "velocity_y", "velocity_z", "pressure",
               "bfield_x", "bfield_y", "bfield_z"]


def load_reference_table(name):
    """tools/l1_error_norm.py:412-471: '#key = value' header lines, then a
    CSV with a header row"""
    path = _os.path.join(GOLDEN_DIR, name)
    skip = 0
    with open(path) as fh:
        for line in fh:
            if line.startswith("#"):
                skip += 1
            else:
                break
    rec = np.genfromtxt(path, skip_header=skip, comments="#", delimiter=",",
                        names=True, encoding="utf-8")
    return {k: rec[k] for k in rec.dtype.names}


def rj2a_setup(axis, ncells=256):
    """method_vlct_{x,y,z}_rj2a_N256.in: 256x4x4 cells on the unit cube, one
    block, outflow along the tube, periodic across it; vlct.incl parameters."""
    cfg = make_config(riemann="hlld", recon="plm", theta=2.0, mhd=True,
                      courant=0.4, gamma=1.6666666666666667, dfloor=1e-200,
                      pfloor=1e-200)
    n = [4, 4, 4]
    n[axis] = ncells
    n, g = tuple(n), (3, 3, 3)
    d = tuple(1.0 / n[a] for a in range(3))
    f = alloc_fields(cfg, n, g)
    blk = oracle.numpy_block(f, n, g, d)
    oracle.ic_shock_tube(blk, (0.0, 0.0, 0.0), cfg.gamma, "rj2a",
                         aligned_ax=axis)
    return cfg, f, blk, n, g, d, 0.2


def table_l1_norm(snap, table, axis, fields):
    """compare_to_1D_reference (tools/l1_error_norm.py:537-628) with --permute:
    the table's vectors are rotated onto the tube's axis, the table is
    broadcast across the tube, residuals are normalised by the cell count."""
    names = "xyz"
    ref = dict(table)
    for pre in ("velocity", "bfield"):
        if all(f"{pre}_{c}" in table for c in names):
            comps = [table[f"{pre}_{c}"] for c in names]
            for k in range(3):
                ref[f"{pre}_{names[(axis + k) % 3]}"] = comps[k]
    shape = [1, 1, 1]
    shape[2 - axis] = -1
    resid = [np.sum(np.abs(snap[k] - ref[k].reshape(shape))) / float(snap[k].size)
             for k in fields]
    return float(np.sqrt(np.sum(np.square(np.array(resid)))))


def run_rj2a(axis, method=None, refresh=None, kind="oracle"):
    cfg, f, blk, n, g, d, t_final = rj2a_setup(axis)
    own = method is None
    if own:
        method = oracle.CpuMethod(cfg, g, kind=kind)
    if refresh is None:
        periodic_axes = 7 & ~(1 << axis)

        def refresh(b):
            # refresh, then Block::update_boundary_ (control_refresh.cpp:229-232)
            oracle.refresh_periodic(b, 0, periodic_axes)
            oracle.boundary(b, axis, 0, "outflow")
            oracle.boundary(b, axis, 1, "outflow")
    dts = evolve(method, blk, t_final, refresh, dump_times=(t_final,))
    if own:
        method.close()
    return cfg, f, g, dts


# ---------------------------------------------------------------------------
# Sod shock tube with dual energy in a Mach-10 frame
# (input/vlct/dual_energy_shock_tube/*.in, run_dual_energy_shock_tube_test.py)
# ---------------------------------------------------------------------------
GOLDEN_SOD_DE = 0.02605861216339738   # run_dual_energy_shock_tube_test.py:64
SOD_FIELDS = ["density", "velocity_x", "velocity_y", "velocity_z"]
SOD_BKG_VELOCITY, SOD_OFFSET = 11.875, 2.96875


def sod_de_setup(axis, flipped=False):
    """method_vlct_sod_{x,y,z}_de_M10[_reverse].in: 508x4x4 cells of width
    1/128, MHD solver (hlld + CT) with B = 0, gamma 1.4, dual energy "modern"
    with eta = 0.00769, the whole tube moving at 11.875 along its axis. (The
    reference cuts the tube into 4 blocks; one block gives the same cells.)"""
    cfg = make_config(riemann="hlld", recon="plm", theta=2.0, mhd=True,
                      courant=0.4, gamma=1.4, dual_energy=True, eta=0.00769,
                      dfloor=1e-200, pfloor=1e-200)
    n = [4, 4, 4]
    n[axis] = 508
    n, g, d = tuple(n), (3, 3, 3), (0.0078125,) * 3
    lower = [0.0, 0.0, 0.0]
    if flipped:
        lower[axis] = -SOD_OFFSET
    f = alloc_fields(cfg, n, g)
    blk = oracle.numpy_block(f, n, g, d)
    oracle.ic_shock_tube(blk, tuple(lower), cfg.gamma, "sod", aligned_ax=axis,
                         axis_velocity=SOD_BKG_VELOCITY, flipped=flipped)
    return cfg, f, blk, n, g, d, 0.25


def sod_de_l1_norm(snap, axis, flipped=False):
    """compare_to_1D_reference with --permute, --bkg_velocity, --offset_soln
    (and --reverse for the flipped tube): the 128-cell table covers the part
    of the domain the tube has drifted into."""
    table = load_reference_table("sod_shock_tube_t0.25_res128.csv")
    names = "xyz"
    ref = {k: v.copy() for k, v in table.items()}
    for c in names:
        ref.setdefault("velocity_" + c, np.zeros_like(table["density"]))
    bkg = [SOD_BKG_VELOCITY, 0.0, 0.0]
    if flipped:                       # reverse_1D_soln (l1_error_norm.py:519-535)
        for k in ref:
            if k != "x":
                ref[k] = ref[k][::-1]
        for c in names:
            ref["velocity_" + c] = -1.0 * ref["velocity_" + c]
        bkg[0] = -SOD_BKG_VELOCITY
    comps = [ref["velocity_" + c] for c in names]
    for k in range(3):
        ref["velocity_" + names[(axis + k) % 3]] = comps[k]
    if axis == 1:                     # testing_utils.py:219-231
        bkg = bkg[2:] + bkg[:2]
    elif axis == 2:
        bkg = bkg[1:] + bkg[:1]
    for i, c in enumerate(names):
        ref["velocity_" + c] = ref["velocity_" + c] + bkg[i]
    sl = [slice(None)] * 3
    sl[2 - axis] = slice(0, 128) if flipped else slice(380, 508)
    shape = [1, 1, 1]
    shape[2 - axis] = 128
    resid = []
    for k in SOD_FIELDS:
        a = snap[k][tuple(sl)]
        resid.append(np.sum(np.abs(a - ref[k].reshape(shape))) / float(a.size))
    return float(np.sqrt(np.sum(np.square(np.array(resid)))))


def outflow_refresh(axis, refresh_periodic, boundary):
    """refresh of a tube: periodic across it, then outflow at its two ends"""
    periodic_axes = 7 & ~(1 << axis)

    def refresh(b):
        refresh_periodic(b, periodic_axes)
        boundary(b, axis, 0, "outflow")
        boundary(b, axis, 1, "outflow")
    return refresh


def run_sod_de(axis, flipped=False, kind="oracle"):
    cfg, f, blk, n, g, d, t_final = sod_de_setup(axis, flipped)
    method = oracle.CpuMethod(cfg, g, kind=kind)
    refresh = outflow_refresh(axis, lambda b, ax: oracle.refresh_periodic(b, 0, ax),
                              oracle.boundary)
    dts = evolve(method, blk, t_final, refresh, dump_times=(t_final,))
    method.close()
    return cfg, f, g, dts


# ---------------------------------------------------------------------------
# passive scalar carried by a sound wave
# (input/vlct/passive_advect_sound_wave/*.in, run_passive_advect_sound_test.py)
# ---------------------------------------------------------------------------
GOLDEN_PASSIVE_SOUND = 6.918011605252798e-08
PASSIVE_FIELDS = ["density", "velocity_x", "velocity_y", "velocity_z",
                  "total_energy", "bfield_x", "bfield_y", "bfield_z", "red"]


def passive_sound_setup(axis):
    """16x4x4 cells, MHD solver with B = 0, one "color" field "red" whose mass
    fraction is a sine a quarter wavelength out of phase with the wave."""
    cfg = make_config(riemann="hlld", recon="plm", theta=2.0, mhd=True,
                      courant=0.4, gamma=1.6666666666666667, n_passive=1,
                      dfloor=1e-200, pfloor=1e-200)
    n = [4, 4, 4]
    n[axis] = 16
    n, g, d = tuple(n), (3, 3, 3), (1.0 / 16,) * 3
    f = alloc_fields(cfg, n, g, passive=("red",))
    blk = oracle.numpy_block(f, n, g, d, ("red",))
    x = (np.arange(n[axis] + 2 * g[axis]) - g[axis] + 0.5) * d[axis]
    shape = [1, 1, 1]
    shape[2 - axis] = -1
    X = x.reshape(shape) * np.ones(f["density"].shape)
    s = np.sin(2.0 * np.pi * X)
    f["density"][...] = 1.0 + 1.e-6 * s
    f["velocity_" + "xyz"[axis]][...] = -1.e-6 * s
    f["total_energy"][...] = 1.5 * (0.6 + 1.e-6 * s) / (1.0 + 1.e-6 * s)
    f["red"][...] = ((0.1 + 1.e-7 * np.sin(2.0 * np.pi * (X - 0.25)))
                     * (1.0 + 1.e-6 * s))
    return cfg, f, blk, n, g, d, 1.0


def passive_snapshot(f, g):
    gx, gy, gz = g
    return {k: f[k][gz:-gz, gy:-gy, gx:-gx].copy() for k in PASSIVE_FIELDS}


def passive_l1_norm(s0, s1):
    resid = [np.sum(np.abs(s0[k] - s1[k])) / float(s0[k].size)
             for k in PASSIVE_FIELDS]
    return float(np.sqrt(np.sum(np.square(np.array(resid)))))


# ---------------------------------------------------------------------------
# cloud in a wind with dual energy: symmetry test
# (input/vlct/dual_energy_cloud/*.in, run_dual_energy_cloud_test.py)
# ---------------------------------------------------------------------------
# max tolerated asymmetry, run_dual_energy_cloud_test.py:80-85
CLOUD_MAX_ASYM = {"hlld": 7.3e-13, "hllc": 4.6e-13, "hlle": 3.2e-13}
# initial_cloud_HD.in:26-63
CLOUD = dict(subsample_n=2, cloud_radius=1.0, center=(0.0, 0.0, 0.0),
             cloud_density=1.610075932356949e+01,
             wind_density=8.944866290871940e-02,
             wind_velocity=1.341500614584355e+01,
             wind_total_energy=1.619661509037183e+02,
             wind_internal_energy=7.198495595720811e+01)
# Boundary:hydro_upwind:value (initial_cloud_HD.in:83-94); metal_density and
# cloud_dye are not in a "color" group there, i.e. not fields of the method
CLOUD_INFLOW = {"density": 8.944866290871940e-02,
                "velocity_x": 1.341500614584355e+01,
                "velocity_y": 0.0, "velocity_z": 0.0,
                "total_energy": 1.619661509037183e+02,
                "internal_energy": 7.198495595720811e+01}
CLOUD_INFLOW_B = {k: 0.0 for k in ("bfield_x", "bfieldi_x", "bfield_y",
                                   "bfieldi_y", "bfield_z", "bfieldi_z")}
CLOUD_LOWER, CLOUD_T_STOP = (-2.0, -2.0, -2.0), 0.0625


def cloud_config(solver):
    """{hlld,hlle}_cloud.in: vlct_de.incl (CT with B = 0); hllc_cloud.in:
    vl_de.incl (no_bfield); theta_limiter 1.5, floors from initial_cloud_HD.in"""
    return make_config(riemann=solver, recon="plm", theta=1.5,
                       mhd=(solver != "hllc"), courant=0.4,
                       gamma=1.6666666666666667, dual_energy=True, eta=0.001,
                       dfloor=1.e-15, pfloor=1.0e-30)


def cloud_setup(solver):
    """32^3 cells on [-2,2]^3 (8 per cloud radius), one block."""
    cfg = cloud_config(solver)
    n, g, d = (32, 32, 32), (3, 3, 3), (0.125,) * 3
    f = alloc_fields(cfg, n, g)
    blk = oracle.numpy_block(f, n, g, d)
    oracle.ic_cloud(blk, CLOUD_LOWER, **CLOUD)
    return cfg, f, blk, n, g, d, CLOUD_T_STOP


def cloud_refresh(mhd, boundary, boundary_inflow):
    """Block::update_boundary_ over Boundary:list = [downwind, hydro_upwind,
    yedge, zedge (, bfield_upwind)] (initial_cloud_HD.in:66-112,
    initial_cloud_MHD.in:27-40): every Boundary object in list order, each over
    the faces it applies to (Cello/mesh_Block.cpp:1057-1077)."""
    def refresh(b):
        boundary(b, 0, 1, "outflow")
        boundary_inflow(b, 0, 0, CLOUD_INFLOW)
        for axis in (1, 2):
            boundary(b, axis, 0, "outflow")
            boundary(b, axis, 1, "outflow")
        if mhd:
            boundary_inflow(b, 0, 0, CLOUD_INFLOW_B)
    return refresh


def cloud_boundary_list(mhd):
    """the same Boundary:list in the form enzo_e_b200.domain.Domain takes"""
    out = [dict(type="outflow", axis=0, face=1),
           dict(type="inflow", axis=0, face=0, values=CLOUD_INFLOW),
           dict(type="outflow", axis=1), dict(type="outflow", axis=2)]
    if mhd:
        out.append(dict(type="inflow", axis=0, face=0, values=CLOUD_INFLOW_B))
    return out


def slice_asym(grid, slice_ax, slice_ind, flip_across):
    """run_dual_energy_cloud_test.py:27-54; grid is indexed (x, y, z)"""
    if slice_ax == 'x':
        slice_arr = grid[slice_ind, :, :]
        flipped = (grid[slice_ind, :, ::-1] if flip_across == 'y'
                   else grid[slice_ind, ::-1, :])
    elif slice_ax == 'y':
        slice_arr = grid[:, slice_ind, :]
        flipped = (grid[:, slice_ind, ::-1] if flip_across == 'x'
                   else grid[::-1, slice_ind, :])
    else:
        slice_arr = grid[:, :, slice_ind]
        flipped = (grid[:, ::-1, slice_ind] if flip_across == 'x'
                   else grid[::-1, :, slice_ind])
    return float(np.sum(np.abs((slice_arr - flipped) / slice_arr)))


def cloud_asymmetries(f, g):
    """check_cloud_asym (run_dual_energy_cloud_test.py:56-77): the three
    slices of the density the reference test inspects"""
    gx, gy, gz = g
    grid = f["density"][gz:-gz, gy:-gy, gx:-gx].transpose(2, 1, 0)
    return [slice_asym(grid, *args) for args in
            (('z', 16, 'x'), ('y', 16, 'x'), ('x', 16, 'z'))]


def run_cloud(solver, kind="oracle"):
    cfg, f, blk, n, g, d, t_stop = cloud_setup(solver)
    method = oracle.CpuMethod(cfg, g, kind=kind)
    refresh = cloud_refresh(cfg.mhd_choice == 1, oracle.boundary,
                            oracle.boundary_inflow)
    dts = evolve(method, blk, t_stop, refresh, dump_times=(t_stop,))
    method.close()
    return cfg, f, g, dts
