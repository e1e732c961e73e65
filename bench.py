#!/usr/bin/env python
"""bench.py -- throughput of the VL+CT hot path (EnzoMethodMHDVlct) on B200.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU path)

Headline workload (config.workload): BASELINE.json configs[2], the 3-D
Orszag-Tang vortex on 512^3 cells per GPU, MHD VL+CT with PLM (theta 1.5) +
HLLD + CT, fp64, one periodic brick per GPU (weak scaling: the global grid
grows with N; N = 8 is a 2x2x2 arrangement = global 1024^3 with x, y and z
ghost exchanges). `--workload` selects another BASELINE configuration as the
headline; by default the other ones are ALSO measured (short runs, same
process) and reported under "workloads":

    turbulence  configs[3]  decaying MHD turbulence, 512^3 per GPU (N = 8:
                            1024^3 as 2x2x2 bricks), all HLLD regions populated
                            along every axis
    ot_s4       configs[4]  the headline problem + 4 passive scalars
    sod256      configs[1]  3-D Sod, hydro, PLM + HLLC + dual energy, 256^3 (N = 1)
    fastwave64  configs[0]  fast magnetosonic linear wave, 128x64x64 (N = 1)

A "step" is one full simulation cycle of the path on one block per GPU:
    timestep() -> min over ranks -> ghost refresh (device wrap / NCCL exchange)
    -> compute() (both VL stages incl. CT)
metric = cell-updates/s = active cells of all ranks * steps / seconds, timed on
the device with CUDA events, max over ranks.

Prints ONE JSON line (rank 0). `roofline` is SURVEY 8(d)'s step figure
(value x algorithmic bytes per cell-update / measured HBM peak); the per-kernel
numbers are under `roofline_families`, the FP64-issue roofline under
`roofline_fp64`. See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cell-updates/sec (fp64 VL+CT MHD)"
UNIT = "cell-updates/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback
# FP64 issue rate of a B200: one DP warp-instruction per 2 cycles per SM
# sub-partition (scripts/microbench/dp_pipe.cu): 148 x 4 x 0.5 x 1.965 GHz
FP64_WARP_INST_PER_S = 148 * 4 * 0.5 * 1.965e9


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# --------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,"
             "clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_median": statistics.median(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------
PARAMS = {
    "Method:mhd_vlct:mhd_choice": "constrained_transport",
    "Method:mhd_vlct:riemann_solver": "hlld",
    "Method:mhd_vlct:reconstruct_method": "plm",
    "Method:mhd_vlct:theta_limiter": 1.5,
    "Method:mhd_vlct:time_scheme": "vl",
    "Method:mhd_vlct:courant": 0.3,
    "Physics:fluid_props:eos:gamma": 5.0 / 3.0,
    "Physics:fluid_props:floors:density": 1e-200,
    "Physics:fluid_props:floors:pressure": 1e-200,
}
HYDRO_DE = {
    "Method:mhd_vlct:mhd_choice": "no_bfield",
    "Method:mhd_vlct:riemann_solver": "hllc",
    "Method:mhd_vlct:reconstruct_method": "plm",
    "Method:mhd_vlct:theta_limiter": 1.5,
    "Method:mhd_vlct:time_scheme": "vl",
    "Method:mhd_vlct:courant": 0.3,
    "Physics:fluid_props:eos:gamma": 1.4,
    "Physics:fluid_props:dual_energy:type": "modern",
    "Physics:fluid_props:dual_energy:eta": 1e-3,
    "Physics:fluid_props:floors:density": 1e-200,
    "Physics:fluid_props:floors:pressure": 1e-200,
}
LINWAVE = dict(PARAMS, **{"Method:mhd_vlct:theta_limiter": 2.0,
                          "Method:mhd_vlct:courant": 0.4,
                          "Physics:fluid_props:eos:gamma": 1.6666666666666667})
GHOST = (3, 3, 3)

# name -> description of a BASELINE configuration. bytes = SURVEY 8(d)'s
# ALGORITHMIC bytes per cell-update: MHD 416 B (+40 B per passive scalar),
# hydro + dual energy 240 B. dp_inst = fp64 warp-instructions per 32
# cell-updates (ncu opmix of the shipped kernels, profiles/; data dependent).
WORKLOADS = {
    "ot": dict(config="configs[2]", title="Orszag-Tang vortex", params=PARAMS,
               n_passive=0, bytes=416.0, multi_gpu=True,
               solver="MHD VL+CT: PLM(theta=1.5) + HLLD + CT, periodic"),
    "turbulence": dict(config="configs[3]", title="decaying MHD turbulence "
                       "(solenoidal modes |k|<=3, seed 20240517, rms Mach 0.5, beta 2)",
                       params=PARAMS, n_passive=0, bytes=416.0, multi_gpu=True,
                       solver="MHD VL+CT: PLM(theta=1.5) + HLLD + CT, periodic"),
    "ot_s4": dict(config="configs[4]", title="Orszag-Tang vortex + 4 passive scalars",
                  params=PARAMS, n_passive=4, bytes=416.0 + 4 * 40.0, multi_gpu=True,
                  solver="MHD VL+CT: PLM(theta=1.5) + HLLD + CT + 4 scalars, periodic"),
    "sod256": dict(config="configs[1]", title="3-D Sod problem", params=HYDRO_DE,
                   n_passive=0, bytes=240.0, multi_gpu=False, size=256,
                   solver="hydro VL: PLM(theta=1.5) + HLLC + dual energy, "
                          "outflow along x, periodic across"),
    "fastwave64": dict(config="configs[0]", title="fast magnetosonic linear wave "
                       "(input/vlct/MHD_linear_wave, N = 64)", params=LINWAVE,
                       n_passive=0, bytes=416.0, multi_gpu=False, size=64,
                       solver="MHD VL+CT: PLM(theta=2) + HLLD + CT, periodic"),
}


def local_shape(name, size):
    if name == "fastwave64":
        return (2 * size, size, size)
    return (size, size, size)


def workload_config(name, size, world, grid, layout):
    w = WORKLOADS[name]
    n = local_shape(name, size)
    glob = tuple(n[a] * grid[a] for a in range(3))
    return {"workload": f"{w['title']}, {n[0]}x{n[1]}x{n[2]} cells per GPU "
                        f"(global {glob[0]}x{glob[1]}x{glob[2]}), {w['solver']} "
                        f"[BASELINE {w['config']}]",
            "name": name, "cells_per_gpu": n[0] * n[1] * n[2], "ghost_depth": 3,
            "riemann_solver": w["params"]["Method:mhd_vlct:riemann_solver"],
            "reconstruct_method": "plm", "n_passive": w["n_passive"],
            "courant": w["params"]["Method:mhd_vlct:courant"],
            "gamma": w["params"]["Physics:fluid_props:eos:gamma"],
            "decomposition": f"{grid[0]}x{grid[1]}x{grid[2]} {layout}",
            "algorithmic_bytes_per_cell_update": w["bytes"],
            "step": "ghost refresh + compute + the next cycle's timestep (folded into the last update kernel: vlct_compute_and_timestep_dev) + its min-reduce over ranks; dt stays on the device; with a z split the z ghost exchange overlaps the interior update",
            "l2_policy": "inputs (>=1.1 GB per field at 512^3, 143 MB at 256^3) exceed the 126 MB L2"
                         if n[0] * n[1] * n[2] >= 256 ** 3 else
                         "L2 flushed between timed steps (a 256 MB buffer is rewritten)"}


def make_fields(name, n_local, lower, width, global_n, dev):
    from enzo_e_b200 import problems
    w = WORKLOADS[name]
    if name in ("ot", "ot_s4"):
        return problems.orszag_tang(n_local, GHOST, lower, width, device=dev,
                                    n_passive=w["n_passive"])
    if name == "turbulence":
        return problems.turbulence(n_local, GHOST, lower, width, global_n, device=dev)
    if name == "sod256":
        return problems.hydro_sod(n_local, GHOST, lower, width, device=dev, gamma=1.4)
    if name == "fastwave64":
        return problems.inclined_wave(n_local, GHOST, lower, width, "fast",
                                      0.7297276562269663, 1.1071487177940904,
                                      device=dev, gamma=1.6666666666666667)
    raise SystemExit(f"unknown workload {name}")


def workload_widths(name, n_local, grid):
    if name == "fastwave64":      # domain 3 x 1.5 x 1.5
        return (3.0 / n_local[0], 1.5 / n_local[1], 1.5 / n_local[2])
    return tuple(1.0 / (n_local[a] * grid[a]) for a in range(3))


def workload_boundaries(name):
    if name == "sod256":
        return [{"type": "outflow", "axis": 0}]
    return None


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuArm:
    """The reference CPU path (oracle/_ref = the reference's own sources when
    present, else the oracle port) on `threads` host threads, one `block`^3
    brick of the workload per thread (the reference's natural unit: one Cello
    block per compute() call; 64^3 as BASELINE.md section 3 states)."""

    def __init__(self, name="ot", block=64, threads=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        from enzo_e_b200 import abi
        self.oracle = oracle
        self.threads = threads or cpu_threads()
        self.block = block
        self.kind = "ref" if oracle.have_ref() else "oracle"
        if self.kind == "oracle" and not oracle.have_oracle():
            oracle.build("oracle")
        w = WORKLOADS[name]
        prm = w["params"]
        # same parameters as the GPU arm, set directly on the struct so that
        # the CPU arm never loads the CUDA library
        cfg = abi.default_config()
        cfg.mhd_choice = abi.MHD_CHOICE[prm["Method:mhd_vlct:mhd_choice"]]
        cfg.riemann_solver = abi.RIEMANN[prm["Method:mhd_vlct:riemann_solver"]]
        cfg.reconstruct_method = abi.RECON["plm"]
        cfg.theta_limiter = prm["Method:mhd_vlct:theta_limiter"]
        cfg.courant = prm["Method:mhd_vlct:courant"]
        cfg.gamma = prm["Physics:fluid_props:eos:gamma"]
        cfg.density_floor = cfg.pressure_floor = 1e-200
        if "Physics:fluid_props:dual_energy:type" in prm:
            cfg.dual_energy = 1
            cfg.dual_energy_eta = prm["Physics:fluid_props:dual_energy:eta"]
        cfg.n_passive = w["n_passive"]
        self.n_passive = w["n_passive"]
        n = (block, block, block)
        # the bricks tile a corner of the 512^3 problem (cell width 1/512)
        size = 512 if w.get("size") is None else w["size"]
        width = (1.0 / size,) * 3
        per_axis = max(1, size // block)
        passive = tuple(f"passive_{k}" for k in range(w["n_passive"]))
        self.workers = []
        for t in range(self.threads):
            c = (t % per_axis, (t // per_axis) % per_axis, (t // per_axis ** 2) % per_axis)
            lower = tuple(c[a] * block * width[a] for a in range(3))
            f = {k: v.numpy().copy() for k, v in make_fields(
                name if name != "fastwave64" else "ot", n, lower, width,
                (size,) * 3, "cpu").items()}
            blk = oracle.numpy_block(f, n, GHOST, width, passive)
            self.workers.append((oracle.CpuMethod(cfg, GHOST, kind=self.kind), blk, f))

    def one_step(self, w):
        m, blk, _ = w
        dt = m.timestep(blk)
        self.oracle.refresh_periodic(blk, self.n_passive)
        m.compute(blk, dt)

    def sample(self, nsteps, threads=None):
        """`nsteps` full cycles (timestep + refresh + compute) of every brick of
        the first `threads` workers, all at once; returns (cell-updates/s, s)"""
        workers = self.workers[:threads or self.threads]

        def loop(w):
            for _ in range(nsteps):
                self.one_step(w)

        ths = [threading.Thread(target=loop, args=(w,)) for w in workers]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        elapsed = time.perf_counter() - t0
        return len(workers) * self.block ** 3 * nsteps / elapsed, elapsed

    def calibrate(self, seconds):
        """steps per sample so that one all-thread sample takes ~`seconds`"""
        _, one = self.sample(1)
        return max(1, min(200, int(seconds / max(one, 1e-6))))

    def close(self):
        for m, _, _ in self.workers:
            m.close()


def run_cpu_baseline(name, seconds_target=12.0, block=64):
    """the bench line's `cpu_baseline`: one bounded all-core sample (+ a
    single-core one) of the reference CPU path on the headline workload"""
    arm = CpuArm(name, block=block)
    nsteps = arm.calibrate(seconds_target * 0.75)
    value, elapsed = arm.sample(nsteps)
    one_core, one_elapsed = arm.sample(max(1, nsteps // 4), threads=1)
    out = {"value": value, "unit": UNIT, "cores": arm.threads,
           "kind": "reference" if arm.kind == "ref" else "port",
           "sample": f"{arm.threads} threads x one {block}^3 brick of the "
                     f"workload x {nsteps} full cycles (timestep+refresh+compute), "
                     f"{elapsed:.1f} s",
           "single_core_value": one_core,
           "single_core_sample": f"1 thread x one {block}^3 brick x "
                                 f"{max(1, nsteps // 4)} cycles, {one_elapsed:.1f} s"}
    arm.close()
    return out


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path
    (oracle/_ref) on all host cores. A "step" of this arm is one bounded
    sample: every thread advances its own 64^3 brick of the workload by a fixed
    number of cycles; value = cell-updates of the timed samples / their wall
    time, so steps x ms_per_step is the time this run really spent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    grid = _grid(args.gpus, args.layout)
    t_all = time.perf_counter()
    arm = CpuArm(name, block=args.cpu_block)
    steps = max(5, args.steps)            # median of >= 5 samples
    per = max(1.0, min(12.0, 100.0 / (steps + args.warmup + 2)))
    cycles = arm.calibrate(per)
    for _ in range(args.warmup):
        arm.sample(max(1, cycles // 2))
    samples = [arm.sample(cycles) for _ in range(steps)]
    single = [arm.sample(max(1, cycles // 2), threads=1) for _ in range(5)]
    arm.close()
    value = statistics.median(v for v, _ in samples)
    ms = 1e3 * statistics.median(t for _, t in samples)
    one_core = statistics.median(v for v, _ in single)
    size = WORKLOADS[name].get("size", args.size)
    n = local_shape(name, size)
    cells_per_step = n[0] * n[1] * n[2] * args.gpus
    sample = (f"{arm.threads} threads x one {args.cpu_block}^3 brick of the workload "
              f"x {cycles} full cycles per sample; median of {steps} samples")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(name, size, args.gpus, grid, args.layout),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.threads,
                             "kind": "reference" if arm.kind == "ref" else "port",
                             "sample": sample, "single_core_value": one_core,
                             "all_core_samples": [v for v, _ in samples],
                             "single_core_samples": [v for v, _ in single]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "step_definition": "one bounded sample (see cpu_baseline.sample); "
                               "ms_per_step is its measured wall time",
            "full_workload_ms_per_step_extrapolated": 1e3 * cells_per_step / value,
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def _grid(world, layout="bricks"):
    from enzo_e_b200.domain import proc_grid
    return proc_grid(world, slabs=(layout == "slabs"))


# Kernel families and their ALGORITHMIC bytes per processed unit (DESIGN.md
# "Measurement"): doubles that must cross HBM once per unit, MHD without dual
# energy / scalars. `units(m, name)` = units one launch processes for a block
# of m^3 cells including ghosts.
FAMILIES = {
    # per face: 8 cell fields (rho, v, etot, B) + 1 face B in, 7 fluxes out
    "k_flux": {"doubles": 16,
               "units": lambda m, name: ((m - 5) * (m - 4) ** 2 if name.endswith("plm")
                                         else (m - 1) * m * m)},
    # per cell: 15 fluxes + 3 face B + 5 start-of-step fields in, 8 fields out
    # (+ "pressure" out for k_update_cfl, the update that also does the CFL work)
    "k_update": {"doubles": 31, "units": lambda m, name: (m - 2) ** 3 * (
        32.0 / 31.0 if name.endswith("cfl") else 1.0)},
    # per cell: v, B (6) + 6 B-fluxes + 3 density fluxes in, 3 edge E out
    "k_edge_efield": {"doubles": 18, "units": lambda m, name: (m - 2) ** 3},
    # per cell: 3 edge E + 3 face B in, 3 face B out
    "k_face_bfield": {"doubles": 9, "units": lambda m, name: (m - 2) ** 3},
    # edge E + face B in one TMA-staged kernel (the edge E never reach HBM):
    # v, B (6) + 6 B-fluxes + 3 density fluxes + 3 face B in, 3 face B out
    "k_ct_tma": {"doubles": 21, "units": lambda m, name: (m - 2) ** 3},
    # per cell: 8 fields in, pressure out (k_timestep_shell: the ghost shell only)
    "k_timestep": {"doubles": 9, "units": lambda m, name: (
        m ** 3 - (m - 6) ** 3 if name.endswith("shell") else m ** 3)},
}


def family_of(name):
    for key in FAMILIES:
        if name.startswith(key):
            return key
    return None


def family_rooflines(report, m, hbm_gbs):
    """per family: total ms, launches, algorithmic GB per launch (average over
    the family's launches), achieved GB/s = bytes / CUDA-event duration"""
    fam = {}
    for name, (ms, calls) in report.items():
        key = family_of(name)
        if key is None or calls == 0:
            continue
        f = fam.setdefault(key, {"ms": 0.0, "launches": 0, "bytes": 0.0})
        f["ms"] += ms
        f["launches"] += calls
        f["bytes"] += calls * FAMILIES[key]["units"](m, name) * FAMILIES[key]["doubles"] * 8.0
    for f in fam.values():
        f["avg_launch_ms"] = f["ms"] / f["launches"]
        f["bytes_per_launch"] = f["bytes"] / f["launches"]
        f["achieved"] = f["bytes"] / (f["ms"] * 1e-3) / 1e9
        f["frac"] = f["achieved"] / hbm_gbs
    return fam


class Setup:
    """one rank's block of a workload + the Method driving it"""

    def __init__(self, name, size, domain_args, dev, layout):
        import torch
        from enzo_e_b200.domain import Domain
        from enzo_e_b200.method import EnzoMethodMHDVlct, Block
        rank, world = domain_args
        w = WORKLOADS[name]
        self.name, self.w = name, w
        self.domain = Domain(rank, world, grid=_grid(world, layout),
                             boundaries=workload_boundaries(name))
        grid = self.domain.grid
        self.n_local = local_shape(name, size)
        self.width = workload_widths(name, self.n_local, grid)
        lower = self.domain.lower_corner(self.n_local, self.width)
        global_n = tuple(self.n_local[a] * grid[a] for a in range(3))
        self.passive = tuple(f"passive_{k}" for k in range(w["n_passive"]))
        self.fields = make_fields(name, self.n_local, lower, self.width, global_n, dev)
        self.method = EnzoMethodMHDVlct(w["params"], n_passive=w["n_passive"])
        self.block = Block(self.fields, self.n_local, GHOST, self.width,
                           passive=self.passive)   # torch's current stream
        assert self.block.stream_is_current
        self.dt_dev = torch.empty(1, dtype=torch.float64, device=dev)
        self.dt_next = torch.empty(1, dtype=torch.float64, device=dev)
        self.primed = False
        self.fold = True       # fold the next cycle's timestep into the update
        self.dev = dev
        self.cells = self.n_local[0] * self.n_local[1] * self.n_local[2]
        # small blocks fit the L2: flush it between timed steps
        self.flush = None
        if self.cells < 256 ** 3:
            self.flush = torch.empty(32 * 1024 * 1024, dtype=torch.float64, device=dev)

    def step(self, overlap=True):
        # dt stays on the device (vlct_timestep_dev / vlct_compute_dev): the
        # cycles queue back to back, nothing waits for the host
        if self.flush is not None:
            self.flush.add_(1.0)
        if not self.fold:
            dt = self.method.timestep_dev(self.block, out=self.dt_dev)
            dt = self.domain.global_dt(dt, self.dev)
            self.domain.step(self.method, self.block, dt, overlap=overlap)
            return dt
        # Enzo-E's cycle is timestep -> refresh -> compute; the timestep of the
        # NEXT cycle only needs what compute leaves behind, so it is evaluated
        # inside the last update kernel (vlct_compute_and_timestep_dev). One
        # cycle = refresh + compute + next timestep + its min-reduction.
        if not self.primed:
            dt = self.method.timestep_dev(self.block, out=self.dt_dev)
            self.domain.global_dt(dt, self.dev)
            self.primed = True
        dt = self.dt_dev
        # refresh + compute; with a z split the z exchange runs under the
        # interior part of the update (Domain.step)
        self.domain.step(self.method, self.block, dt, overlap=overlap,
                         dt_next=self.dt_next)
        self.domain.global_dt(self.dt_next, self.dev)
        self.dt_dev, self.dt_next = self.dt_next, self.dt_dev
        return dt

    def close(self):
        self.method.close()
        self.fields.clear()
        self.block = None


def timed_steps(setup, stream, steps, warmup, sync_all, overlap=True):
    import torch
    for _ in range(warmup):
        setup.step(overlap)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush_ms = 0.0
    if setup.flush is not None:      # cost of the L2 flush alone, subtracted below
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(steps):
            setup.flush.add_(1.0)
        f1.record(stream)
        sync_all()
        flush_ms = f0.elapsed_time(f1)
    e0.record(stream)
    for _ in range(steps):
        dt = setup.step(overlap)
    e1.record(stream)
    sync_all()
    return e0.elapsed_time(e1) - flush_ms, dt


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the VL+CT path has no "
                         "CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pin_note = pin_rank_to_cores(local_rank, world)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        try:   # NCCL's own stream above the compute stream: the ghost slabs
            # move while the interior update runs
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    name = args.workload
    if world > 1 and not WORKLOADS[name]["multi_gpu"]:
        raise SystemExit(f"workload {name} is a single-GPU configuration")
    size = WORKLOADS[name].get("size", args.size)
    hbm_gbs, peak_kind = measured_peaks()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce_max(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        setup = Setup(name, size, (rank, world), dev, args.layout)
        method, block, domain = setup.method, setup.block, setup.domain
        grid = domain.grid
        setup.fold = not args.no_fold
        overlap = not args.no_overlap

        for _ in range(args.warmup):
            setup.step(overlap)
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = method.kernel_launches()
        elapsed_ms, dt = timed_steps(setup, stream, args.steps, 0, sync_all, overlap)
        launches = method.kernel_launches() - launches0
        clocks = sampler.stop() if rank == 0 else None
        elapsed_ms = reduce_max(elapsed_ms)
        cells = setup.cells * world * args.steps
        value = cells / (elapsed_ms * 1e-3)

        # ---- compute()-only timing and per-kernel profile (live, same process)
        sync_all()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(max(2, args.steps // 2)):
            method.compute(block, dt)
        c1.record(stream)
        sync_all()
        compute_ms = c0.elapsed_time(c1) / max(2, args.steps // 2)

        # per-kernel profile with whole-block launches (the overlapped step
        # splits every launch in three, which would not be "one launch" of the
        # roofline accounting)
        method.profile(True)
        nprof = 4
        for _ in range(nprof):
            setup.step(overlap=False)
        sync_all()
        report = method.profile_report()
        method.profile(False)
        last_dt = float(dt.item())
        scratch_gb = method.scratch_bytes() / 1e9

    b_alg = WORKLOADS[name]["bytes"]
    groups = dict(report)
    total_prof_ms = sum(ms for ms, _ in groups.values())
    m = setup.n_local[0] + 2 * GHOST[0]
    fam = family_rooflines(report, m, hbm_gbs) if name in ("ot", "turbulence") else {}
    step_gbs = value * b_alg / 1e9 / world
    # the path's roofline is the whole step against HBM
    roofline = {"bound": "hbm", "kernel": "step (all kernels of one cycle)",
                "achieved": step_gbs, "peak": hbm_gbs, "peak_kind": peak_kind,
                "unit": "GB/s", "frac": step_gbs / hbm_gbs, "traffic": None,
                "bytes_per_cell_update": b_alg,
                "algorithmic_bytes_per_launch": b_alg * setup.cells,
                "avg_launch_ms": elapsed_ms / args.steps,
                "note": "per GPU; achieved = cell-updates/s x algorithmic bytes per "
                        "cell-update (SURVEY 8d; a 'launch' = one step); traffic is "
                        "not measured in this run (ncu per-kernel DRAM bytes: profiles/)"}
    roofline_families = {k: {"achieved": v["achieved"], "frac": v["frac"],
                             "avg_launch_ms": v["avg_launch_ms"],
                             "algorithmic_bytes_per_launch": v["bytes_per_launch"],
                             "share_of_step": v["ms"] / total_prof_ms}
                         for k, v in sorted(fam.items())}
    roofline_fp64 = fp64_roofline(name, value / world)
    kernels = {k: {"ms_per_step": ms / nprof, "launches_per_step": calls / nprof}
               for k, (ms, calls) in sorted(groups.items())}

    # ---- end-to-end through the C ABI with HOST (pinned) buffers ----------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, setup, world, dev, pin_note)
    n_local, width = setup.n_local, setup.width
    setup.close()
    del setup, method, block
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, short runs ---------------------
    extras = {}
    if not args.no_extras:
        for other in ("turbulence", "ot_s4", "sod256", "fastwave64", "ot"):
            if other == name or (world > 1 and not WORKLOADS[other]["multi_gpu"]):
                continue
            extras[other] = run_extra(other, args, rank, world, dev, stream,
                                      sync_all, reduce_max, hbm_gbs)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = run_cpu_baseline(name, seconds_target=args.cpu_seconds,
                                        block=args.cpu_block)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(name, size, world, grid, args.layout),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "roofline_fp64": roofline_fp64,
                "roofline_families": roofline_families,
                "cpu_baseline": cpu_baseline,
                "workloads": extras,
                "compute_only_ms": compute_ms,
                "compute_only_value": n_local[0] * n_local[1] * n_local[2] / (compute_ms * 1e-3),
                "kernels": kernels, "last_dt": last_dt,
                "scratch_gb": scratch_gb}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def fp64_roofline(name, value_per_gpu):
    """fp64 warp-instructions per second against the issue rate of the FP64
    pipe. The instruction count per cell-update is the ncu opmix of the shipped
    kernels (profiles/dp_inst.json, per workload; data dependent), not a live
    measurement."""
    try:
        with open(os.path.join(ROOT, "profiles", "dp_inst.json")) as fh:
            table = json.load(fh)
        per32 = float(table[name]["dp_warp_inst_per_32_cell_updates"])
        src = table[name].get("source")
    except Exception:
        return None
    achieved = value_per_gpu / 32.0 * per32
    return {"bound": "fp64", "achieved": achieved, "peak": FP64_WARP_INST_PER_S,
            "unit": "fp64 warp-instructions/s", "frac": achieved / FP64_WARP_INST_PER_S,
            "dp_warp_inst_per_32_cell_updates": per32, "source": src,
            "note": "peak = 148 SMs x 4 sub-partitions x 1 DP warp-instruction per "
                    "2 cycles x 1.965 GHz (the non-FMA arithmetic bit-parity needs "
                    "counts one flop per lane per instruction)"}


def run_extra(name, args, rank, world, dev, stream, sync_all, reduce_max, hbm_gbs):
    """a short device-resident run of another BASELINE configuration"""
    import torch
    size = WORKLOADS[name].get("size", args.size)
    try:
        with torch.cuda.stream(stream):
            setup = Setup(name, size, (rank, world), dev, args.layout)
            setup.fold = not args.no_fold
            steps = max(3, min(args.steps, 5))
            ms, dt = timed_steps(setup, stream, steps, 3, sync_all,
                                 overlap=not args.no_overlap)
            ms = reduce_max(ms)
            finite = bool(torch.isfinite(setup.fields["density"]).all())
            value = setup.cells * world * steps / (ms * 1e-3)
            out = {"value": value, "unit": UNIT, "ms_per_step": ms / steps,
                   "steps": steps, "warmup": 3,
                   "config": workload_config(name, size, world, setup.domain.grid,
                                             args.layout),
                   "roofline_frac": value / world * WORKLOADS[name]["bytes"] / 1e9 / hbm_gbs,
                   "last_dt": float(dt.item()), "finite": finite}
            setup.close()
        del setup
        torch.cuda.empty_cache()
        return out
    except Exception as exc:   # an extra must never take the headline line down
        return {"error": repr(exc)}


def pin_rank_to_cores(local_rank, world):
    """give every rank its own slice of the host cores (no migration of the
    thread that feeds the copy engines); returns a note for the JSON line"""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if world <= 1 or len(cores) < 2 * world:
            return f"{len(cores)} cores, not pinned"
        per = len(cores) // world
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return f"rank pinned to {per} of {len(cores)} cores"
    except Exception as exc:
        return f"not pinned ({exc})"


def run_e2e(args, setup, world, dev, pin_note):
    """Same metric through the C ABI with HOST buffers: every step copies the
    block's fields pinned-host -> device, runs compute + timestep, and copies
    the results back, all inside the library's calls. Headline:
    vlct_compute_and_timestep (one upload + one download per cycle); beside
    it the two separate calls vlct_timestep + vlct_compute."""
    import torch
    import torch.distributed as dist
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block

    host = {}
    for k, v in setup.fields.items():
        t = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        t.copy_(v)
        host[k] = t
    torch.cuda.synchronize(dev)
    host_np = {k: v.numpy() for k, v in host.items()}
    w = setup.w
    m2 = EnzoMethodMHDVlct(w["params"], n_passive=w["n_passive"])
    hb = Block(host_np, setup.n_local, GHOST, setup.width, passive=setup.passive)
    steps = max(2, min(args.steps, 4))
    cells = setup.cells * world * steps
    refresh_host = None   # the host code's own refresh is not part of the path

    def timed(fused):
        dt = m2.timestep(hb)
        if fused:
            dt = m2.compute_and_timestep(hb, dt)
        else:
            m2.compute(hb, dt)
        if world > 1:
            dist.barrier()
        h2d0, d2h0 = m2.staged_bytes()
        t0 = time.perf_counter()
        for _ in range(steps):
            if fused:
                dt = m2.compute_and_timestep(hb, dt)
            else:
                dt = m2.timestep(hb)
                m2.compute(hb, dt)
        elapsed = time.perf_counter() - t0
        h2d1, d2h1 = m2.staged_bytes()
        t = torch.tensor([elapsed], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
        return {"value": cells / elapsed, "unit": UNIT,
                "h2d_bytes_per_step": (h2d1 - h2d0) // steps,
                "d2h_bytes_per_step": (d2h1 - d2h0) // steps,
                "steps": steps, "ms_per_step": 1e3 * elapsed / steps,
                "pcie_gbs_h2d": (h2d1 - h2d0) / elapsed / 1e9,
                "pcie_gbs_d2h": (d2h1 - d2h0) / elapsed / 1e9}

    out = timed(True)
    out["path"] = ("vlct_compute_and_timestep with mem_space=HOST (pinned host "
                   "arrays): per cycle ONE upload of compute's inputs and ONE "
                   "download of its outputs + pressure, as a z-pass pipeline of "
                   "H2D / kernels / D2H; the CFL kernel of the next cycle runs on "
                   "the device copy behind the update (include/vlct.h)")
    two = timed(False)
    two["path"] = ("vlct_timestep + vlct_compute as two calls (each uploads its "
                   "inputs and downloads its outputs)")
    out["two_calls"] = two
    out["host_cores"] = pin_note
    out["limiter"] = ("PCIe: both directions run concurrently at the rates in "
                      "pcie_gbs_*; at N > 1 all ranks share one host memory system "
                      "(every GPU of the box hangs off NUMA node 0)")
    m2.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ot", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=512,
                    help="cells per axis per GPU (workloads without a fixed size)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-block", type=int, default=64)
    ap.add_argument("--layout", default="bricks", choices=["slabs", "bricks"],
                    help="N > 1: bricks (1x1x2, 1x2x2, 2x2x2; default) or z slabs (1x1xN)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: exchange ghosts before the update instead of under it")
    ap.add_argument("--no-fold", action="store_true",
                    help="separate timestep kernel every cycle instead of the CFL "
                         "work folded into the last update (vlct_compute_and_timestep_dev)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the short runs of the other BASELINE configurations")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
