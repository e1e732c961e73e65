#!/usr/bin/env python
"""bench.py -- throughput of the VL+CT hot path (EnzoMethodMHDVlct) on B200.

    python bench.py --gpus N --steps K --warmup W            (our CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference CPU path)

Workload (config.workload): BASELINE.json configs[2], the 3-D Orszag-Tang
vortex on 512^3 cells per GPU, MHD VL+CT with PLM (theta 1.5) + HLLD + CT,
fp64, one periodic brick per GPU (weak scaling: the global grid grows with N).

A "step" is one full simulation cycle of the path on one block per GPU:
    timestep() -> min over ranks -> ghost refresh (device wrap / NCCL exchange)
    -> compute() (both VL stages incl. CT)
metric = cell-updates/s = active cells of all ranks * steps / seconds, timed on
the device with CUDA events, max over ranks.

Prints ONE JSON line (rank 0). See DESIGN.md "Measurement" for the roofline
accounting (algorithmic bytes per kernel launch and per cell-update).
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
B_ALG_STEP = 416.0        # compulsory bytes per MHD cell-update
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


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
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------
# workload
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
GHOST = (3, 3, 3)


def workload_config(size, world, grid):
    return {"workload": f"Orszag-Tang vortex, {size}^3 cells per GPU "
                        f"(global {size * grid[0]}x{size * grid[1]}x{size * grid[2]}), "
                        "MHD VL+CT: PLM(theta=1.5) + HLLD + CT, periodic",
            "cells_per_gpu": size ** 3, "ghost_depth": 3,
            "riemann_solver": "hlld", "reconstruct_method": "plm",
            "courant": 0.3, "gamma": 5.0 / 3.0,
            "decomposition": f"{grid[0]}x{grid[1]}x{grid[2]} bricks",
            "step": "timestep + dt min-reduce + ghost refresh + compute (dt device-resident: vlct_timestep_dev / vlct_compute_dev; with a z split the z ghost exchange overlaps the interior update)",
            "l2_policy": "inputs (>=1.1 GB per field) exceed the 126 MB L2"}


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_sample(seconds_target=12.0, block=48, threads=None, steps_cap=200):
    """Time the reference CPU path (oracle/_ref when present, else the oracle
    port) on `threads` host threads, one `block`^3 brick of the Orszag-Tang
    workload per thread, for about `seconds_target` seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import oracle
    from enzo_e_b200 import abi, problems

    threads = threads or cpu_threads()
    kind = "ref" if oracle.have_ref() else "oracle"
    if kind == "oracle" and not oracle.have_oracle():
        oracle.build("oracle")
    # same parameters as PARAMS, set directly on the struct so that the CPU
    # arm never loads the CUDA library
    cfg = abi.default_config()
    cfg.mhd_choice = abi.MHD_CHOICE["constrained_transport"]
    cfg.riemann_solver = abi.RIEMANN["hlld"]
    cfg.reconstruct_method = abi.RECON["plm"]
    cfg.theta_limiter = 1.5
    cfg.courant = 0.3
    cfg.gamma = 5.0 / 3.0
    cfg.density_floor = cfg.pressure_floor = 1e-200
    n, g = (block, block, block), GHOST
    width = (1.0 / 512,) * 3
    workers = []
    for t in range(threads):
        lower = ((t % 8) * block * width[0], ((t // 8) % 8) * block * width[1],
                 (t // 64) * block * width[2])
        f = {k: v.numpy().copy() for k, v in problems.orszag_tang(
            n, g, lower, width, device="cpu").items()}
        blk = oracle.numpy_block(f, n, g, width)
        workers.append((oracle.CpuMethod(cfg, g, kind=kind), blk, f))

    def one_step(w):
        m, blk, _ = w
        dt = m.timestep(blk)
        oracle.refresh_periodic(blk, 0)
        m.compute(blk, dt)

    # calibrate with one step on one thread
    t0 = time.perf_counter()
    one_step(workers[0])
    per_step = time.perf_counter() - t0
    nsteps = max(2, min(steps_cap, int(seconds_target / max(per_step, 1e-6))))

    def loop(w):
        for _ in range(nsteps):
            one_step(w)

    ths = [threading.Thread(target=loop, args=(w,)) for w in workers]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    elapsed = time.perf_counter() - t0
    for m, _, _ in workers:
        m.close()
    cells = threads * block ** 3 * nsteps
    return {"value": cells / elapsed, "unit": UNIT, "cores": threads,
            "kind": "reference" if kind == "ref" else "port",
            "sample": f"{threads} threads x one {block}^3 brick of the "
                      f"Orszag-Tang workload x {nsteps} full steps "
                      f"(timestep+refresh+compute), {elapsed:.1f} s",
            "seconds": elapsed, "steps": nsteps}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    grid = _grid(args.gpus, args.layout)
    t_all = time.perf_counter()
    samples = []
    for _ in range(args.warmup):
        run_cpu_sample(seconds_target=1.0, block=args.cpu_block)
    per = max(2.0, min(20.0, 150.0 / max(1, args.steps)))
    for _ in range(args.steps):
        samples.append(run_cpu_sample(seconds_target=per, block=args.cpu_block))
    value = statistics.median(s["value"] for s in samples)
    cells_per_step = args.size ** 3 * args.gpus
    best = samples[0]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * cells_per_step / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.size, args.gpus, grid),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": best["cores"],
                             "kind": best["kind"], "sample": best["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "ms_per_step = time the CPU path would need for one step of "
                    "the full workload at the sampled per-cell rate",
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


def _grid(world, layout="slabs"):
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
    "k_update": {"doubles": 31, "units": lambda m, name: (m - 2) ** 3},
    # per cell: v, B (6) + 6 B-fluxes + 3 density fluxes in, 3 edge E out
    "k_edge_efield": {"doubles": 18, "units": lambda m, name: (m - 2) ** 3},
    # per cell: 3 edge E + 3 face B in, 3 face B out
    "k_face_bfield": {"doubles": 9, "units": lambda m, name: (m - 2) ** 3},
    # per cell: 8 fields in, pressure out
    "k_timestep": {"doubles": 9, "units": lambda m, name: m ** 3},
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from enzo_e_b200 import problems
    from enzo_e_b200.domain import Domain
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block

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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        opts = None
        try:   # NCCL's own stream above the compute stream: the ghost slabs
            # move while the interior update runs
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        except Exception:
            pass
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    domain = Domain(rank, world, grid=_grid(world, args.layout))
    grid = domain.grid
    size = args.size
    n_local = (size, size, size)
    width = tuple(1.0 / (size * grid[a]) for a in range(3))
    lower = domain.lower_corner(n_local, width)

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        fields = problems.orszag_tang(n_local, GHOST, lower, width, device=dev)
        method = EnzoMethodMHDVlct(PARAMS)
        block = Block(fields, n_local, GHOST, width)   # torch's current stream
        assert block.stream_is_current

        dt_dev = torch.empty(1, dtype=torch.float64, device=dev)

        def step(overlap=not args.no_overlap):
            # dt stays on the device (vlct_timestep_dev / vlct_compute_dev):
            # the cycles queue back to back, nothing waits for the host
            dt = method.timestep_dev(block, out=dt_dev)
            dt = domain.global_dt(dt, dev)
            # refresh + compute; with a z split the z exchange runs under the
            # interior part of the update (Domain.step)
            domain.step(method, block, dt, overlap=overlap)
            return dt

        def sync_all():
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize(dev)

        for _ in range(args.warmup):
            step()
        sync_all()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = method.kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            dt = step()
        e1.record(stream)
        sync_all()
        elapsed_ms = e0.elapsed_time(e1)
        launches = method.kernel_launches() - launches0
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        cells = size ** 3 * world * args.steps
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
            step(overlap=False)
        sync_all()
        report = method.profile_report()
        method.profile(False)

    hbm_gbs, peak_kind = measured_peaks()
    groups = dict(report)
    total_prof_ms = sum(ms for ms, _ in groups.values())
    m = size + 2 * GHOST[0]
    fam = family_rooflines(report, m, hbm_gbs)
    # dominant kernel = the family with the largest share of the step
    dom = max(fam, key=lambda k: fam[k]["ms"])
    d = fam[dom]
    traffic = None
    try:   # per-launch DRAM bytes from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom + "*" if dom == "k_flux" else dom,
                "achieved": d["achieved"], "peak": hbm_gbs,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": d["frac"],
                "traffic": traffic,
                "algorithmic_bytes_per_launch": d["bytes_per_launch"],
                "avg_launch_ms": d["avg_launch_ms"],
                "launches_per_step": d["launches"] / nprof,
                "share_of_step": d["ms"] / total_prof_ms if total_prof_ms else None,
                "note": "the flux kernels are bound by the FP64 pipe, not HBM "
                        "(ncu: profiles/), so their HBM fraction is low by "
                        "construction; see roofline_families for the "
                        "memory-bound kernels"}
    roofline_families = {k: {"achieved": v["achieved"], "frac": v["frac"],
                             "avg_launch_ms": v["avg_launch_ms"],
                             "share_of_step": v["ms"] / total_prof_ms}
                         for k, v in sorted(fam.items())}
    step_gbs = value * B_ALG_STEP / 1e9 / world
    roofline_step = {"bound": "hbm", "bytes_per_cell_update": B_ALG_STEP,
                     "achieved": step_gbs, "peak": hbm_gbs, "unit": "GB/s",
                     "frac": step_gbs / hbm_gbs,
                     "note": "per-GPU; 416 B = compulsory traffic of one MHD "
                             "cell-update (SURVEY 8d); the path is bound by "
                             "the FP64 pipe, see DESIGN.md"}
    kernels = {k: {"ms_per_step": ms / nprof, "launches_per_step": calls / nprof}
               for k, (ms, calls) in sorted(groups.items())}

    # ---- end-to-end through the C ABI with HOST (pinned) buffers ----------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, method, fields, n_local, width, world, dev)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = run_cpu_sample(seconds_target=args.cpu_seconds,
                                      block=args.cpu_block)
        cpu_baseline = {k: cpu_baseline[k] for k in
                        ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(size, world, grid),
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "roofline_families": roofline_families,
                "roofline_step": roofline_step,
                "cpu_baseline": cpu_baseline,
                "compute_only_ms": compute_ms,
                "compute_only_value": size ** 3 / (compute_ms * 1e-3),
                "kernels": kernels, "last_dt": float(dt.item()),
                "scratch_gb": method.scratch_bytes() / 1e9}
        print(json.dumps(line), flush=True)
    method.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, method, fields, n_local, width, world, dev):
    """Same metric through the C ABI with HOST buffers: every step copies the
    block's fields pinned-host -> device, runs timestep + compute, and copies
    the results back (all inside vlct_timestep / vlct_compute)."""
    import torch
    import torch.distributed as dist
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    from enzo_e_b200 import problems

    host = {}
    for k, v in fields.items():
        t = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        t.copy_(v)
        host[k] = t
    torch.cuda.synchronize(dev)
    host_np = {k: v.numpy() for k, v in host.items()}
    m2 = EnzoMethodMHDVlct(PARAMS)
    hb = Block(host_np, n_local, GHOST, width)
    steps = max(2, min(args.steps, 4))
    cells = n_local[0] * n_local[1] * n_local[2] * world * steps

    def timed(reuse):
        m2.set_option("host_mirror_reuse", reuse)
        for _ in range(1):
            dt = m2.timestep(hb)
            m2.compute(hb, dt)
        if world > 1:
            dist.barrier()
        h2d0, d2h0 = m2.staged_bytes()
        t0 = time.perf_counter()
        for _ in range(steps):
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
                "steps": steps, "ms_per_step": 1e3 * elapsed / steps}

    out = timed(0)
    out["path"] = ("vlct_timestep + vlct_compute with mem_space=HOST (pinned host "
                   "arrays; every call uploads its inputs and downloads its "
                   "outputs, as a z-pass pipeline of H2D / kernels / D2H)")
    # informational: the same loop with the caller's promise that nobody writes
    # the fields between compute and the next timestep (Enzo-E's cycle order),
    # so that timestep reuses the device copy compute left behind
    reuse = timed(1)
    reuse["path"] = "as e2e, option host_mirror_reuse = 1 (include/vlct.h)"
    out["with_host_mirror_reuse"] = reuse
    m2.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512,
                    help="cells per axis per GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--cpu-block", type=int, default=48)
    ap.add_argument("--layout", default="slabs", choices=["slabs", "bricks"],
                    help="N > 1: z slabs (1x1xN, default) or bricks (2x2x2 at N=8)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N > 1: exchange ghosts before the update instead of under it")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
