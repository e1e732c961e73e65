#!/usr/bin/env python
"""Informational throughput of the other BASELINE.json configurations on one
B200 (bench.py measures configs[2]a, the headline): device-resident cycles
(timestep + refresh + compute), CUDA-event timing, per-kernel breakdown.

  sod256      configs[1]a  3-D Sod, hydro, PLM(1.5) + HLLC + dual energy, 256^3,
                           outflow along x, periodic across
  sedov256    configs[1]b  Sedov-like blast, same solver, 256^3, periodic
  blast512    configs[2]b  MHD blast, PLM(1.5) + HLLD + CT, 512^3
  ot512s4     configs[4]   Orszag-Tang 512^3 + 4 passive scalars
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HYDRO = {"Method:mhd_vlct:mhd_choice": "no_bfield",
         "Method:mhd_vlct:riemann_solver": "hllc",
         "Method:mhd_vlct:reconstruct_method": "plm",
         "Method:mhd_vlct:theta_limiter": 1.5,
         "Method:mhd_vlct:courant": 0.3,
         "Physics:fluid_props:dual_energy:type": "modern",
         "Physics:fluid_props:dual_energy:eta": 1e-3,
         "Physics:fluid_props:floors:density": 1e-200,
         "Physics:fluid_props:floors:pressure": 1e-200}


def run(name, steps=10, warmup=3):
    import torch
    from bench import PARAMS, GHOST
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    dev = torch.device("cuda", 0)
    passive, outflow_x = (), False
    if name == "sod256":
        size, params = 256, dict(HYDRO, **{"Physics:fluid_props:eos:gamma": 1.4})
        make = lambda n, w: problems.hydro_sod(n, GHOST, (0, 0, 0), w, device=dev, gamma=1.4)
        outflow_x, b_alg = True, 240.0
    elif name == "sedov256":
        size, params = 256, dict(HYDRO, **{"Physics:fluid_props:eos:gamma": 5.0 / 3.0})
        make = lambda n, w: problems.hydro_blast(n, GHOST, (0, 0, 0), w, device=dev)
        b_alg = 240.0
    elif name == "blast512":
        size, params = 512, PARAMS
        make = lambda n, w: problems.mhd_blast(n, GHOST, (0, 0, 0), w, device=dev)
        b_alg = 416.0
    elif name == "ot512s4":
        size, params = 512, PARAMS
        passive = tuple(f"passive_{k}" for k in range(4))
        make = lambda n, w: problems.orszag_tang(n, GHOST, (0, 0, 0), w, device=dev, n_passive=4)
        b_alg = 416.0 + 4 * 40.0
    else:
        raise SystemExit(f"unknown config {name}")
    n, width = (size,) * 3, (1.0 / size,) * 3
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        fields = make(n, width)
        method = EnzoMethodMHDVlct(params, n_passive=len(passive))
        block = Block(fields, n, GHOST, width, passive=passive)
        dt_dev = torch.empty(1, dtype=torch.float64, device=dev)

        def step():
            dt = method.timestep_dev(block, out=dt_dev)
            if outflow_x:
                method.refresh_periodic(block, 6)
                method.boundary(block, 0, 0, "outflow")
                method.boundary(block, 0, 1, "outflow")
            else:
                method.refresh_periodic(block, 7)
            method.compute(block, dt)

        for _ in range(warmup):
            step()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / steps
        method.profile(True)
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        rep = method.profile_report()
        method.profile(False)
        finite = all(bool(torch.isfinite(v).all()) for v in fields.values())
    value = size ** 3 / (ms * 1e-3)
    out = {"config": name, "cells": size ** 3, "ms_per_step": ms,
           "cell_updates_per_s": value, "bytes_per_cell_update": b_alg,
           "hbm_gbs_algorithmic": value * b_alg / 1e9, "all_finite": finite,
           "dt": float(dt_dev.item()),
           "kernels_ms_per_step": {k: round(v[0] / 3, 3) for k, v in sorted(rep.items())}}
    method.close()
    return out


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["sod256", "sedov256", "blast512", "ot512s4"]):
        print(json.dumps(run(name)), flush=True)
