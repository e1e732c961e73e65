#!/usr/bin/env python
"""Throughput of many small blocks (Enzo-E's usual operating point): one
launch sequence per block (vlct_timestep + vlct_compute in a loop) against
the batched entry points (vlct_timestep_batch + vlct_compute_batch), device
resident, MHD PLM + HLLD + CT. usage: bench_batch.py [block_size nblocks]..."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_host(size, nb, steps=3):
    """the same batch with HOST (pinned) blocks: what a drop-in into Cello's
    host-resident fields sees. Compares one-shot staging (host_batch_blocks =
    nb) with the double-buffered sub-batch pipeline (auto)."""
    import torch
    from bench import PARAMS, GHOST
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    per_axis = round(nb ** (1 / 3))
    width = (1.0 / (size * per_axis),) * 3
    n = (size,) * 3
    blocks, keep = [], []
    for b in range(nb):
        c = (b % per_axis, (b // per_axis) % per_axis, b // per_axis ** 2)
        lower = tuple(c[a] * size * width[a] for a in range(3))
        f = problems.orszag_tang(n, GHOST, lower, width, device="cpu")
        pinned = {}
        for k, v in f.items():
            t = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            t.copy_(v)
            pinned[k] = t
        keep.append(pinned)
        blocks.append(Block({k: v.numpy() for k, v in pinned.items()}, n, GHOST, width))
    method = EnzoMethodMHDVlct(PARAMS)
    out = {"block": size, "nblocks": nb, "cells": nb * size ** 3, "mem_space": "HOST"}
    for name, sub, mode in (("one_shot", nb, 0), ("pipelined", 0, 0),
                            ("pipelined_zero_copy_kernels", 0, 1),
                            ("pipelined_copy_per_block_field", 0, 2)):
        method.set_option("host_batch_blocks", sub)
        method.set_option("host_batch_copy_mode", mode)
        dt = method.timestep_batch(blocks)
        method.compute_batch(blocks, dt)
        h0 = method.staged_bytes()
        t0 = time.perf_counter()
        for _ in range(steps):
            dt = method.timestep_batch(blocks)
            method.compute_batch(blocks, dt)
        el = (time.perf_counter() - t0) / steps
        h1 = method.staged_bytes()
        out[name] = {"ms_per_step": 1e3 * el, "cell_updates_per_s": nb * size ** 3 / el,
                     "h2d_gb_per_step": (h1[0] - h0[0]) / steps / 1e9,
                     "d2h_gb_per_step": (h1[1] - h0[1]) / steps / 1e9}
    out["speedup"] = out["one_shot"]["ms_per_step"] / out["pipelined"]["ms_per_step"]
    method.close()
    return out


def run(size, nb, steps=3):
    import torch
    from bench import PARAMS, GHOST
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    dev = torch.device("cuda", 0)
    per_axis = round(nb ** (1 / 3))
    width = (1.0 / (size * per_axis),) * 3
    n = (size,) * 3
    blocks, keep = [], []
    for b in range(nb):
        c = (b % per_axis, (b // per_axis) % per_axis, b // per_axis ** 2)
        lower = tuple(c[a] * size * width[a] for a in range(3))
        f = problems.orszag_tang(n, GHOST, lower, width, device=dev)
        keep.append(f)
        blocks.append(Block(f, n, GHOST, width))
    method = EnzoMethodMHDVlct(PARAMS)
    out = {"block": size, "nblocks": nb, "cells": nb * size ** 3}

    def loop_step():
        dt = min(method.timestep(b) for b in blocks)
        for b in blocks:
            method.compute(b, dt)

    def batch_step():
        dt = method.timestep_batch(blocks)
        method.compute_batch(blocks, dt)

    for name, step in (("batch", batch_step), ("loop", loop_step)):
        step()
        torch.cuda.synchronize()
        l0 = method.kernel_launches()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / steps
        out[name] = {"ms_per_step": 1e3 * el,
                     "cell_updates_per_s": nb * size ** 3 / el,
                     "launches_per_step": (method.kernel_launches() - l0) / steps}
    out["speedup"] = out["loop"]["ms_per_step"] / out["batch"]["ms_per_step"]
    method.close()
    return out


if __name__ == "__main__":
    host = "--host" in sys.argv
    args = [int(a) for a in sys.argv[1:] if a != "--host"] or [32, 512, 16, 4096]
    for size, nb in zip(args[::2], args[1::2]):
        print(json.dumps((run_host if host else run)(size, nb)), flush=True)
