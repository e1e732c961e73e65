#!/usr/bin/env python
"""Experiment (NOT a product path): what would FMA contraction buy, and what
would it cost in parity? Run with VLCT_B200_LIB pointing at a library whose
kernels were compiled with -fmad=true; compares one and ten steps of a seeded
random MHD state (PLM + HLLD + CT) with the CPU oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from helpers import make_config, random_state, copy_state, oracle
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    cfg = make_config(riemann="hlld", recon="plm", theta=1.5, mhd=True)
    n, g, d = (48, 32, 24), (3, 3, 3), (1 / 48, 1 / 48, 1 / 48)
    host = random_state(cfg, n, g, seed=7)
    out = {"lib": os.environ.get("VLCT_B200_LIB", "default")}
    for nsteps in (1, 10):
        ref = copy_state(host)
        cpu = oracle.CpuMethod(cfg, g)
        blk = oracle.numpy_block(ref, n, g, d)
        dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
        m = EnzoMethodMHDVlct(config=cfg)
        b = Block(dev, n, g, d)
        for _ in range(nsteps):
            dt = cpu.timestep(blk)
            oracle.refresh_periodic(blk, 0)
            cpu.compute(blk, dt)
            m.refresh_periodic(b, 7)
            m.compute(b, dt)       # same dt: isolates the update's arithmetic
        m.synchronize()
        got = {k: v.cpu().numpy() for k, v in dev.items()}
        m.close()
        cpu.close()
        act = (slice(3, -3),) * 3
        res = {}
        for k in ref:
            if k == "pressure":
                continue
            a, r = got[k][act], ref[k][act]
            scale = np.max(np.abs(r))
            res[k] = {"max_abs_over_field_scale": float(np.max(np.abs(a - r)) / scale),
                      "max_cellwise_rel": float(np.max(np.abs(a - r) / np.maximum(np.abs(r), 1e-300))),
                      "bit_identical": bool(np.array_equal(a, r))}
        out[f"steps_{nsteps}"] = res
    print(json.dumps(out))


if __name__ == "__main__":
    main()
