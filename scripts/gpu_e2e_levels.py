#!/usr/bin/env python
"""e2e (vlct_compute_and_timestep, pinned HOST block, 512^3) against the depth
of the z-pass staging pipeline (option host_pipeline_levels)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import PARAMS, GHOST
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    size = 512
    dev = torch.device("cuda", 0)
    n, width = (size,) * 3, (1.0 / size,) * 3
    fields = problems.orszag_tang(n, GHOST, (0, 0, 0), width, device=dev)
    host = {}
    for k, v in fields.items():
        t = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        t.copy_(v)
        host[k] = t
    del fields
    torch.cuda.synchronize()
    host_np = {k: v.numpy() for k, v in host.items()}
    for levels in [int(a) for a in sys.argv[1:]] or [-1, 6, 8, 12, 24, 34]:
        m = EnzoMethodMHDVlct(PARAMS)
        m.set_option("host_pipeline_levels", levels)
        hb = Block(host_np, n, GHOST, width)
        dt = m.timestep(hb)
        dt = m.compute_and_timestep(hb, dt)
        t0 = time.perf_counter()
        steps = 3
        for _ in range(steps):
            dt = m.compute_and_timestep(hb, dt)
        el = (time.perf_counter() - t0) / steps
        print(json.dumps({"host_pipeline_levels": levels, "ms_per_cycle": 1e3 * el,
                          "launches": m.kernel_launches()}), flush=True)
        m.close()


if __name__ == "__main__":
    main()
