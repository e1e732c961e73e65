#!/usr/bin/env python
"""Board power and SM clock while ONE kernel family of the step runs in a loop
(512^3 Orszag-Tang): is the step's energy, not its critical path, the floor?
Uses the handle option "debug_kernel_mask" (fields end up as garbage).
Output: one JSON line per family."""
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Sampler:
    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        self.rows = []
        self.proc = subprocess.Popen(
            ["nvidia-smi", "--query-gpu=clocks.sm,power.draw.instant,power.draw,clocks.mem",
             "--format=csv,noheader,nounits", "-lms", "50", "-i", "0"],
            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            try:
                self.rows.append([float(x) for x in line.strip().split(",")])
            except ValueError:
                pass

    def stop(self):
        self.proc.terminate()
        r = self.rows[len(self.rows) // 3:]     # steady part
        if not r:
            return {}
        med = lambda i: sorted(x[i] for x in r)[len(r) // 2]
        return {"sm_mhz": med(0), "power_instant_w": med(1), "power_avg_w": med(2),
                "samples": len(r)}


def main():
    import torch
    from bench import PARAMS, GHOST
    from enzo_e_b200 import problems
    from enzo_e_b200.method import EnzoMethodMHDVlct, Block
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    dev = torch.device("cuda", 0)
    n, width = (size,) * 3, (1.0 / size,) * 3
    stream = torch.cuda.Stream(device=dev)
    K = {"scal": 1, "flux_xy": 2, "flux_z": 4, "edge": 8, "face": 16, "update": 32}
    # ("ct": edge + face together take the fused TMA-staged kernel, k_ct_tma)
    families = {"all": 63, "flux": 6, "cell": 56, "ct": 24, "edge": 8, "face": 16, "update": 32,
                "flux_xy": 2, "flux_z": 4}
    with torch.cuda.stream(stream):
        fields = problems.orszag_tang(n, GHOST, (0, 0, 0), width, device=dev)
        method = EnzoMethodMHDVlct(PARAMS)
        block = Block(fields, n, GHOST, width)
        dt = torch.full((1,), 1e-5, dtype=torch.float64, device=dev)
        for name, mask in families.items():
            if only is not None and name not in only:
                continue
            saved = {k: v.clone() for k, v in fields.items()} if name == "all" else None
            method.set_option("debug_kernel_mask", mask)
            for _ in range(3):
                method.compute(block, dt)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            method.compute(block, dt)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            one = e0.elapsed_time(e1)
            reps = max(5, int(seconds * 1e3 / max(one, 0.1)))
            s = Sampler()
            s.start()
            e0.record(stream)
            for _ in range(reps):
                method.compute(block, dt)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / reps
            out = s.stop()
            out.update({"family": name, "ms_per_step": ms, "first_ms": one, "reps": reps,
                        "joule_per_step": out.get("power_instant_w", 0) * ms * 1e-3})
            print(json.dumps(out), flush=True)
    method.close()


if __name__ == "__main__":
    main()
