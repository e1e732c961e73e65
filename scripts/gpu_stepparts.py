"""Where does a step's time go? Times timestep / refresh / compute loops
separately and together (512^3 Orszag-Tang, device-resident)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import enzo_e_b200  # noqa
from enzo_e_b200 import problems
from enzo_e_b200.method import EnzoMethodMHDVlct, Block
from enzo_e_b200.domain import Domain
import bench

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    n = (size,) * 3
    width = (1.0 / size,) * 3
    fields = problems.orszag_tang(n, bench.GHOST, (0., 0., 0.), width, device=dev)
    m = EnzoMethodMHDVlct(bench.PARAMS)
    blk = Block(fields, n, bench.GHOST, width)
    dom = Domain()
    dt = torch.empty(1, dtype=torch.float64, device=dev)

    def timeit(label, fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"{label:28s} {e0.elapsed_time(e1) / reps:8.3f} ms/iter  (host enqueue {1e3 * (t1 - t0) / reps:.3f} ms)")

    m.timestep_dev(blk, out=dt)
    timeit("timestep_dev", lambda: m.timestep_dev(blk, out=dt))
    timeit("refresh", lambda: dom.refresh(m, blk))
    timeit("compute", lambda: m.compute(blk, dt))
    def step():
        m.timestep_dev(blk, out=dt); dom.refresh(m, blk); m.compute(blk, dt)
    timeit("step", step)
    def step2():
        d = m.timestep(blk); dom.refresh(m, blk); m.compute(blk, d)
    timeit("step (host dt)", step2)
    timeit("compute", lambda: m.compute(blk, dt))
