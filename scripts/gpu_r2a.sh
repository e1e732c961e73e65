#!/bin/bash
# round 2, first call: the whole GPU test-suite (new: active floors, at-size
# parity, fused compute+timestep), smoke, the reworked bench line (+ extras),
# the reference arm, and per-kernel fp64 instruction / DRAM byte counts of one
# step for the Orszag-Tang and the turbulence workloads
mkdir -p gpurun_out
TAG=${TAG:-r2a}
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 2400 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$TAG.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
for W in ot turbulence; do
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:^k_ -s 57 -c 19 --csv --log-file gpurun_out/step_metrics_${W}_$TAG.csv \
  python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_step_${W}_$TAG.log 2>&1; echo "ncu $W rc=$?"
done
tail -5 gpurun_out/pytest_gpu_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log
python - <<PY
import json
for n in ("bench_$TAG", "bench_ref_$TAG"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), d.get("clocks"), (d.get("roofline") or {}).get("frac"))
        e = d.get("e2e") or {}
        print(" e2e", e.get("value"), e.get("ms_per_step"), (e.get("two_calls") or {}).get("ms_per_step"))
        print(" cpu", d.get("cpu_baseline"))
        for k, v in (d.get("workloads") or {}).items():
            print(" extra", k, v.get("value"), v.get("ms_per_step"), v.get("roofline_frac"), v.get("error"))
        if "kernels" in d:
            print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
        try: print(open(f"gpurun_out/{n}.err").read()[-2000:])
        except Exception: pass
PY
