#!/bin/bash
# first GPU contact: parity tests, smoke, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py --size 256 --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "bench256 rc=$?"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench512 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_512.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_256.json; cat gpurun_out/bench_512.json; tail -3 gpurun_out/bench_512.err
