#!/bin/bash
# round-1 final artefacts: parity tests, smoke, bench (256^3, 512^3), reference
# arm, ncu launch list and ncu --set full of the flux kernels at 512^3
mkdir -p gpurun_out
TAG=${TAG:-r1i}
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py --size 256 --steps 5 --warmup 3 > gpurun_out/bench_256.json 2> gpurun_out/bench_256.err; echo "bench256 rc=$?"
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_512.json 2> gpurun_out/bench_512.err; echo "bench512 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_512.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
if [ "${NCU_FULL:-1}" = 1 ]; then
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_flux -s 24 -c 6 \
  -f -o gpurun_out/prof_flux512_$TAG python bench.py --size 512 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_flux512_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
python - <<'PY'
import json
for n in ("256", "512", "ref"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d.get("value"), d.get("ms_per_step"), d.get("clocks"), (d.get("roofline") or {}).get("frac"),
              (d.get("e2e") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value"))
        if "kernels" in d:
            print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
