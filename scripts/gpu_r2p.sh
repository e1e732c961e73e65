#!/bin/bash
# round 2: full GPU suite, smoke, bench (all extras), reference arm, ncu launch
# list of the bench command, ncu --set full of the TMA-staged edge kernel
mkdir -p gpurun_out
TAG=${TAG:-r2p}
timeout 1500 python -m pytest tests -m gpu -q --durations=5 --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -12 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$TAG.log; tail -2 gpurun_out/smoke_$TAG.log
if [ "${BENCH:-1}" = 1 ]; then
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
fi
if [ "${REF:-0}" = 1 ]; then
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref rc=$?"
fi
if [ "${NCU_LIST:-1}" = 1 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
fi
if [ "${NCU_FULL:-1}" = 1 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_ct_tma} -s 2 -c 2 -o gpurun_out/prof_${NCU_KERNEL:-k_ct_tma}_$TAG -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_${NCU_KERNEL:-k_ct_tma}_$TAG.ncu-rep --page details --csv > gpurun_out/ncu_full_${NCU_KERNEL:-k_ct_tma}_$TAG.csv 2>/dev/null
fi
python - <<PY
import json
for n in ("bench_$TAG", "bench_ref_$TAG"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), d.get("clocks"), (d.get("roofline") or {}).get("frac"), (d.get("roofline_fp64") or {}).get("frac"))
        e = d.get("e2e") or {}
        print(" e2e", e.get("value"), e.get("ms_per_step"), (e.get("two_calls") or {}).get("ms_per_step"))
        print(" cpu", d.get("cpu_baseline"))
        for k, v in (d.get("workloads") or {}).items():
            print(" extra", k, v.get("value"), v.get("ms_per_step"), v.get("roofline_frac"), v.get("error"))
        if "kernels" in d:
            print({k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
        print(d.get("roofline_families"))
    except Exception as e:
        print(n, "failed", e)
        try: print(open(f"gpurun_out/{n}.err").read()[-2000:])
        except Exception: pass
PY
