#!/bin/bash
# multi-GPU evidence: decomposition invariance (bricks / slabs / overlapped
# exchange; z-extruded vortex and 3-D turbulence) with the log kept for
# profiles/, then the bench line at N GPUs (bricks, all extras)
mkdir -p gpurun_out
N=${NGPU:-2}
TAG=${TAG:-r2}
nvidia-smi -L | wc -l
rm -f gpurun_out/${TAG}_multi_gpu_invariance_$N.log
VLCT_MULTI_GPU_LOG=gpurun_out/${TAG}_multi_gpu_invariance_$N.log timeout 1800 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "${KEXPR:-test_bricks}" 2>&1 | tail -4 | tee gpurun_out/${TAG}_multi_gpu_pytest_$N.log
cat gpurun_out/${TAG}_multi_gpu_invariance_$N.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 $BENCH_FLAGS > gpurun_out/${TAG}_bench_multi_$N.json 2> gpurun_out/${TAG}_bench_multi_$N.err
echo "bench $N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_multi_$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms/step %.2f %s clocks %s" % (d["value"], d["ms_per_step"], d["config"]["decomposition"], d["clocks"]))
    e = d.get("e2e") or {}
    print(" e2e", e.get("value"), e.get("ms_per_step"), (e.get("two_calls") or {}).get("ms_per_step"))
    for k, v in (d.get("workloads") or {}).items():
        print(" extra", k, v.get("value"), v.get("ms_per_step"), v.get("roofline_frac"), v.get("error"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/${TAG}_bench_multi_$N.err").read()[-3000:])
PY
