#!/bin/bash
# multi-GPU: decomposition-invariance test + bench at N GPUs (N = $NGPU)
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
for n in ${BENCH_N:-$N}; do
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_multi_$n.json 2> gpurun_out/bench_multi_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_multi_$n.json 2> gpurun_out/bench_multi_$n.err
  fi
  echo "bench $n rc=$?"; tail -3 gpurun_out/bench_multi_$n.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_multi_$n.json").read().strip().splitlines()[-1])
    print("N=$n value %.4g ms/step %.2f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value")))
    print({k: round(v["ms_per_step"],3) for k,v in d["kernels"].items()})
except Exception as e: print("failed", e)
PY
done
