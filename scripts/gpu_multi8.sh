#!/bin/bash
# 8 GPUs (short: charged 8x): decomposition invariance of the overlapped step, bench slabs / bricks
mkdir -p gpurun_out
N=${NGPU:-8}
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "${N}-overlap" 2>&1 | tail -3
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e $2 > gpurun_out/bench_multi_${N}_$1.json 2> gpurun_out/bench_multi_${N}_$1.err
  echo "bench $N $1 rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_multi_${N}_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 value %.4g ms/step %.2f %s clocks %s" % (d["value"], d["ms_per_step"], d["config"]["decomposition"], d["clocks"]))
except Exception as e: print("failed", e); print(open("gpurun_out/bench_multi_${N}_$1.err").read()[-1500:])
PY
}
run slabs_overlap ""
run bricks_overlap "--layout bricks"
run slabs_plain "--no-overlap"
