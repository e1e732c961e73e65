#!/bin/bash
# multi-GPU: decomposition-invariance tests + bench at N GPUs (N = $NGPU), with
# and without the overlapped z exchange, slabs and bricks
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
run() { # name, extra flags
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e $2 > gpurun_out/bench_multi_${N}_$1.json 2> gpurun_out/bench_multi_${N}_$1.err
  echo "bench $N $1 rc=$?"; tail -2 gpurun_out/bench_multi_${N}_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_multi_${N}_$1.json").read().strip().splitlines()[-1])
    print("N=$N $1 value %.4g ms/step %.2f %s" % (d["value"], d["ms_per_step"], d["config"]["decomposition"]))
    print({k: round(v["ms_per_step"],3) for k,v in d["kernels"].items() if "slab" in k or "wrap" in k})
except Exception as e: print("failed", e)
PY
}
run slabs_overlap ""
run slabs_plain "--no-overlap"
if [ "$N" -ge 4 ]; then run bricks_overlap "--layout bricks"; fi
timeout 600 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_multi_1.json 2> gpurun_out/bench_multi_1.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_multi_1.json').read().strip().splitlines()[-1]); print('N=1 value %.4g ms/step %.2f' % (d['value'], d['ms_per_step']))"
