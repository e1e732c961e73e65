timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_multi_2_r1j.json 2> gpurun_out/bench_multi_2_r1j.err; echo rc=$?
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_multi_2_r1j.json").read().strip().splitlines()[-1])
print("N=2 value %.4g ms/step %.2f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"],3) for k,v in d["kernels"].items() if "slab" in k or "wrap" in k}, d.get("gpu_launches"))
PY
