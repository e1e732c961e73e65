N=${NGPU:-4}
for G in ${GRIDS:-2,2,1 2,1,2}; do
echo "=== grid $G"
VLCT_TEST_GRID=$G timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu_worker.py 3 plain turbulence 2>&1 | grep "MULTI\|MISMATCH\|dt seq\|Error" | tail -30
done
