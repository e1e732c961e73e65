"""GPU, N > 1: block-decomposition invariance of the NCCL ghost exchange.
Skipped on a single-GPU box (the gloo test covers the host logic there)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("problem", ["ot", "turbulence"])
@pytest.mark.parametrize("mode", ["plain", "overlap", "overlap-slabs"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_bricks_match_single_block(world, mode, problem):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world),
           os.path.join(HERE, "multi_gpu_worker.py"), "3", mode, problem]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:]
    # keep the evidence: scripts/gpu_multi*.sh copy this log to profiles/
    log = os.environ.get("VLCT_MULTI_GPU_LOG")
    if log:
        with open(log, "a") as fh:
            fh.write([l for l in out.stdout.splitlines() if "MULTI_GPU_" in l][-1] + "\n")
