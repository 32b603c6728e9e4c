"""Multi-GPU tests of the fused step (2 ranks, NCCL + NVLink peer memory).  Need >= 2 visible GPUs: run with
``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``; skipped on a single-GPU box.  The host-side logic
of the same routes (sharding, the gloo all-reduce) is covered on CPU in tests/test_cabi_host.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif("_gpus() < 2", reason="needs two GPUs")
def test_two_rank_fused_step_p2p_nccl_and_timeout():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_multi_gpu_worker.py")]
    env = dict(os.environ)
    env.pop("VS_NCCL", None)
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0 and "multi-gpu worker ok" in p.stdout, p.stdout[-4000:]
