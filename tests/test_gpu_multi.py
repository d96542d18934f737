"""The default multi-GPU product path over REAL peers: P2PRowShardedTrainer (hole_shard_step: rows
gathered from the owners' shards and deltas staged at the owners over NVLink / CUDA IPC) and the
candidate-sharded ranking, one process per GPU, against a single-GPU engine on the same global
batches (tools/multi_gpu_check.py).  Needs >= 2 GPUs; on a 1-GPU box the virtual-rank tests of
tests/test_gpu_train.py cover the same kernels with local "peers"."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_p2p_row_sharded_training_and_ranking_match_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from graphembeddings_b200 import build
    build.build()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "MULTI_GPU_CHECK OK" in r.stdout, tail
    assert '"trainer": "P2PRowShardedTrainer"' in r.stdout, tail
