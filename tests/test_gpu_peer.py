"""Peer mode (csrc/peer.cu) on real GPUs: needs >= 2 devices, one process per GPU.  Runs
tools/dp_check.py under torchrun: after three minibatches with the gradient exchange fused into the
dW epilogue (reduce-scatter over NVLink peer memory, owner-sharded update, copy-engine all-gather)
every rank must hold the parameters of ONE rank stepping the full minibatch."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["peer-fused", "peer-ce", "peer-kernel", "nccl"])
def test_data_parallel_matches_single_gpu(mode):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    # peer mode once per gradient transport: NVLink stores from the dW epilogue / local staging + copy engines /
    # local staging + the co-resident copy kernel
    env = dict(os.environ, VBNN_DP=mode.split("-")[0])
    if mode.startswith("peer-"):
        env["VBNN_PEER_TRANSPORT"] = {"peer-fused": "1", "peer-ce": "2", "peer-kernel": "3"}[mode]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP CHECK OK" in r.stdout
