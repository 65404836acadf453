"""Gradient exchange over NVLink peer memory (csrc/peer.cu, pcoe.dp.PeerExchange) on real GPUs.  The workers run under
torch.distributed.run in a subprocess (one process per GPU): world 1 on any box (the kernel talks to itself through
the same flags), world 2 where two GPUs are visible."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world: int, script: str, env_extra=None, port: int = 29631):
    env = dict(os.environ, OMP_NUM_THREADS="1", **(env_extra or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script)]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("world", [1, 2])
def test_peer_allreduce_matches_nccl_exactly(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = _run(world, "tests/peer_worker.py", port=29631 + world)
    assert r.returncode == 0 and "peer_worker: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
def test_training_gradients_overlap_vs_single_allreduce_world2(exchange):
    """the overlapped two-bucket exchange (through NCCL or through the peer kernel), with and without
    gradient accumulation under no_sync(), == one torch.distributed all-reduce after backward; every rank holds the
    same reduced buffer."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run(2, "tools/dp_overlap_check.py", {"PCOE_EXCHANGE": exchange}, port=29641 + (exchange == "peer"))
    assert r.returncode == 0 and "dp_overlap_check: PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
