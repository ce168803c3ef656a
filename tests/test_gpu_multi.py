"""Multi-GPU parity: the sharded path (NCCL exchange of factor rows + all-reduce of partial Grams)
must reproduce the oracle exactly like the single-GPU path.  Needs >= 2 GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("peer_store", ["1", "0", "pc"])
@pytest.mark.parametrize("world", [2, 4])
def test_sharded_training_matches_oracle(world, peer_store):
    """peer_store=1: finished rows are stored into every replica by the sweep kernels (CUDA IPC over
    NVLink); 0: NCCL broadcasts after the sweep.  Both must land on the oracle's factors."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, EALS_PEER_STORE="0" if peer_store == "0" else "1",
               EALS_PEER_PRED_CACHE="1" if peer_store == "pc" else "0", EALS_CHECK_REPLICAS="1")
    peer_store, port = ("1", 20) if peer_store == "pc" else (peer_store, 10 * int(peer_store))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world + port),
           os.path.join(ROOT, "tests", "dist_parity_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "dist parity ok" in res.stdout
    assert f"peer_store {bool(int(peer_store))}" in res.stdout
