"""Two-rank NCCL run of train_bpe_sharded (one process per GPU).  Needs >= 2 GPUs: skipped otherwise (the host
logic is covered on CPU by tests/test_sharded_cpu.py, the per-GPU C ABI by tests/test_gpu_train.py)."""
import os
import pathlib
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_train_bpe_sharded_nccl(world, tmp_path):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tests" / "helpers" / "nccl_train_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        assert (tmp_path / ("ok.%d" % rank)).exists()
