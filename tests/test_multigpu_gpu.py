"""Multi-GPU parity: one process per GPU (torchrun), row bands, peer-halo and NCCL-exchange modes vs the unsharded run.
Skipped on boxes with a single GPU (the sharding logic itself is covered on CPU with gloo in test_sharding_cpu.py and on
one GPU by test_denoiser_gpu.py::test_peer_halo_mode / test_record_halo_exchange_mode)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_band_sharding_across_gpus(tmp_path, world):
    if _gpu_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert open(tmp_path / "verdict.txt").read() == "ok", p.stdout[-3000:]
