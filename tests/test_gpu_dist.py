"""Multi-GPU tests (need >= 2 GPUs; skipped otherwise) and single-GPU virtual-rank tests of the fused exchange."""
import os
import subprocess
import sys

import pytest

import parity_cases as pc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("P,dims", [(8, (64, 512, 64)), (2, (16, 512, 128)), (4, (32, 64, 40)), (1, (8, 1024, 16))])
def test_slab_scatter_virtual_ranks(P, dims):
    from jtransforms_b200 import _lib
    _lib._lib = None
    pc.slab_scatter_virtual(_lib.get(), "Double", dims, P, torch_device="cuda:0")


def test_slab_two_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(HERE, "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
