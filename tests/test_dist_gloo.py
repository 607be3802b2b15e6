"""world_size-2 gloo test of the slab-decomposed 3-D FFT host logic (jtransforms_b200/dist.py).
Kernels run through the g++-emulated build of the library sources (tests/emu); the exchange is a real
torch.distributed all-to-all between two processes."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
from jtransforms_b200 import _lib
_lib.use(%(emu)r)
from jtransforms_b200.dist import SlabFFT3D
from oracle import jt_oracle as o
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
for (S, R, Cn) in [(4, 4, 8), (8, 2, 4), (6, 10, 5)]:
    x = o.fill_uniform(2 * S * R * Cn, seed=7)
    want = o.complex_forward_3d(x, S, R, Cn).reshape(S, R, 2 * Cn)
    f = SlabFFT3D(S, R, Cn)
    Ls = S // 2
    loc = torch.from_numpy(x.reshape(S, -1)[rank * Ls:(rank + 1) * Ls].copy().ravel())
    res = f.forward(loc)
    host = torch.zeros(2 * S * R * Cn, dtype=torch.float64)
    f.scatter_to_host(res, host)
    got = host.numpy().reshape(S, R, 2 * Cn)
    Rh = R // 2
    mine = slice(rank * Rh, (rank + 1) * Rh)
    err = o.rel_l2(got[:, mine], want[:, mine])
    assert err < 1e-12 * 12, err
    other = slice((1 - rank) * Rh, (2 - rank) * Rh)
    assert not got[:, other].any()
    # distributed round trip: inverse of the k2-slabbed spectrum gives back this rank's slices
    for scale in (True, False):
        back = f.inverse(res.clone(), scale)
        ref = x.reshape(S, -1)[rank * Ls:(rank + 1) * Ls].ravel() * (1.0 if scale else S * R * Cn)
        err = o.rel_l2(back.numpy(), ref)
        assert err < 1e-12 * 12, ("inverse", scale, err)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_slab_fft3d_two_ranks(tmp_path):
    emu = os.path.join(HERE, "emu", "_build", "libjtb200_emu.so")
    subprocess.run(["sh", os.path.join(HERE, "emu", "build_emu.sh")], check=True, capture_output=True)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "emu": emu, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    for p, out in zip(procs, outs):
        assert p.returncode == 0, out
