"""torchrun worker: slab-decomposed 3-D FFT over the visible GPUs, both exchange modes, against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from jtransforms_b200.dist import SlabFFT3D
from oracle import jt_oracle as o

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mode in ("p2p", "nccl"):
    for (S, R, Cn) in [(64, 64, 64), (16, 512, 32), (8 * world, 64, 24), (2 * world, 512, 512), (128, 128, 128),
                       (256, 256, 32), (6 * world, 5 * world, 10)]:
        x = o.fill_uniform(2 * S * R * Cn, seed=11)
        want = o.complex_forward_3d(x, S, R, Cn).reshape(S, R, 2 * Cn)
        f = SlabFFT3D(S, R, Cn, device_index=local, exchange=mode)
        Ls, Rh = S // world, R // world
        for it in range(3):      # several steps: exercises the double buffering and the barrier epochs
            loc = torch.from_numpy(x.reshape(S, -1)[rank * Ls:(rank + 1) * Ls].copy().ravel()).cuda()
            res = f.forward(loc)
            torch.cuda.synchronize()
            got = res.cpu().numpy().reshape(S, Rh, 2 * Cn)
            err = o.rel_l2(got, want[:, rank * Rh:(rank + 1) * Rh])
            assert err < 1e-12 * 20, (mode, S, R, Cn, it, err)
            f.status()           # no device-side wait timed out
            if it == 0:          # distributed round trip (inverse of the k2-slabbed spectrum)
                back = f.inverse(res.clone(), True)
                torch.cuda.synchronize()
                ref = x.reshape(S, -1)[rank * Ls:(rank + 1) * Ls].ravel()
                err = o.rel_l2(back.cpu().numpy(), ref)
                assert err < 1e-12 * 20, ("inverse", mode, S, R, Cn, err)
        f.close()
dist.barrier()
dist.destroy_process_group()
print("rank %d ok" % rank)
