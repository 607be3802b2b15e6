"""torchrun worker: pipelined slab exchange at 512^3, status after every step"""
import os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch, torch.distributed as dist
from jtransforms_b200 import _lib
from jtransforms_b200.dist import SlabFFT3D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(os.environ.get("DBG_N", "512"))
f = SlabFFT3D(n, n, n, device_index=local)
a = torch.rand(f.local_elements(), dtype=torch.float64, device="cuda")
sync_each = os.environ.get("DBG_SYNC", "1") == "1"
for it in range(int(os.environ.get("DBG_STEPS", "6"))):
    t0 = time.perf_counter()
    f.forward(a)
    if sync_each:
        torch.cuda.synchronize()
        try:
            f.status()
            print("rank %d step %d ok %.1f ms" % (rank, it, (time.perf_counter() - t0) * 1e3), flush=True)
        except Exception as e:
            print("rank %d step %d FAILED %s %.1f ms" % (rank, it, e, (time.perf_counter() - t0) * 1e3), flush=True)
torch.cuda.synchronize()
import ctypes as C
lib = _lib.get()
lib.jtb_slab_profile(f._m, 1)
for rep in range(2):
    dist.barrier()
    torch.cuda.synchronize()
    f.forward(a)
    torch.cuda.synchronize()
    t3 = (C.c_float * 3)(); nb = C.c_int(); st = (C.c_float * 16)(); dn = (C.c_float * 16)()
    lib.jtb_slab_last_times(f._m, t3)
    lib.jtb_slab_chunk_times(f._m, C.byref(nb), st, dn)
    print("rank %d phases %s stored %s done %s" % (rank, ["%.3f" % v for v in t3], ["%.3f" % st[j] for j in range(nb.value)],
                                                     ["%.3f" % dn[j] for j in range(nb.value)]), flush=True)
lib.jtb_slab_profile(f._m, 0)
try:
    f.status()
    print("rank %d final ok" % rank, flush=True)
except Exception as e:
    print("rank %d final FAILED %s" % (rank, e), flush=True)
dist.barrier()
f.close()
dist.destroy_process_group()
