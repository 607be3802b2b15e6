"""torchrun probe: concurrent pinned H2D / D2H bandwidth per rank, with and without NUMA affinity."""
import os, sys, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
aff0 = len(os.sched_getaffinity(0))
mode = os.environ.get("AFF", "1")
err = ""
if mode == "1":
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception as e:
        err = repr(e)
aff1 = sorted(os.sched_getaffinity(0))
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def bw(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9
h2d = bw(lambda: d.copy_(h, non_blocking=True))
d2h = bw(lambda: h.copy_(d, non_blocking=True))
s1 = torch.cuda.Stream()
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s1):
        h2.copy_(d2, non_blocking=True)
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
bi = bw(both)
out = {"rank": rank, "aff_before": aff0, "aff_after": "%d cpus [%d..%d]" % (len(aff1), aff1[0], aff1[-1]), "err": err,
       "h2d_GBps": round(h2d, 1), "d2h_GBps": round(d2h, 1), "bidir_each_GBps": round(bi, 1)}
outs = [None] * world
dist.all_gather_object(outs, out)
if rank == 0:
    for o in outs: print(json.dumps(o))
dist.destroy_process_group()
