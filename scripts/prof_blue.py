"""Batched Bluestein (config 3) on a small batch: target of ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtransforms_b200 as jt
n, b = 1000003, int(os.environ.get("BLUE_BATCH", "256"))
plan = jt.FloatFFT_1D(n)
a = torch.rand(2 * n * b, dtype=torch.float32, device="cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    plan.complexForwardBatch(a, b, 2 * n)
torch.cuda.synchronize()
