"""Runs the plan-API 512^3 transform (the bench's device-resident path) a few times: target of the ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jtransforms_b200 as jt
prec = os.environ.get("PROF_PREC", "f64")
n = int(os.environ.get("N3D", "512"))
dt = torch.float64 if prec == "f64" else torch.float32
plan = (jt.DoubleFFT_3D if prec == "f64" else jt.FloatFFT_3D)(n, n, n)
a = torch.rand(2 * n ** 3, dtype=dt, device="cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    plan.complexForward(a)
torch.cuda.synchronize()
