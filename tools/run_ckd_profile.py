"""Runs the run_ckd LW core at the bench shape a few times (profiling target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ecckd_b200 as E  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
print(bench.bench_flux(torch, E, ctx, dev, 6552.3)["fused_run_ckd"])
