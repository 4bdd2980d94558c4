"""Runs ecckd_radiative_transfer_lw_dev at the bench shape a few times (profiling target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
ncol, nlay, ng = 65536, 54, 32
g = torch.Generator(device=dev).manual_seed(1)
od = torch.rand((ncol, nlay, ng), dtype=torch.float64, device=dev, generator=g).mul_(3.0)
plh = torch.rand((ncol, nlay + 1, ng), dtype=torch.float64, device=dev, generator=g).add_(1.0)
pls = torch.rand((ncol, ng), dtype=torch.float64, device=dev, generator=g).add_(2.0)
emis = torch.ones(ng, dtype=torch.float64, device=dev)
dn = torch.empty((ncol, nlay + 1, ng), dtype=torch.float64, device=dev)
up = torch.empty_like(dn)
for _ in range(5):
    E.radiative_transfer_lw_dev(ctx, ncol, nlay, ng, plh, od, emis, pls, dn, up)
torch.cuda.synchronize()
print("ok", float(dn[0, -1, 0]))
