"""The Planck look-up table leg of the full build alone (python tools/planck_lut_probe.py): 231 temperatures over the
bench-sized g-point map, three calls, for a launch list under ncu."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import ecckd_b200 as E  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
n = 1 << 23
wn = 0.0002 + (3260.0 - 0.0002) / (n - 1) * torch.arange(n, device=dev, dtype=torch.float64)
dwn = torch.full_like(wn, float(wn[1] - wn[0]))
g, ng = bench.make_g_point(torch, wn, bench.smooth_key(torch, wn, seed=7), "ranked", seed=7)
plan = E.GPointPlan(ctx, g, n, ng)
t_lut = np.arange(120.0, 351.0, 1.0)
out = torch.empty((t_lut.size, ng), device=dev, dtype=torch.float64)
for it in range(3):
    ctx.synchronize()
    t0 = time.perf_counter()
    E.planck_lut_dev(ctx, plan, wn, dwn, t_lut, out)
    ctx.synchronize()
    print(f"planck_lut_dev call {it}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
print("checksum", float(out.sum()))
os.environ["ECCKD_PLANCK_LUT_GENERAL"] = "1"  # the reduction path with one group per temperature
ref = torch.empty_like(out)
for it in range(2):
    ctx.synchronize()
    t0 = time.perf_counter()
    E.planck_lut_dev(ctx, plan, wn, dwn, t_lut, ref)
    ctx.synchronize()
    print(f"general path call {it}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
rel = ((out - ref).abs() / ref.abs()).max().item()
print("largest relative difference between the two paths", rel)

