"""Device-resident averaging of ONE spectrum (53 layers x 2^23): the configuration the host-buffer call reduces in."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
nlay, nwav = 53, 1 << 23
dnu = (3260.0 - 0.0002) / (nwav - 1)
wn = 0.0002 + dnu * torch.arange(nwav, device=dev, dtype=torch.float64)
dwn = torch.full((nwav,), dnu, dtype=torch.float64, device=dev)
key = bench.smooth_key(torch, wn, seed=7)
g_point, ng = bench.make_g_point(torch, wn, key, "ranked", seed=7)
p_hl = bench.pressure_grid(nlay)
p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
t_hl = bench.temperature_profile(p_hl)
t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
od = bench.make_spectrum(torch, wn, nlay, p_hl, 1001, 4.15e-4)
plan = E.GPointPlan(ctx, g_point, nwav, ng)
out, mn, mx = (torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(3))
warn = torch.zeros(ng, dtype=torch.int32, device=dev)
for nl in (53, 12, 1):
    def run():
        E.average_optical_depth_to_g_point_dev(ctx, plan, 4.15e-4, p_fl[:nl], p_hl[:nl + 1], od, nwav, E.WEIGHT_PLANCK, None, 0,
                                               t_fl[:nl], wn, dwn, "transmission", out, mn, mx, ng, warn)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{nl} layers x 1 spectrum: {ms:.3f} ms  ({ms / nl:.3f} ms per layer, {nl * nwav * 8 / ms / 1e6:.0f} GB/s)")
