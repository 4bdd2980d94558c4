"""Where the host-buffer call spends its time: the call at several layer counts against the raw pinned copy."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
log2 = 23
nwav = 1 << log2
dnu = (3260.0 - 0.0002) / (nwav - 1)
wn = 0.0002 + dnu * torch.arange(nwav, device=dev, dtype=torch.float64)
dwn = torch.full((nwav,), dnu, dtype=torch.float64, device=dev)
key = bench.smooth_key(torch, wn, seed=7)
g_point, ng = bench.make_g_point(torch, wn, key, "ranked", seed=7)
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)  # noqa: E731
h_wn, h_dwn, h_g = pin(wn).numpy(), pin(dwn).numpy(), pin(g_point).numpy()
for nlay in (53, 26, 8, 1):
    p_hl = bench.pressure_grid(nlay)
    p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
    t_hl = bench.temperature_profile(p_hl)
    t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
    od = bench.make_spectrum(torch, wn, nlay, p_hl, 1001, 4.15e-4)
    h = pin(od)
    d = torch.empty_like(od)
    res = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32))
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    raw = (time.time() - t0) / 3
    hn = h.numpy()
    for _ in range(2):
        E.planck_average_optical_depth_to_g_point(ctx, ng, 4.15e-4, p_fl, p_hl, h_g, hn, t_fl, h_wn, h_dwn, "transmission", out=res)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        E.planck_average_optical_depth_to_g_point(ctx, ng, 4.15e-4, p_fl, p_hl, h_g, hn, t_fl, h_wn, h_dwn, "transmission", out=res)
    torch.cuda.synchronize()
    call = (time.time() - t0) / 3
    print(f"nlay {nlay:2d}: raw od copy {raw * 1e3:6.2f} ms, call {call * 1e3:6.2f} ms, difference {(call - raw) * 1e3:5.2f} ms")
    del od, h, d
