"""Times the host-buffer averaging call with float64 and float32 spectra and the raw pinned H2D copies."""
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
nlay, log2 = 53, int(sys.argv[1]) if len(sys.argv) > 1 else 23
nwav = 1 << log2
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
pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)  # noqa: E731
h64 = pin(od)
h32 = torch.empty(od.shape, dtype=torch.float32, pin_memory=True)
h32.copy_(od)
h_wn, h_dwn, h_g = pin(wn).numpy(), pin(dwn).numpy(), pin(g_point).numpy()
res = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32))
d64 = torch.empty_like(od)
d32 = torch.empty(od.shape, dtype=torch.float32, device=dev)
for name, h, d in (("f64", h64, d64), ("f32", h32, d32)):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    print(f"raw pinned H2D {name}: {dt * 1e3:.1f} ms  {h.numel() * h.element_size() / dt / 1e9:.1f} GB/s")
for name, h in (("f64", h64.numpy()), ("f32", h32.numpy()), ("f64", h64.numpy()), ("f32", h32.numpy())):
    for _ in range(2):
        E.planck_average_optical_depth_to_g_point(ctx, ng, 4.15e-4, p_fl, p_hl, h_g, h, t_fl, h_wn, h_dwn, "transmission", out=res)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(3):
        E.planck_average_optical_depth_to_g_point(ctx, ng, 4.15e-4, p_fl, p_hl, h_g, h, t_fl, h_wn, h_dwn, "transmission", out=res)
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 3
    print(f"host call {name}: {dt * 1e3:.1f} ms  checksum {res[0].sum():.6e}")
