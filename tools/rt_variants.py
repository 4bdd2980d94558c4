"""Times ecckd_radiative_transfer_lw_dev at the bench shape for several ECCKD_RT knob settings
(run on the GPU box from the repo root)."""
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
res = {}
for name, env in (("plain", "1"), ("staged", None)):
    if env:
        os.environ["ECCKD_RT_PLAIN"] = env
    else:
        os.environ.pop("ECCKD_RT_PLAIN", None)
    for _ in range(3):
        E.radiative_transfer_lw_dev(ctx, ncol, nlay, ng, plh, od, emis, pls, dn, up)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        E.radiative_transfer_lw_dev(ctx, ncol, nlay, ng, plh, od, emis, pls, dn, up)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name] = (dn.clone(), up.clone())
    gbs = ncol * ng * 1760 / (ms * 1e-3) / 1e9
    print(f"{name}: {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  ({gbs / 6552.3:.3f} of peak)")
print("max |staged - plain|: dn %.3e up %.3e" % (float((res["staged"][0] - res["plain"][0]).abs().max()),
                                                   float((res["staged"][1] - res["plain"][1]).abs().max())))
# a wider spectral dimension (per-row copies): ng = 64
ng2 = 64
od2 = od.reshape(ncol // 2, nlay, ng2).contiguous()
pl2 = plh.reshape(ncol // 2, nlay + 1, ng2).contiguous()
ps2 = pls.reshape(ncol // 2, ng2).contiguous()
em2 = torch.ones(ng2, dtype=torch.float64, device=dev)
dn2, up2 = torch.empty_like(pl2), torch.empty_like(pl2)
for name, env in (("plain ng=64", "1"), ("staged ng=64", None)):
    if env:
        os.environ["ECCKD_RT_PLAIN"] = env
    else:
        os.environ.pop("ECCKD_RT_PLAIN", None)
    for _ in range(3):
        E.radiative_transfer_lw_dev(ctx, ncol // 2, nlay, ng2, pl2, od2, em2, ps2, dn2, up2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        E.radiative_transfer_lw_dev(ctx, ncol // 2, nlay, ng2, pl2, od2, em2, ps2, dn2, up2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = ncol * ng * 1760 / (ms * 1e-3) / 1e9
    print(f"{name}: {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  ({gbs / 6552.3:.3f} of peak)")

# shortwave (direct + Lambertian up): 24 B per column.layer.g algorithmic
ssi = torch.rand(ng, dtype=torch.float64, device=dev, generator=g).add_(0.5)
alb = torch.rand(ng, dtype=torch.float64, device=dev, generator=g).mul_(0.4)
import numpy as np
mu = np.full(ncol, 0.5)
for name, env in (("sw plain", "1"), ("sw staged", None)):
    if env:
        os.environ["ECCKD_RT_PLAIN"] = env
    else:
        os.environ.pop("ECCKD_RT_PLAIN", None)
    for _ in range(3):
        E.radiative_transfer_sw_dev(ctx, ncol, nlay, ng, mu, ssi, od, alb, dn, up)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        E.radiative_transfer_sw_dev(ctx, ncol, nlay, ng, mu, ssi, od, alb, dn, up)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = ncol * ng * (8 * nlay + 16 * (nlay + 1)) / (ms * 1e-3) / 1e9
    print(f"{name}: {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  ({gbs / 6552.3:.3f} of peak)")
