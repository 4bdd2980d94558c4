"""Where does the host-buffer entry point spend its time?  (diagnostic)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ecckd_b200 as E
import bench

log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 23
nlay, nwav = 53, 1 << log2
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
dnu = (3260.0 - 0.0002) / (nwav - 1)
wn = 0.0002 + dnu * torch.arange(nwav, device=dev, dtype=torch.float64)
dwn = torch.full((nwav,), dnu, dtype=torch.float64, device=dev)
key = bench.smooth_key(torch, wn, seed=7)
g_point, ng = bench.make_g_point(torch, wn, key, "ranked", seed=7)
p_hl = bench.pressure_grid(nlay); p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
t_hl = bench.temperature_profile(p_hl); t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
od = bench.make_spectrum(torch, wn, nlay, p_hl, 1001, 1e-3)
h_od = torch.empty((nlay, nwav), dtype=torch.float64, pin_memory=True); h_od.copy_(od)
torch.cuda.synchronize()
def T(f, n=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("H2D 53 x nwav pinned: %.1f ms" % T(lambda: od.copy_(h_od, non_blocking=True)))
print("plan create: %.1f ms" % T(lambda: E.GPointPlan(ctx, g_point, nwav, ng).close()))
plan = E.GPointPlan(ctx, g_point, nwav, ng)
out = torch.empty((nlay, ng), dtype=torch.float64, device=dev); mn = torch.empty_like(out); mx = torch.empty_like(out)
def devcall(n):
    E.average_optical_depth_to_g_point_dev(ctx, plan, 1e-3, p_fl[:n], p_hl[:n + 1], od, nwav, E.WEIGHT_PLANCK, None, nwav,
                                           t_fl[:n], wn, dwn, "transmission", out, mn, mx, ng)
for n in (53, 27, 7, 1):
    print("dev call, %2d layers x 1 spectrum: %.2f ms" % (n, T(lambda: devcall(n))))
h_wn, h_dwn, h_g = wn.cpu().numpy(), dwn.cpu().numpy(), g_point.cpu().numpy()
res = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32))
print("host call: %.1f ms" % T(lambda: E.planck_average_optical_depth_to_g_point(ctx, ng, 1e-3, p_fl, p_hl, h_g, h_od.numpy(), t_fl, h_wn, h_dwn, "transmission", out=res)))
