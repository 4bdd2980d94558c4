"""Throughput of the rows that sit either side of the headline path, at the sizes of SURVEY.md 8d
(configs D and E): the batched candidate-interval error of find_g_points, the forward cost of
optimize_lut, and the per-band ranking.  Prints one JSON object (run on the GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
PEAK = 6552.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
out = {"peak_gbs": PEAK}
gen = torch.Generator(device=dev).manual_seed(5)


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- interval error (LW, "transmission"): one band of 2^22 ranked wavenumbers x 54 layers, 16 candidate intervals
nlay, n = 54, 1 << 22
r = lambda *s: torch.rand(s, dtype=torch.float64, device=dev, generator=gen)  # noqa: E731
bg = r(nlay, n).mul_(2.0)
metric = r(nlay, n).mul_(0.9)
planck = r(nlay + 1, n).add_(1.0)
hr = r(nlay, n).mul_(1e-5)
emis, surf, fds, fut = r(n).mul_(0.1).add_(0.9), r(n).add_(2.0), r(n), r(n)
p_hl = torch.linspace(1.0, 1.0e5, nlay + 1, dtype=torch.float64, device=dev)
lw = torch.full((nlay,), 1.0 / nlay, dtype=torch.float64, device=dev)
nint = 16
cuts = np.linspace(0, n, nint + 1).astype(np.int64)
i1, i2 = cuts[:-1], cuts[1:] - 1
err = torch.empty(nint, dtype=torch.float64, device=dev)
ms = timed(lambda: E.interval_error_batch_lw_dev(ctx, "transmission", nlay, n, n, p_hl, planck, emis, surf, bg, metric,
                                                 fds, fut, hr, 0.2, lw, i1, i2, err))
# algorithmic bytes: fit pass reads metric, planck, hr (24 B per point.layer); flux pass reads bg, planck (16 B)
alg = nlay * n * 40.0
out["interval_error_lw"] = {"ms": ms, "points_x_layers_per_s": nlay * n / (ms * 1e-3), "intervals": nint,
                            "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / PEAK,
                            "config": f"{nint} intervals partitioning a band of 2^22 wavenumbers x {nlay} layers, transmission"}
del bg, metric, planck, hr

# ---- forward cost (config E): 100 scenes x 50 profiles (LW, ng = 64, 13 bands)
nprof, ng, nband = 50, 64, 13
bm = torch.tensor(np.sort(np.arange(ng) % nband), dtype=torch.int32, device=dev)
scene = dict(p=torch.linspace(1.0, 1.0e5, nlay + 1, dtype=torch.float64, device=dev).repeat(nprof, 1).contiguous(),
             planck=r(nprof, nlay + 1, ng).add_(1.0), emis=r(nprof, nband).mul_(0.1).add_(0.9), surf=r(nprof, ng).add_(2.0),
             od=r(nprof, nlay, ng), dn=r(nprof, nlay + 1, nband), up=r(nprof, nlay + 1, nband), hr=r(nprof, nlay, nband))
cp, ct, pt = (torch.empty(k, dtype=torch.float64, device=dev) for k in (nprof, 1, 1))
L = E._lib.lib()


def cost_scene():
    s = scene
    E.check(L.ecckd_cost_function_ckd_lw_dev(ctx.handle, nprof, nlay, ng, nband, s["p"].data_ptr(), s["planck"].data_ptr(),
                                             s["emis"].data_ptr(), s["surf"].data_ptr(), s["od"].data_ptr(),
                                             s["dn"].data_ptr(), s["up"].data_ptr(), s["hr"].data_ptr(), None, None, 0.2, 0.2,
                                             0.8, 0.0, None, 0.5, 1.0e4, None, None, bm.data_ptr(), cp.data_ptr(),
                                             ct.data_ptr(), pt.data_ptr()))


def cost_all():
    for _ in range(100):
        cost_scene()


ms = timed(cost_all, n=3, warm=1)
out["forward_cost_lw"] = {"ms_per_100_scenes": ms, "columns_x_g_per_s": 100 * nprof * ng / (ms * 1e-3),
                          "config": "optimize_lut forward cost, 100 scenes x 50 profiles x 54 layers x 64 g points, 13 bands; "
                                    "one launch pair per scene (latency-bound: 50 CTAs per launch)"}
# the same 100 scenes through the batched entry point (one launch for all scenes)
scenes100 = [dict(nprof=nprof, pressure_hl=scene["p"], optical_depth=scene["od"], flux_dn=scene["dn"], flux_up=scene["up"],
                  hr=scene["hr"], planck_hl=scene["planck"], surf_emissivity=scene["emis"], surf_planck=scene["surf"])
             for _ in range(100)]
cs100 = torch.empty(100, dtype=torch.float64, device=dev)
ms = timed(lambda: E.cost_function_ckd_scenes_dev(ctx, False, scenes100, nlay, ng, nband, 0.2, 0.2, 0.8, 0.0,
                                                  d_band_mapping=bm, d_cost_scene=cs100, d_cost_total=ct), n=5, warm=2)
out["forward_cost_lw_scenes"] = {"ms_per_100_scenes": ms, "columns_x_g_per_s": 100 * nprof * ng / (ms * 1e-3),
                                 "config": "the same 100 scenes x 50 profiles through ecckd_cost_function_ckd_scenes_dev "
                                           "(one launch of 5000 CTAs + the per-scene sums)"}
# one big scene of 5000 profiles (what a column-sharded rank would batch)
nprof = 5000
scene = dict(p=scene["p"][:1].repeat(nprof, 1).contiguous(), planck=r(nprof, nlay + 1, ng).add_(1.0),
             emis=r(nprof, nband).mul_(0.1).add_(0.9), surf=r(nprof, ng).add_(2.0), od=r(nprof, nlay, ng),
             dn=r(nprof, nlay + 1, nband), up=r(nprof, nlay + 1, nband), hr=r(nprof, nlay, nband))
cp = torch.empty(nprof, dtype=torch.float64, device=dev)
ms = timed(cost_scene, n=10, warm=2)
out["forward_cost_lw_batched"] = {"ms_per_5000_profiles": ms, "columns_x_g_per_s": nprof * ng / (ms * 1e-3),
                                  "config": "the same 5000 profiles as one batch"}

# ---- the reverse mode of the same 5000 profiles: dJ/dtau, then its transpose of the look-up-table
# interpolation into the tables of an 8-gas model
grad_od = torch.empty((nprof, nlay, ng), dtype=torch.float64, device=dev)


def grad_scene():
    s = scene
    E.check(L.ecckd_cost_function_ckd_lw_grad_dev(
        ctx.handle, nprof, nlay, ng, nband, s["p"].data_ptr(), s["planck"].data_ptr(), s["emis"].data_ptr(),
        s["surf"].data_ptr(), s["od"].data_ptr(), s["dn"].data_ptr(), s["up"].data_ptr(), s["hr"].data_ptr(), None, None,
        0.2, 0.2, 0.8, 0.0, None, 0.5, 1.0e4, None, None, bm.data_ptr(), grad_od.data_ptr()))


ms = timed(grad_scene, n=10, warm=2)
out["gradient_cost_lw_batched"] = {"ms_per_5000_profiles": ms, "columns_x_g_per_s": nprof * ng / (ms * 1e-3),
                                   "config": "ecckd_cost_function_ckd_lw_grad_dev on the same 5000 profiles x 54 layers x 64 g points"}
try:
    from tests import synthetic as S
    ck = S.ckd_model_case(ncol=64, nlay=nlay, ng=ng, seed=1234)
    model = E.CkdModel(ctx, ck["gases"], ck["log_pressure"], ck["temperature"], ck["temperature_planck"], ck["planck_lut"])
    rep = (nprof + 63) // 64
    tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep, 1))[:nprof])).to(dev)
    p_hl, t_hl = tile(ck["pressure_hl"]), tile(ck["temperature_hl"])
    vmr = [None if v is None else tile(v) for v in ck["vmr_fl"]]
    grads = [torch.zeros(np.asarray(g["molar_abs"]).shape, dtype=torch.float64, device=dev) for g in ck["gases"]]
    ms = timed(lambda: model.calc_total_optical_depth_adjoint_dev(vmr, nprof, nlay, p_hl, t_hl, grad_od, grads), n=5, warm=2)
    out["gradient_lut_adjoint"] = {"ms_per_5000_profiles": ms,
                                   "config": "ecckd_calc_total_optical_depth_adjoint_dev: dJ/dtau of 5000 profiles x 54 layers x 64 g "
                                             "points gathered into the tables of 8 gases (one with 12 concentrations), fixed order"}
    model.close()
except Exception as exc:  # noqa: BLE001
    out["gradient_lut_adjoint"] = {"error": repr(exc)}

# ---- ranking (config D): 2^23 keys, 13 bands
nw = 1 << 23
wn = torch.linspace(0.0002, 3260.0, nw, dtype=torch.float64, device=dev)
key = torch.randn(nw, dtype=torch.float64, device=dev, generator=gen)
edges = [0.0, 350.0, 500.0, 630.0, 700.0, 820.0, 980.0, 1080.0, 1180.0, 1390.0, 1480.0, 1800.0, 2080.0, 3260.0]
rank = torch.empty(nw, dtype=torch.int32, device=dev)
ib = torch.empty(nw, dtype=torch.int32, device=dev)
ms = timed(lambda: E.rank_bands_dev(ctx, nw, wn, key, np.array(edges[:-1]), np.array(edges[1:]), ib, rank), n=5, warm=2)
t0 = time.time()
ko = key.cpu().numpy()
order = np.argsort(ko[: nw // 8], kind="stable")
cpu_s = (time.time() - t0) * 8
out["rank_bands"] = {"ms": ms, "keys_per_s": nw / (ms * 1e-3), "numpy_stable_argsort_keys_per_s_1_core": nw / cpu_s,
                     "config": "2^23 keys, 13 narrow bands, stable per-band rank"}
print(json.dumps(out))
