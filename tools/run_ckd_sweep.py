"""Timing of ecckd_run_ckd_dev (65536 columns x 54 layers x 32 g points, longwave): the layer-major path and the
column-major kernel it replaced (ECCKD_RUN_CKD_LEGACY=1).  Prints one line per setting."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402
from tests import synthetic as S  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    ctx = E.Context(0, st.cuda_stream)
    ncol, nlay, ng, nbase = 1 << 16, 54, 32, 64
    kind = sys.argv[1] if len(sys.argv) > 1 else "synthetic"
    c = S.ckd_model_case(ncol=nbase, nlay=nlay, ng=ng, seed=1234)
    if kind == "coherent":  # neighbouring columns alike in humidity too (what a model grid looks like)
        rng = np.random.default_rng(0)
        prof = np.exp(np.linspace(np.log(3e-6), np.log(2e-2), nlay))
        c["vmr_fl"][1] = prof[None, :] * np.exp(rng.normal(0, 0.15, (nbase, nlay)))
    model = E.CkdModel(ctx, c["gases"], c["log_pressure"], c["temperature"], c["temperature_planck"], c["planck_lut"])
    rep = ncol // nbase
    t = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, (rep, 1)))).to(dev)
    p_hl, t_hl = t(c["pressure_hl"]), t(c["temperature_hl"])
    vmr = [None if v is None else t(v) for v in c["vmr_fl"]]
    od = torch.empty((ncol, nlay, ng), dtype=torch.float64, device=dev)
    plh = torch.empty((ncol, nlay + 1, ng), dtype=torch.float64, device=dev)
    pls = torch.empty((ncol, ng), dtype=torch.float64, device=dev)
    dn, up = torch.empty_like(plh), torch.empty_like(plh)
    bdn = torch.empty((ncol, nlay + 1), dtype=torch.float64, device=dev)
    bup = torch.empty_like(bdn)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    full = lambda: model.run_ckd_dev(vmr, 1361.0, ncol, nlay, p_hl, t_hl, od, plh, pls, dn, up, bdn, bup)
    odp = lambda: model.run_ckd_dev(vmr, 1361.0, ncol, nlay, p_hl, t_hl, od, plh, pls)
    ref_up = None
    settings = [("1", "1"), ("0", "1"), ("0", "2"), ("0", "3")]  # column-major kernel; layer-major kernels 1, 2, 3 (the default)
    if len(sys.argv) > 2 and sys.argv[2] == "one":  # a single short run for ncu
        os.environ["ECCKD_RUN_CKD_LEGACY"] = "0"
        os.environ.pop("ECCKD_TOD_VARIANT", None)
        full()
        full()
        torch.cuda.synchronize()
        return
    for legacy, mb in settings:
        os.environ["ECCKD_RUN_CKD_LEGACY"] = legacy
        os.environ["ECCKD_TOD_VARIANT"] = mb
        ms = timed(full)
        ms_od = timed(odp)
        full()
        torch.cuda.synchronize()
        if ref_up is None:
            ref_up = up.clone()
        err = float((up - ref_up).abs().max())
        print(f"{kind} legacy={legacy} variant={mb}: fused {ms:.3f} ms ({ncol * ng / ms / 1e6:.2f} G col.g/s), "
              f"optical depth + Planck only {ms_od:.3f} ms, max |flux_up - legacy| {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
