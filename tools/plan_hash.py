"""SHA-256 of the exported g-point plans of a few maps and the time to build the bench-sized one
(python tools/plan_hash.py): used to check that a change of the plan kernels leaves the plans bit-identical."""
import hashlib
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
for log2n, seed in ((18, 3), (20, 7), (23, 7)):
    n = 1 << log2n
    wn = 0.0002 + (3260.0 - 0.0002) / (n - 1) * torch.arange(n, device=dev, dtype=torch.float64)
    g, ng = bench.make_g_point(torch, wn, bench.smooth_key(torch, wn, seed=seed), "ranked", seed=seed)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan = E.GPointPlan(ctx, g, n, ng)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    info, rounds = plan.export()
    h = hashlib.sha256(rounds.tobytes()).hexdigest()
    print(f"2^{log2n} wavenumbers, {ng} g points: {info["nrounds"]} rounds, plan built in {dt * 1e3:.1f} ms, sha256 {h[:16]}")
    plan.close()
