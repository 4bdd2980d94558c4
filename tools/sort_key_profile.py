"""Runs the longwave sort-key kernel and the band ranking at the config-D shape (profiling target):
[54][2^22] optical depths, 13 bands."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(0, st.cuda_stream)
nlay, nwav = 54, 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 22
p_hl = np.concatenate([[1.0], np.logspace(0.2, 5.0, nlay)])
wn = torch.linspace(0.0002, 3260.0, nwav, dtype=torch.float64, device=dev)
dwn = torch.full_like(wn, 3260.0 / nwav)
od = torch.rand(nlay, nwav, dtype=torch.float64, device=dev) * torch.exp(4.0 * torch.sin(wn * 0.37) - 3.0)[None, :]
key = torch.empty(nwav, dtype=torch.float64, device=dev)
rank = torch.empty(nwav, dtype=torch.int32, device=dev)
edges = [0.0, 350.0, 500.0, 630.0, 700.0, 820.0, 980.0, 1080.0, 1180.0, 1390.0, 1480.0, 1800.0, 2080.0, 3260.0]
for _ in range(3):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    E.sort_key_lw_dev(ctx, nlay, nwav, p_hl, wn, dwn, od, nwav, 0.5, key)
    e1.record()
    E.rank_bands_dev(ctx, nwav, wn, key, np.array(edges[:-1]), np.array(edges[1:]), None, rank)
    e2.record()
    torch.cuda.synchronize()
    print("sort_key_lw %.3f ms, rank_bands %.3f ms" % (e0.elapsed_time(e1), e1.elapsed_time(e2)))
