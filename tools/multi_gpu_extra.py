"""Configs D and E of SURVEY.md 8d/8e on N GPUs of one box (one process per GPU, NCCL):

  D  find_g_points ranking: every rank computes the longwave sort keys of its wavenumber range
     (ecckd_sort_key_lw_dev on the column slice), ONE all-gather of the keys (8 B per wavenumber),
     then the stable per-band rank replicated on every rank (bit-exact, no exchange inside the sort).
  E  optimize_lut forward cost: the profiles of every scene are sharded over the ranks, every rank
     evaluates its columns, ONE all-reduce of the {cost, penalty} scalars of all scenes.

Rank 0 also evaluates both unsharded and checks the sharded results against them.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/multi_gpu_extra.py            (or plain `python tools/multi_gpu_extra.py` for N = 1)
Prints one JSON object on rank 0.  Timing: CUDA events on the work stream, max over ranks."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402
from ecckd_b200 import sharding  # noqa: E402

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    sys.stdout.flush()  # the "NCCL version ..." banner of the first collective goes to stderr, not into the JSON
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.all_reduce(torch.zeros(1, device=dev))
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(st)
ctx = E.Context(local_rank, st.cuda_stream)
L = E._lib.lib()
out = {"n_gpus": world}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


# ------------------------------------------------------------------ config D
nlay, nwav = 54, 1 << 23
p_hl = np.concatenate([[1.0], np.logspace(0.2, 5.0, nlay)])  # Pa, top to surface
edges = [0.0, 350.0, 500.0, 630.0, 700.0, 820.0, 980.0, 1080.0, 1180.0, 1390.0, 1480.0, 1800.0, 2080.0, 3260.0]
b1, b2 = np.array(edges[:-1]), np.array(edges[1:])


def spectrum(lo, hi):
    """Synthetic optical depths [nlay][hi-lo]: an element-wise function of the global wavenumber index."""
    w = torch.arange(lo, hi, dtype=torch.float64, device=dev)
    wn = 0.0002 + w * ((3260.0 - 0.0002) / (nwav - 1))
    dwn = torch.full_like(wn, (3260.0 - 0.0002) / (nwav - 1))
    shape = torch.exp(4.0 * torch.sin(wn * 0.37) + 2.0 * torch.sin(wn * 7.1 + 1.0) - 3.0)
    col = torch.tensor(np.diff(p_hl) / 1.0e5 * 40.0, dtype=torch.float64, device=dev)
    lay = torch.arange(nlay, dtype=torch.float64, device=dev)
    od = shape[None, :] * col[:, None] * (1.0 + 0.3 * torch.sin(wn[None, :] * 0.011 + lay[:, None]))
    return wn, dwn, od.contiguous()


lo, hi = sharding.shard_range(nwav, rank, world)
wn_l, dwn_l, od_l = spectrum(lo, hi)
wn_full = 0.0002 + torch.arange(nwav, dtype=torch.float64, device=dev) * ((3260.0 - 0.0002) / (nwav - 1))
key_l = torch.empty(hi - lo, dtype=torch.float64, device=dev)
ib = torch.empty(nwav, dtype=torch.int32, device=dev)
rk = torch.empty(nwav, dtype=torch.int32, device=dev)
state = {}


def key_fn(a, b):
    E.sort_key_lw_dev(ctx, nlay, b - a, p_hl, wn_l, dwn_l, od_l, b - a, 0.5, key_l)
    return key_l


def config_d():
    state["key"] = sharding.sharded_sort_keys(key_fn, nwav, rank, world)
    E.rank_bands_dev(ctx, nwav, wn_full, state["key"], b1, b2, ib, rk)


ms_d = timed(config_d)
ms_key = timed(lambda: key_fn(lo, hi))
ms_rank = timed(lambda: E.rank_bands_dev(ctx, nwav, wn_full, state["key"], b1, b2, ib, rk))
check_d = None
if rank == 0:
    if world > 1:
        del od_l
        wn_f, dwn_f, od_f = spectrum(0, nwav)
        key_f = torch.empty(nwav, dtype=torch.float64, device=dev)
        E.sort_key_lw_dev(ctx, nlay, nwav, p_hl, wn_f, dwn_f, od_f, nwav, 0.5, key_f)
        rk_f = torch.empty(nwav, dtype=torch.int32, device=dev)
        E.rank_bands_dev(ctx, nwav, wn_full, key_f, b1, b2, None, rk_f)
        torch.cuda.synchronize()
        check_d = {"keys_bit_equal": bool(torch.equal(key_f.view(torch.int64), state["key"].view(torch.int64))),
                   "rank_equal": bool(torch.equal(rk_f, rk))}
        del od_f, key_f, rk_f
    r = rk.cpu().numpy()
    check_d = dict(check_d or {}, rank_is_permutation=bool(np.array_equal(np.sort(r), np.arange(nwav))))
# key kernel: 8 B per point.layer in + 8 B per point out (+ 16 B wavenumber grid); all-gather 8 B per point
out["config_d_ranking"] = {
    "ms": ms_d, "ms_sort_key_shard": ms_key, "ms_rank_bands_replicated": ms_rank, "wavenumbers_per_s": nwav / (ms_d * 1e-3),
    "sort_key_GBps_per_gpu": ((hi - lo) * (8.0 * nlay + 24.0)) / (ms_key * 1e-3) / 1e9,
    "allgather_bytes": 8 * nwav if world > 1 else 0, "check": check_d,
    "config": f"sort_key_lw on [54][2^23 / {world}] per rank -> all-gather keys -> stable rank in 13 bands on every rank"}
barrier()

# ------------------------------------------------------------------ config E
nscene, nprof, ng, nband = 100, 50, 64, 13
gen = torch.Generator(device=dev).manual_seed(11)  # the same scenes on every rank
r = lambda *s: torch.rand(s, dtype=torch.float64, device=dev, generator=gen)  # noqa: E731
bm = torch.tensor(np.sort(np.arange(ng) % nband), dtype=torch.int32, device=dev)
sc = dict(p=torch.linspace(1.0, 1.0e5, nlay + 1, dtype=torch.float64, device=dev).repeat(nscene, nprof, 1).contiguous(),
          planck=r(nscene, nprof, nlay + 1, ng).add_(1.0), emis=r(nscene, nprof, nband).mul_(0.1).add_(0.9),
          surf=r(nscene, nprof, ng).add_(2.0), od=r(nscene, nprof, nlay, ng).sub_(0.001),
          dn=r(nscene, nprof, nlay + 1, nband), up=r(nscene, nprof, nlay + 1, nband), hr=r(nscene, nprof, nlay, nband))
cp = torch.empty(nprof, dtype=torch.float64, device=dev)
acc = torch.zeros(2, nscene, dtype=torch.float64, device=dev)  # {cost_total, penalty} per scene


def scene_table(a, b):
    """Profiles [a, b) of every scene as the scene list of ecckd_cost_function_ckd_scenes_dev."""
    return [dict(nprof=b - a, pressure_hl=sc["p"][s, a:b], optical_depth=sc["od"][s, a:b], flux_dn=sc["dn"][s, a:b],
                 flux_up=sc["up"][s, a:b], hr=sc["hr"][s, a:b], planck_hl=sc["planck"][s, a:b],
                 surf_emissivity=sc["emis"][s, a:b], surf_planck=sc["surf"][s, a:b]) for s in range(nscene)]


def forward(table, acc_t, reduce):
    """All scenes in one launch; acc_t[:, 0] = scene costs (penalty included), acc_t[:, 1] = penalties."""
    acc_t.zero_()
    if table:
        cs, pn = acc_t[0], acc_t[1]
        E.cost_function_ckd_scenes_dev(ctx, False, table, nlay, ng, nband, 0.2, 0.2, 0.8, 0.0, d_band_mapping=bm,
                                       d_cost_scene=cs, d_penalty_scene=pn)
    if reduce and world > 1:
        dist.all_reduce(acc_t, op=dist.ReduceOp.SUM)


c0, c1 = sharding.shard_columns(nprof, rank, world)
mine = scene_table(c0, c1) if c1 > c0 else []
ms_e = timed(lambda: forward(mine, acc, True), n=5, warm=2)
check_e = None
if rank == 0:
    full = torch.zeros_like(acc)
    sharded = acc.clone()
    forward(scene_table(0, nprof), full, False)
    torch.cuda.synchronize()
    rel = ((sharded - full).abs() / full.abs()).max().item()
    check_e = {"max_rel_diff_vs_unsharded": rel, "ok": bool(rel < 1e-12)}
out["config_e_forward_cost"] = {
    "ms_per_100_scenes": ms_e, "columns_x_g_per_s": nscene * nprof * ng / (ms_e * 1e-3),
    "allreduce_bytes": int(acc.numel() * 8) if world > 1 else 0, "check": check_e,
    "config": f"100 scenes x 50 profiles x 54 layers x 64 g points, 13 bands; profiles {c0}-{c1 - 1} of every scene on rank 0, "
              "all scenes in one launch per rank (ecckd_cost_function_ckd_scenes_dev), one all-reduce of the 100 x {cost, penalty} scalars"}
barrier()
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
