"""Real-NCCL check of the N-independent sharded reduction: run under torchrun with any number of ranks;
every rank reduces its run of the SAME six wavenumber segments, the slots are exchanged with
ecckd_average_optical_depth_exchange_dev over NCCL and combined in segment order.  Rank 0 prints the SHA-256 of
the finalized look-up-table slices - the digest must be the same for every world size (and is compared with
the oracle at 1e-12)."""
import hashlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ecckd_b200 as E  # noqa: E402
from ecckd_b200 import sharding  # noqa: E402
from tests import synthetic as S  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(st)
    ctx = E.Context(local, st.cuda_stream)
    nseg = 6
    c = S.lw_spectrum_case(nlay=9, nwav=300007, ng=22, nband=4, seed=41)
    method = sys.argv[1] if len(sys.argv) > 1 else "transmission"
    nlay, ng, nwav = c["nlay"], c["ng"], c["nwav"]
    spans = sharding.segment_ranges(nwav, nseg)
    sb = sharding.segment_owners(nseg, world)
    mine = range(sb[rank], sb[rank + 1])
    n3, n2 = nlay * ng * 3, nlay * ng * 2
    seg_sums = torch.zeros((nseg, n3), dtype=torch.float64, device=dev)
    seg_minmax = torch.zeros((nseg, n2), dtype=torch.float64, device=dev)
    seg_count = torch.zeros((nseg, ng), dtype=torch.int64, device=dev)
    comm = E.Comm(ctx, rank, world, exchange=lambda box: dist.broadcast_object_list(box, src=0)) if world > 1 else None
    keep = []
    for s in mine:  # this rank uploads ONLY its own segments
        a, b = spans[s]
        t = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(dev, dt)
        g = t(c["g_point"][a:b], torch.int32)
        wn, dwn = t(c["wavenumber"][a:b], torch.float64), t(c["d_wavenumber"][a:b], torch.float64)
        od = t(c["od"][:, a:b], torch.float64)
        plan = E.GPointPlan(ctx, g, b - a, ng)
        seg_count[s].copy_(torch.from_numpy(plan.counts()))
        E.average_optical_depth_partial_dev(ctx, plan, nlay, c["pressure_fl"], [od], b - a, E.WEIGHT_PLANCK, None, 0,
                                            c["temperature_fl"], wn, dwn, method, seg_sums[s], seg_minmax[s])
        keep += [g, wn, dwn, od, plan]
    sums = torch.empty(n3, dtype=torch.float64, device=dev)
    minmax = torch.empty(n2, dtype=torch.float64, device=dev)
    count = torch.empty(ng, dtype=torch.int64, device=dev)
    E.average_optical_depth_exchange_dev(ctx, comm, sb, seg_sums, seg_minmax, sums, minmax, seg_count, count)
    out = [torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(3)]
    E.average_optical_depth_finalize_dev(ctx, ng, count, [c["vmr"]], c["pressure_fl"], c["pressure_hl"], method, sums, minmax,
                                         [out[0]], [out[1]], [out[2]], ng)
    torch.cuda.synchronize()
    res = [o.cpu().numpy() for o in out]
    digest = hashlib.sha256(b"".join(r.tobytes() for r in res)).hexdigest()
    digests = [None] * world
    if world > 1:
        dist.all_gather_object(digests, digest)
    else:
        digests = [digest]
    if rank == 0:
        from tests.oracle_lib import oracle
        O = oracle()
        planck = O.planck_function(c["temperature_fl"], c["wavenumber"], c["d_wavenumber"])
        ref = O.average_optical_depth_to_g_point(ng, c["vmr"], c["pressure_fl"], c["pressure_hl"], c["g_point"], c["od"],
                                                 planck, method)
        S.assert_avg_close(tuple(res) + (None,), ref, method=method, scale=S.out_scale(c))
        assert len(set(digests)) == 1, "ranks disagree"
        print(f"world={world} method={method} sha256={digest} (all {world} ranks agree; oracle parity ok)", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
