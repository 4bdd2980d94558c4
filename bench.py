#!/usr/bin/env python
"""bench.py - throughput of ecCKD's hot path on B200 (one JSON line on stdout).

Workload (BASELINE.json configs[1], "create_look_up_table LW, 7 gases, 64 g
points"): the spectral reduction of one temperature index of the look-up-table
build = the 19 (gas, concentration) spectra that share that temperature's
Planck weights (composite, H2O x 12 concentrations, O3, CO2, CH4, N2O, CFC-11,
CFC-12; test/create_lut_lw.sh:146-200), each [53 layers][nwav = 2^23] FP64, 64 g
points over 13 bands, averaging_method "transmission", Planck weights computed
in the kernel.  The whole config is 6 such steps (114 blocks, 405 GB of optical
depth), which does not fit one GPU's HBM at once; one temperature index
(67.6 GB) does.  All data are synthetic and CKDMIP-shaped (no CKDMIP database).

  value  = point.layers/s of the device-resident reduction, whole job (all ranks).
  e2e    = the same metric through the reference-facing host-buffer call
           (planck_function + average_optical_depth_to_g_point on one block,
           host pinned buffers, H2D/D2H inside the timed region).
  N > 1  = the wavenumber axis is sharded (each rank owns its own range, weak
           scaling); per step the ranks all-reduce the per-g-point partial
           accumulators over NCCL and finalize redundantly.

`--impl reference` times the CPU oracle port of the reference algorithm
(oracle/, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LW_BAND_EDGES = [0.0, 350.0, 500.0, 630.0, 700.0, 820.0, 980.0, 1080.0, 1180.0, 1390.0, 1480.0, 1800.0, 2080.0,
                 3260.0]  # 13 "narrow" longwave bands (CKDMIP/RRTMG-like)
NG_PER_BAND = [6, 5, 5, 5, 5, 5, 4, 4, 5, 4, 5, 5, 6]  # 64 g points


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-nwav", type=int, default=23)
    ap.add_argument("--nspec", type=int, default=19)
    ap.add_argument("--nlay", type=int, default=53)
    ap.add_argument("--gmap", default="ranked", choices=["ranked", "random", "bands"],
                    help="ranked: g points from ranking a smooth synthetic spectrum (as find_g_points does); "
                         "random: no spectral coherence at all (adversarial)")
    ap.add_argument("--weights", default="planck", choices=["planck", "matrix"])
    ap.add_argument("--method", default="transmission")
    ap.add_argument("--cpu-log2-nwav", type=int, default=23, help="CPU arm block size (2^23 = the GPU arm's block)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU reduces its own 2^log2-nwav range (the driver's default); strong: ONE 2^log2-nwav "
                         "spectrum cut into --segments pieces shared out over the GPUs")
    ap.add_argument("--segments", type=int, default=0,
                    help="N-independent wavenumber segments per 2^log2-nwav spectrum (default 1 weak, 8 strong)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-flux", action="store_true")
    ap.add_argument("--no-full", action="store_true", help="skip the full look-up-table build leg (N = 1 only)")
    return ap.parse_args()


# ------------------------------------------------------------------ helpers
def pressure_grid(nlay):
    return np.exp(np.linspace(np.log(1.0), np.log(1.0e5), nlay + 1))


def temperature_profile(p_hl, offset=0.0):
    return 200.0 + 90.0 * (np.log(p_hl) / np.log(1.0e5)) + offset


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms
    period, so even a 100 ms region gets tens of samples); `nvidia-smi -lms` as the fallback."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None
        self.f = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.mx, self.reasons = [], [], set()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def _poll(self, nv, h):
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))]
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=float(np.median(self.sm)), sm_max_mhz=float(max(self.mx)),
                           reasons=sorted(self.reasons), samples=len(self.sm), source="nvml")
            return out
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm), source="nvidia-smi")
        return out


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ------------------------------------------------------------------ synthetic data on the device
def smooth_key(torch, wn, seed, ncomp=40):
    """Smooth multi-scale function of wavenumber (log of a synthetic absorption spectrum)."""
    rng = np.random.default_rng(seed)
    x = wn / 3260.0
    f = torch.zeros_like(wn)
    for _ in range(ncomp):
        freq = 2.0 ** rng.uniform(0, 17)  # features from the whole spectrum down to ~0.02 cm-1
        f += float(rng.normal() / np.sqrt(ncomp)) * torch.sin((2 * np.pi * freq) * x + float(rng.uniform(0, 2 * np.pi)))
    return f


def make_g_point(torch, wn, key, mode, seed):
    """g-point map: rank `key` inside each band, cut the rank axis at seeded fractions."""
    rng = np.random.default_rng(seed)
    g = torch.full((wn.numel(),), -1, dtype=torch.int32, device=wn.device)
    if mode == "random":
        key = torch.rand(wn.numel(), dtype=torch.float64, device=wn.device,
                         generator=torch.Generator(device=wn.device).manual_seed(seed))
    g0 = 0
    edges = LW_BAND_EDGES
    for b in range(len(edges) - 1):
        cuts = np.sort(rng.uniform(0.05, 0.95, NG_PER_BAND[b] - 1))
        if b < len(edges) - 2:
            m = (wn >= edges[b]) & (wn < edges[b + 1])
        else:
            m = (wn >= edges[b]) & (wn <= edges[b + 1])
        idx = torch.nonzero(m).flatten()
        n = idx.numel()
        if n > 0 and mode == "bands":  # one g point per band: a single lane segment per round (diagnostic)
            g[idx] = g0
        elif n > 0:
            order = torch.argsort(key[idx], stable=True)
            frac = (torch.arange(n, device=wn.device, dtype=torch.float64) + 0.5) / n
            gb = torch.bucketize(frac, torch.tensor(cuts, device=wn.device)).to(torch.int32) + g0
            g[idx[order]] = gb
        g0 += NG_PER_BAND[b]
    return g, g0


def make_spectrum(torch, wn, nlay, p_hl, seed, vmr):
    """One synthetic spectrum [nlay][nwav]: smooth absorption x layer mass x lognormal noise, 2 % zeros."""
    dev = wn.device
    gen = torch.Generator(device=dev).manual_seed(seed)
    base = torch.exp(2.5 * smooth_key(torch, wn, seed))
    mass = torch.tensor((p_hl[1:] - p_hl[:-1]) * vmr / (9.80665 * 0.001 * 28.970), device=dev)
    od = torch.empty((nlay, wn.numel()), dtype=torch.float64, device=dev)
    od.normal_(0.0, 0.3, generator=gen)
    od.exp_()
    od.mul_(base[None, :])
    od.mul_((3.0e-4 * mass)[:, None])
    for l in range(nlay):  # zeros sprinkled in (exercise the zero-handling branches), row by row to bound memory
        z = torch.rand(wn.numel(), device=dev, generator=gen) < 0.02
        od[l].masked_fill_(z, 0.0)
        # the CKDMIP spectra files store float32 (merge_spectra.cpp:108) and the reference widens on read:
        # the float64 values of the real workload are widened floats, and so are these
        od[l].copy_(od[l].float())
    return od


# ------------------------------------------------------------------ reference arm (CPU)
def use_all_host_threads():
    """The CPU arm runs on every host thread: torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would time the OpenMP oracle on one core (86 M instead of 560 M point*layers/s on the 16-thread box)."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


class CpuSample:
    """The oracle port of planck_function + average_optical_depth_to_g_point (the reference's
    per-block call pair, create_look_up_table.cpp:320-331) on a sample of the workload."""

    def __init__(self, args, log2_nwav):
        from tests.oracle_lib import oracle
        from tests import synthetic as S

        self.O = oracle()
        use_all_host_threads()
        self.method = args.method
        self.nwav = nwav = 1 << log2_nwav
        self.nlay = nlay = args.nlay
        rng = np.random.default_rng(7)
        self.p_hl = pressure_grid(nlay)
        self.p_fl = 0.5 * (self.p_hl[1:] + self.p_hl[:-1])
        t_hl = temperature_profile(self.p_hl)
        self.t_fl = 0.5 * (t_hl[:-1] * self.p_hl[:-1] + t_hl[1:] * self.p_hl[1:]) / self.p_fl
        self.wn = np.linspace(0.0002, 3260.0, nwav)
        self.dwn = np.full(nwav, self.wn[1] - self.wn[0])
        key = np.log(S.smooth_spectrum(self.wn, rng, ncomp=40))
        self.g, self.ng = S.g_point_map(self.wn, key, np.array(LW_BAND_EDGES), NG_PER_BAND, rng)
        row = np.exp(rng.normal(-3.0, 1.5, nwav))
        self.od = np.ascontiguousarray(row[None, :] * np.linspace(0.2, 3.0, nlay)[:, None])

    def run(self):
        t0 = time.perf_counter()
        planck = self.O.planck_function(self.t_fl, self.wn, self.dwn)
        self.O.average_optical_depth_to_g_point(self.ng, 4.0e-4, self.p_fl, self.p_hl, self.g, self.od, planck,
                                                self.method)
        return time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    steps, warm = args.steps, args.warmup
    sample = CpuSample(args, args.cpu_log2_nwav)
    for _ in range(warm):
        sample.run()
    total_t = sum(sample.run() for _ in range(steps))
    value = steps * args.nlay * sample.nwav / total_t
    line = {
        "impl": "reference", "metric": "LBL spectral points*layers/s reduced to LUT", "value": value,
        "unit": "point*layers/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"create_look_up_table LW (BASELINE configs[1]), one temperature index: {args.nspec} spectra x "
                               f"{args.nlay} layers x 2^{args.log2_nwav} wavenumbers per GPU, 64 g points / 13 bands, "
                               f"'{args.method}', weights={args.weights}, g map={args.gmap}",
                   "sample": "each step is ONE of the workload's %d spectra blocks (%d layers x 2^%d wavenumbers): "
                             "planck_function + average_optical_depth_to_g_point, the reference's per-block call pair "
                             "(create_look_up_table.cpp:320-331)" % (args.nspec, args.nlay, args.cpu_log2_nwav)},
        "cpu_baseline": {"value": value, "unit": "point*layers/s", "cores": cores, "kind": "port",
                         "sample": f"{args.nlay} x 2^{args.cpu_log2_nwav} per step; oracle/ecckd_oracle.cpp "
                                   "(C++/OpenMP restatement incl. the reference's per-g rescans), all host threads"},
        "e2e": {"value": value, "unit": "point*layers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import ecckd_b200 as E

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: the communicator is created (first collective) with
        # stdout pointed at stderr, so that the "NCCL version ..." banner does not land on it
        sys.stdout.flush()
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
    # One explicit non-default stream for everything (torch ops, NCCL, our kernels, the timing
    # events): the legacy default stream has handle 0, which the C ABI reads as "make a private stream".
    work_stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(work_stream)
    stream = work_stream.cuda_stream
    assert stream != 0
    ctx = E.Context(local_rank, stream)

    from ecckd_b200 import sharding

    nlay, nspec = args.nlay, args.nspec
    strong = args.scaling == "strong"
    nwav_glob = (1 << args.log2_nwav) * (1 if strong else world)
    nseg_total = (args.segments or (8 if strong else 1)) * (1 if strong else world)
    spans = sharding.segment_ranges(nwav_glob, nseg_total)
    seg_begin = sharding.segment_owners(nseg_total, world)
    my_segs = list(range(seg_begin[rank], seg_begin[rank + 1]))
    assert my_segs, "more ranks than segments"
    lo, hi = spans[my_segs[0]][0], spans[my_segs[-1]][1]
    nwav = hi - lo  # this rank's wavenumbers
    dnu = (3260.0 - 0.0002) / (nwav_glob - 1)
    wn = 0.0002 + dnu * (torch.arange(nwav, device=dev, dtype=torch.float64) + float(lo))
    dwn = torch.full((nwav,), dnu, dtype=torch.float64, device=dev)
    if strong and world > 1:  # the map of the WHOLE spectrum (ranks inside each band), this rank's slice of it
        wn_all = 0.0002 + dnu * torch.arange(nwav_glob, device=dev, dtype=torch.float64)
        g_all, ng = make_g_point(torch, wn_all, smooth_key(torch, wn_all, seed=7), args.gmap, seed=7)
        g_point = g_all[lo:hi].clone()
        del wn_all, g_all
    else:
        key = smooth_key(torch, wn, seed=7)
        g_point, ng = make_g_point(torch, wn, key, args.gmap, seed=7)
        del key
    p_hl = pressure_grid(nlay)
    p_fl = 0.5 * (p_hl[1:] + p_hl[:-1])
    t_hl = temperature_profile(p_hl)
    t_fl = 0.5 * (t_hl[:-1] * p_hl[:-1] + t_hl[1:] * p_hl[1:]) / p_fl
    vmrs = [1.0] + [1e-7 * 10 ** (0.5 * i) for i in range(12)] + [5e-6, 4.15e-4, 1.9e-6, 3.3e-7, 2.3e-10, 5e-10]
    vmrs = (vmrs * ((nspec + len(vmrs) - 1) // len(vmrs)))[:nspec]
    ods = [make_spectrum(torch, wn, nlay, p_hl, 1000 + s + 100 * rank, vmrs[s]) for s in range(nspec)]
    weight = None
    if args.weights == "matrix":
        weight = torch.empty((nlay, nwav), dtype=torch.float64, device=dev)
        E.planck_function_dev(ctx, t_fl, wn, dwn, nwav, weight)
    wmode = E.WEIGHT_PLANCK if weight is None else E.WEIGHT_MATRIX
    # One plan per segment; a segment's partial accumulators go to its slot of seg_sums / seg_minmax, the
    # exchange (grouped in-place NCCL broadcasts of the owners' slots, the only cross-GPU traffic) makes all
    # slots visible everywhere and combines them in segment order: bitwise the same for any number of GPUs.
    n3, n2 = nlay * nspec * ng * 3, nlay * nspec * ng * 2
    seg_sums = torch.zeros((nseg_total, n3), dtype=torch.float64, device=dev)
    seg_minmax = torch.zeros((nseg_total, n2), dtype=torch.float64, device=dev)
    seg_count = torch.zeros((nseg_total, ng), dtype=torch.int64, device=dev)
    seg = []  # (slot, offset into this rank's arrays, length, plan)
    for sidx in my_segs:
        a, b = spans[sidx][0] - lo, spans[sidx][1] - lo
        plan_s = E.GPointPlan(ctx, g_point[a:b], b - a, ng)
        seg_count[sidx].copy_(torch.from_numpy(plan_s.counts()))
        seg.append((sidx, a, b - a, plan_s))
    plan = seg[0][3]
    comm = None
    if world > 1:
        comm = E.Comm(ctx, rank, world, exchange=lambda box: dist.broadcast_object_list(box, src=0))
    sums = torch.empty(n3, dtype=torch.float64, device=dev)
    minmax = torch.empty(n2, dtype=torch.float64, device=dev)
    count = torch.zeros(ng, dtype=torch.int64, device=dev)
    # the per-g wavenumber counts are exchanged once, with the map
    E.average_optical_depth_exchange_dev(ctx, comm, seg_begin, None, None, None, None, seg_count, count)
    outs = [torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(nspec)]
    mins = [torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(nspec)]
    maxs = [torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(nspec)]

    def step():
        for sidx, a, n, plan_s in seg:
            E.average_optical_depth_partial_dev(
                ctx, plan_s, nlay, p_fl, [od[:, a:] for od in ods], nwav, wmode, None if weight is None else weight[:, a:],
                nwav, t_fl, wn[a:], dwn[a:], args.method, seg_sums[sidx], seg_minmax[sidx])
        E.average_optical_depth_exchange_dev(ctx, comm, seg_begin, seg_sums, seg_minmax, sums, minmax)
        E.average_optical_depth_finalize_dev(ctx, ng, count, vmrs, p_fl, p_hl, args.method, sums, minmax, outs, mins,
                                             maxs, ng)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ctx.set_timing(True)
    ctx.reduce_kernel_time()
    launches0 = ctx.kernel_launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    kms, klaunch = ctx.reduce_kernel_time()
    ctx.set_timing(False)
    launches = ctx.kernel_launches - launches0
    kstat = torch.tensor([kms / args.steps], dtype=torch.float64, device=dev)  # reduction kernel time per step
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kstat, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    units_per_step = nspec * nlay * nwav_glob
    value = units_per_step * args.steps / (total_ms * 1e-3)

    # roofline of the dominant kernel, algorithmic bytes per step per GPU (one launch per segment):
    # 8 B per point.layer of optical depth; per wavenumber 4 B g point + 8 B nu + 8 B dnu (Planck
    # weights recomputed) or + 8 B per point.layer of weights per distinct weight row (streamed).
    alg_bytes = nspec * nlay * nwav * 8 + nwav * 20
    if weight is not None:
        alg_bytes = nspec * nlay * nwav * 8 + nlay * nwav * 8 + nwav * 4
    k_ms = float(kstat.item())
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None  # DRAM bytes of one launch from the committed ncu capture, if it is of this workload
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "reduce_traffic.json")))
        w = t["workload"]
        if (w["nspec"], w["nlay"], w["log2_nwav"], w["method"], w["weights"], w["gmap"]) == (
                nspec, nlay, args.log2_nwav, args.method, args.weights, args.gmap):
            traffic = t["dram_bytes_per_launch"]
    except Exception:
        traffic = None
    peak, peak_src = measured_peak_gbs()
    kernel_name = "reduce_kernel" if os.environ.get("ECCKD_REDUCE_GEN") == "1" or os.environ.get("ECCKD_PLAN_WINDOW") == "0" \
        or args.method not in ("linear", "transmission", "transmission-2", "transmission-3", "transmission-10", "square-root") \
        else "reduce2_kernel"

    # ---- parity of what was just timed, at the full size: block 1 of the batch against the oracle
    # (planck_function + average_optical_depth_to_g_point on the same arrays, all host threads); the same
    # oracle run is the cpu_baseline sample, so that baseline and GPU arm are on the SAME block.
    validated, cpu, oracle_ref = None, None, None
    host = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from tests.oracle_lib import oracle
        from tests import synthetic as S
        O = oracle()
        ncores = use_all_host_threads()
        host = dict(od=ods[1].cpu().numpy(), wn=wn.cpu().numpy(), dwn=dwn.cpu().numpy(), g=g_point.cpu().numpy())
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            planck = O.planck_function(t_fl, host["wn"], host["dwn"])
            oracle_ref = O.average_optical_depth_to_g_point(ng, vmrs[1], p_fl, p_hl, host["g"], host["od"], planck,
                                                            args.method)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            del planck
        got = (outs[1].cpu().numpy(), mins[1].cpu().numpy(), maxs[1].cpu().numpy(), None)
        scale = S.out_scale(dict(vmr=vmrs[1], pressure_fl=p_fl, pressure_hl=p_hl))
        S.assert_avg_close(got, oracle_ref, method=args.method, scale=scale, what="bench output vs oracle: ")
        fin = np.isfinite(oracle_ref[0]) & (oracle_ref[0] != 0)
        rel = np.where(fin, np.abs(got[0] - oracle_ref[0]) / np.where(fin, np.abs(oracle_ref[0]), 1.0), 0.0)
        tol = S.avg_tolerance(oracle_ref[0], args.method, scale)
        iw = np.unravel_index(int(np.argmax(rel)), rel.shape)
        kd_tau = {"transmission": 1.66, "transmission-2": 3.32, "transmission-3": 4.98, "transmission-10": 16.6}.get(args.method, 0.0) \
            * abs(float(oracle_ref[0][iw])) / float(scale[iw[0]])
        validated = {"against": "oracle/ecckd_oracle.cpp on the timed step's block 1 (%d x 2^%d x %d g)" % (nlay, args.log2_nwav, ng),
                     "tolerance": "1e-12 relative (+ the transmission conditioning term, tests/synthetic.py)",
                     "max_rel_err": float(np.max(np.abs(got[0][fin] - oracle_ref[0][fin]) / np.abs(oracle_ref[0][fin]))),
                     "max_err_over_tolerance": float(np.max(np.where(fin, np.abs(got[0] - oracle_ref[0]) / tol, 0.0))),
                     "cells_within_1e-12": int(np.sum(rel <= 1e-12)), "cells": int(rel.size),
                     "worst_cell": {"layer": int(iw[0]), "g": int(iw[1]), "kD_tau": kd_tau,
                                    "note": "relative error is amplified by exp(kD tau)/(kD tau) where the averaged "
                                            "transmission is within rounding of 0 (tests/synthetic.py::assert_avg_close)"},
                     "min_max_exact": bool(np.array_equal(got[1], oracle_ref[1]) and np.array_equal(got[2], oracle_ref[2]))}
        cpu = {"value": nlay * nwav / best, "unit": "point*layers/s", "cores": ncores, "kind": "port",
               "sample": f"one block of the timed workload ({nlay} x 2^{args.log2_nwav}): planck_function + "
                         f"average_optical_depth_to_g_point, oracle/ecckd_oracle.cpp, OpenMP all threads, {best:.2f} s"}

    # ---- end to end: the create_look_up_table block loop through the C ABI with HOST buffers.
    # ecckd_lut_builder_* holds the g-point map, its plan and the wavenumber grid on the device (the tool
    # reads them once per gas); every step is one block: its optical depths go host->device inside the
    # timed region, its LUT slice comes back.  Two distinct blocks alternate.  `value` is the float32
    # path (the spectra files store float32; widened on the device, bit-identical), the float64 path and
    # the one-shot call (map + grid uploaded, plan looked up in the cache every call) are beside it.
    e2e = None
    if not args.no_e2e:
        pin = lambda t, dt=None: torch.empty(t.shape, dtype=dt or t.dtype, pin_memory=True).copy_(t)
        p_wn, p_dwn, p_g = pin(wn), pin(dwn), pin(g_point)
        h_wn, h_dwn, h_g = p_wn.numpy(), p_dwn.numpy(), p_g.numpy()
        blocks = [1, 2 % nspec]
        res = (np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros((nlay, ng)), np.zeros(ng, dtype=np.int32),
               np.zeros(nlay))
        t_hl_b = temperature_profile(p_hl)
        builder = E.LutBuilder(ctx, ng, h_g, h_wn, h_dwn)
        e2e_steps = max(2, min(args.steps, 6))

        def timed(call):
            for k in range(2):
                call(k)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(e2e_steps):
                call(k)
            e1.record()
            barrier()
            ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            return float(ems.item()) / e2e_steps

        def check_block(what):
            if oracle_ref is not None:  # the last call of every timed loop was on block 1
                from tests import synthetic as S
                S.assert_avg_close(res[:3] + (None,), oracle_ref, method=args.method,
                                   scale=S.out_scale(dict(vmr=vmrs[1], pressure_fl=p_fl, pressure_hl=p_hl)),
                                   what=f"e2e {what} vs oracle: ")

        order = blocks if (e2e_steps - 1) % 2 == 0 else blocks[::-1]  # the last timed call is on block 1

        h32 = [pin(ods[b], torch.float32) for b in blocks]
        np32 = {b: h.numpy() for b, h in zip(blocks, h32)}
        ms32 = timed(lambda k: builder.block(vmrs[order[k % 2]], p_hl, t_hl_b, np32[order[k % 2]], args.method, out=res))
        check_block("float32 builder")
        res32 = [r.copy() for r in res[:3]]
        del h32, np32
        h64 = [pin(ods[b]) for b in blocks]
        np64 = {b: h.numpy() for b, h in zip(blocks, h64)}
        ms64 = timed(lambda k: builder.block(vmrs[order[k % 2]], p_hl, t_hl_b, np64[order[k % 2]], args.method, out=res))
        check_block("float64 builder")
        same = all(np.array_equal(a, b) for a, b in zip(res32, res[:3]))
        res1 = res[:4]
        ms1 = timed(lambda k: E.planck_average_optical_depth_to_g_point(
            ctx, ng, vmrs[order[k % 2]], p_fl, p_hl, h_g, np64[order[k % 2]], t_fl, h_wn, h_dwn, args.method, out=res1))
        check_block("one-shot call")
        builder.close()
        del h64, np64
        per = lambda ms: world * nlay * nwav / (ms * 1e-3)
        e2e = {"value": per(ms32), "unit": "point*layers/s",
               "h2d_bytes_per_step": int(nlay * nwav * 4), "d2h_bytes_per_step": int(3 * nlay * ng * 8 + ng * 4),
               "call": "ecckd_lut_builder_block, float32 spectra as the CKDMIP files store them (one [%d][2^%d] block per "
                       "step from pinned host memory, two blocks alternating; g-point map, plan and wavenumber grid "
                       "resident in the builder)" % (nlay, args.log2_nwav),
               "ms_per_step": ms32, "float32_equals_float64_bitwise": bool(same),
               "float64_spectra": {"value": per(ms64), "unit": "point*layers/s", "ms_per_step": ms64,
                                   "h2d_bytes_per_step": int(nlay * nwav * 8),
                                   "call": "ecckd_lut_builder_block, float64 spectra"},
               "one_shot_call": {"value": per(ms1), "unit": "point*layers/s", "ms_per_step": ms1,
                                 "h2d_bytes_per_step": int(nlay * nwav * 8 + nwav * 20),
                                 "call": "ecckd_planck_average_optical_depth_to_g_point (float64; map and grid uploaded "
                                         "and the cached plan verified every call)",
                                 "plan_cache_hits_misses": list(ctx.plan_cache_stats())}}
    host = None

    # ---- the FULL look-up-table build of config B on one GPU (north_star's target sentence): g-point
    # compaction, plan, then for each of the 6 temperature indices the composite block merged from its
    # constituents (read_merged_spectrum.cpp:149-168) and the reduction of the 19 blocks with that index's
    # Planck weights, then gpoint_fraction and the Planck look-up table (create_look_up_table.cpp:111-160,
    # 293-505, 512-591).  The 405 GB of the whole config do not fit one GPU: the 6 temperature steps re-read
    # the resident 67.6 GB (every byte still comes from HBM, the spectra being far larger than L2).
    full = None
    if rank == 0 and world == 1 and not args.no_full and not strong and args.weights == "planck":
        try:
            full = bench_full_build(torch, E, ctx, dev, args, nlay, nspec, nwav, ng, wn, dwn, g_point, ods, vmrs, p_fl,
                                    p_hl, t_fl)
        except Exception as exc:  # a secondary leg must not hide the primary metric
            full = {"error": repr(exc)}

    # ---- metric 2: CKD flux columns.g-points/s (run_ckd LW, config A LUT with columns replicated)
    flux = None
    if not args.no_flux and rank == 0:
        try:
            flux = bench_flux(torch, E, ctx, dev, peak, with_cpu=not args.no_cpu)
        except Exception as exc:  # secondary metric must not hide the primary one
            flux = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "LBL spectral points*layers/s reduced to LUT", "value": value, "unit": "point*layers/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"create_look_up_table LW (BASELINE configs[1]), one temperature index: {nspec} spectra x "
                            f"{nlay} layers x 2^{args.log2_nwav} wavenumbers " + ("in total" if strong else "per GPU") +
                            f", {ng} g points / 13 bands, '{args.method}', weights={args.weights}, g map={args.gmap}",
                "sharding": f"{nseg_total} wavenumber segments (boundaries independent of the GPU count), contiguous runs per "
                            "rank; grouped in-place NCCL broadcasts of the per-segment partial accumulators, combined in "
                            "segment order (ecckd_average_optical_depth_exchange_dev): bitwise independent of the GPU count",
                "l2": "inputs (%.1f GB per GPU) exceed L2; no flush needed" % (nspec * nlay * nwav * 8 / 1e9),
                "exchange_bytes_received_per_step": int((n3 + n2) * 8 * (nseg_total - len(my_segs))) if world > 1 else 0,
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms, "launches_per_step": len(seg), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes // len(seg)),
                         "basis": "8 B per point.layer + 20 B per wavenumber per launch: the per-wavenumber data are "
                                  "shared by the %d spectra of the launch (%.2f B per point.layer); SURVEY.md 8d's "
                                  "per-unit figure for a single spectrum is 8.4 B, on which this kernel time is "
                                  "%.3f of the peak (its 70 %% target: 546 G point.layers/s per GPU)"
                                  % (nspec, alg_bytes / (nspec * nlay * nwav), 8.4 * nspec * nlay * nwav / (k_ms * 1e-3) / 1e9 / peak)
                         if weight is None else "8 B per point.layer + 8 B per point.layer of each distinct weight row"},
            "cpu_baseline": cpu,
            "validated": validated,
            "e2e": e2e,
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
        }
        if flux is not None:
            line["flux_solver"] = flux
        if full is not None:
            line["full_build"] = full
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_full_build(torch, E, ctx, dev, args, nlay, nspec, nwav, ng, wn, dwn, g_point, ods, vmrs, p_fl, p_hl, t_fl):
    """The whole LW look-up-table build (BASELINE configs[1]) as one timed sequence on one GPU; see the call site."""
    nt = 6
    g_work = g_point.clone()
    comp = torch.empty_like(ods[0])  # the composite block of a temperature index, merged from two constituents
    n_merge = 2
    outs = [[torch.empty((nlay, ng), dtype=torch.float64, device=dev) for _ in range(nspec)] for _ in range(3)]
    w1 = np.arange(0.0, 3260.0, 50.0)
    w2 = w1 + 50.0
    d_frac = torch.empty((ng, w1.size), dtype=torch.float64, device=dev)
    t_lut = np.arange(120.0, 351.0)
    d_plut = torch.empty((t_lut.size, ng), dtype=torch.float64, device=dev)
    state = {}

    marks = []

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, e))

    def build():
        del marks[:]
        mark("start")
        g_work.copy_(g_point)
        ng2, _ = E.compact_g_points_dev(ctx, ng, nwav, g_work)  # a4
        assert ng2 == ng
        plan = E.GPointPlan(ctx, g_work, nwav, ng)
        mark("compaction_and_plan")
        for it in range(nt):
            E.merge_spectra_dev(ctx, nlay, nwav, [ods[(13 + k) % nspec] for k in range(n_merge)], [nwav] * n_merge,
                                [0.5 + 0.25 * k for k in range(n_merge)], comp, nwav)  # a8, one pass
            E.average_optical_depth_batch_dev(ctx, plan, vmrs, p_fl, p_hl, [comp] + ods[1:], nwav, E.WEIGHT_PLANCK, None,
                                              nwav, t_fl + 20.0 * (it - 2.5), wn, dwn, args.method, outs[0], outs[1],
                                              outs[2], ng)
        mark("merges_and_reductions")
        E.gpoint_fraction_dev(ctx, ng, nwav, g_work, wn, dwn, w1, w2, d_frac)  # a5
        mark("gpoint_fraction")
        E.planck_lut_dev(ctx, plan, wn, dwn, t_lut, d_plut)  # a6
        mark("planck_lut")
        state["plan"] = plan

    build()  # warm-up (workspaces, plan kernels' first launch)
    torch.cuda.synchronize()
    state["plan"].close()
    ctx.set_timing(True)
    ctx.reduce_kernel_time()
    l0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    build()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    kms, _ = ctx.reduce_kernel_time()
    ctx.set_timing(False)
    launches = ctx.kernel_launches - l0
    # checks: the composite block's first layer against the same merge on the host, gpoint_fraction rows sum to 1
    h = sum((0.5 + 0.25 * k) * ods[(13 + k) % nspec][0].cpu().numpy() for k in range(n_merge))
    merge_ok = bool(np.allclose(comp[0].cpu().numpy(), h, rtol=1e-15, atol=0.0))
    frac = d_frac.cpu().numpy()
    frac_ok = bool(np.allclose(frac.sum(axis=1), 1.0, rtol=1e-12))
    state["plan"].close()
    units = nt * nspec * nlay * nwav
    return {"metric": "LBL spectral points*layers/s reduced to LUT, whole build", "value": units / (ms * 1e-3),
            "unit": "point*layers/s", "ms": ms, "blocks": nt * nspec, "gpu_launches": int(launches),
            "reduce_kernel_ms": kms, "reduce_kernel_share": kms / ms,
            "phases_ms": {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks[:-1], marks[1:])},
            "optical_depth_bytes_read": int(units * 8 + nt * n_merge * nlay * nwav * 8),
            "config": {"workload": f"create_look_up_table LW, all {nt} temperature indices x {nspec} blocks of {nlay} x 2^{args.log2_nwav}: "
                                   f"g-point compaction + plan + per index [merge of the composite block from {n_merge} constituents + "
                                   "reduction of the 19 blocks] + gpoint_fraction + Planck look-up table (231 temperatures); "
                                   "the resident 67.6 GB re-read per temperature index"},
            "validated": {"composite_merge_layer0_equals_host": merge_ok, "gpoint_fraction_rows_sum_to_1": frac_ok}}


def bench_flux(torch, E, ctx, dev, peak, with_cpu=True):
    """Metric 2, CKD flux columns*g-points/s: the fused run_ckd longwave core (ecckd_run_ckd_dev: total
    optical depth of 8 gases from the look-up tables, Planck function, flux sweeps, broadband sums;
    run_ckd.cpp:119-132, 265-356) on 2^16 columns x 54 layers x 32 g points (BASELINE configs[0]'s model with
    the columns replicated), checked against the oracle, with its own CPU baseline and host-buffer e2e."""
    from tests import synthetic as S

    ncol, nlay, ng, nbase = 1 << 16, 54, 32, 64
    c = S.ckd_model_case(ncol=nbase, nlay=nlay, ng=ng, seed=1234)
    model = E.CkdModel(ctx, c["gases"], c["log_pressure"], c["temperature"], c["temperature_planck"], c["planck_lut"])
    rep = ncol // nbase
    tile = lambda a: np.ascontiguousarray(np.tile(a, (rep, 1)))
    h_p, h_t = tile(c["pressure_hl"]), tile(c["temperature_hl"])
    h_vmr = [None if v is None else tile(v) for v in c["vmr_fl"]]
    t = lambda a: torch.from_numpy(a).to(dev)
    p_hl, t_hl = t(h_p), t(h_t)
    vmr = [None if v is None else t(v) for v in h_vmr]
    od = torch.empty((ncol, nlay, ng), dtype=torch.float64, device=dev)
    plh = torch.empty((ncol, nlay + 1, ng), dtype=torch.float64, device=dev)
    pls = torch.empty((ncol, ng), dtype=torch.float64, device=dev)
    dn, up = torch.empty_like(plh), torch.empty_like(plh)
    bdn = torch.empty((ncol, nlay + 1), dtype=torch.float64, device=dev)
    bup = torch.empty_like(bdn)
    run = lambda: model.run_ckd_dev(vmr, 1361.0, ncol, nlay, p_hl, t_hl, od, plh, pls, dn, up, bdn, bup)

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

    launches0 = ctx.kernel_launches
    run()
    launches = ctx.kernel_launches - launches0
    ms = timed(run)
    bytes_per = 8 * (54 + 3 * 55) + 8  # 1760 B per column.g-point (OD, Planck, two fluxes written once)

    # parity of the timed output: the first replica against the oracle's run_ckd, every replica against the first
    from tests.oracle_lib import oracle
    O = oracle()
    ref = O.run_ckd(c, is_sw=False)
    got = {"optical_depth": od, "planck_hl": plh, "flux_dn": dn, "flux_up": up, "bb_flux_dn": bdn, "bb_flux_up": bup}
    worst = {}
    for k, arr in got.items():
        a = arr[:nbase].cpu().numpy()
        if k == "optical_depth":
            np.testing.assert_allclose(a, ref[k], rtol=2e-11, atol=1e-18)
            worst[k + "_rel"] = float(np.max(np.abs(a - ref[k]) / np.maximum(np.abs(ref[k]), 1e-300)))
        elif k == "planck_hl":
            np.testing.assert_allclose(a, ref[k], rtol=1e-14)
        else:
            err = float(np.max(np.abs(a - ref[k])))
            assert err <= 1e-6, (k, err)  # north_star: fluxes within 1e-6 W m-2
            worst[k + "_abs_W_m2"] = err
        assert torch.equal(arr.view(rep, nbase, *arr.shape[1:])[-1], arr[:nbase]), k
    validated = {"against": "oracle run_ckd on the %d base columns; the %d replicas bitwise equal" % (nbase, rep), **worst}

    # the flux solver alone (radiative_transfer_lw.cpp:25-60) on resident optical depth and Planck arrays
    emis = torch.ones(ng, dtype=torch.float64, device=dev)
    ms_rt = timed(lambda: E.radiative_transfer_lw_dev(ctx, ncol, nlay, ng, plh, od, emis, pls, dn, up))
    gbs_rt = ncol * ng * bytes_per / (ms_rt * 1e-3) / 1e9
    # optical depth + Planck function alone (no sweep): bytes written 8*(54+55+1) per column.g-point + inputs
    ms_od = timed(lambda: model.run_ckd_dev(vmr, 1361.0, ncol, nlay, p_hl, t_hl, od, plh, pls))
    od_bytes = ncol * ng * 8 * (nlay + nlay + 2) + ncol * 8 * (2 * (nlay + 1) + 7 * nlay)

    # CPU baseline: the oracle's run_ckd (a scalar restatement, like the reference's serial column loop,
    # run_ckd.cpp:347) on every host thread, columns dealt out in chunks of 64
    cpu = None
    if with_cpu:
        from concurrent.futures import ThreadPoolExecutor
        nthr = os.cpu_count() or 1
        nchunk = max(nthr * 2, 8)
        one = lambda _: O.run_ckd(c, is_sw=False)
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(one, range(nthr)))
            t0 = time.perf_counter()
            list(ex.map(one, range(nchunk)))
            dt = time.perf_counter() - t0
        cpu = {"value": nchunk * nbase * ng / dt, "unit": "column*g-points/s", "cores": nthr, "kind": "port",
               "sample": f"{nchunk} chunks of {nbase} columns x {nlay} layers x {ng} g points through oracle run_ckd "
                         f"(orc_run_ckd, scalar per call) on {nthr} threads, {dt:.2f} s"}

    # end to end: ecckd_run_ckd with HOST arrays (profiles and 7 concentration arrays up; optical depth, Planck
    # function, spectral and broadband fluxes down), pinned
    pinned = lambda a: torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
    hp, ht = pinned(h_p), pinned(h_t)
    hp.copy_(torch.from_numpy(h_p))
    ht.copy_(torch.from_numpy(h_t))
    hv = []
    for v in h_vmr:
        if v is None:
            hv.append(None)
        else:
            b = pinned(v)
            b.copy_(torch.from_numpy(v))
            hv.append(b.numpy())
    outs = {"optical_depth": pinned(od), "planck_hl": pinned(plh), "planck_surf": pinned(pls), "flux_dn": pinned(dn),
            "flux_up": pinned(up), "bb_flux_dn": pinned(bdn), "bb_flux_up": pinned(bup)}
    outs_np = {k: v.numpy() for k, v in outs.items()}
    ms_e2e = timed(lambda: model.run_ckd(hv, hp.numpy(), ht.numpy(), out=outs_np), n=3, warm=1)
    assert np.array_equal(outs_np["flux_up"][:nbase], up[:nbase].cpu().numpy())
    h2d = sum(a.nbytes for a in [h_p, h_t] + [v for v in h_vmr if v is not None])
    d2h = sum(v.numel() * 8 for v in outs.values())
    model.close()

    rate = lambda m: ncol * ng / (m * 1e-3)
    gbs = ncol * ng * bytes_per / (ms * 1e-3) / 1e9
    return {"metric": "CKD flux columns*g-points/s", "value": rate(ms), "unit": "column*g-points/s", "ms_per_step": ms,
            "higher_is_better": True, "dtype": "f64", "gpu_launches": int(launches),
            "config": {"workload": f"run_ckd LW core, fused (ecckd_run_ckd_dev): {ncol} columns x {nlay} layers x {ng} g points, "
                                   "8 gases (BASELINE configs[0] model, its columns replicated); inputs resident"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "traffic": None, "kernel": "total_od_planck3_kernel + rt_lw_staged_kernel (whole fused call)",
                         "algorithmic_bytes_per_launch": int(ncol * ng * bytes_per)},
            "validated": validated,
            "cpu_baseline": cpu,
            "e2e": {"value": rate(ms_e2e), "unit": "column*g-points/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "call": "ecckd_run_ckd (host arrays, pinned)"},
            "parts": {"optical_depth_and_planck": {"ms": ms_od, "GBps": od_bytes / (ms_od * 1e-3) / 1e9,
                                                   "kernel": "total_od_planck3_kernel"},
                      "flux_sweep": {"ms": ms_rt, "GBps": gbs_rt, "frac": gbs_rt / peak, "kernel": "rt_lw_staged_kernel",
                                     "note": "radiative_transfer_lw on resident optical depth and Planck arrays"}}}


if __name__ == "__main__":
    main()
