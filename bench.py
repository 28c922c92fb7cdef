#!/usr/bin/env python
"""MuPS benchmark (the driver's contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "Nesti-Net MoE default"): every one of the 100 000 points of
a synthetic PCPNet-shape cloud is a query; 4 scales (0.01/0.03/0.05/0.07 of the bbox diagonal),
512-point patches, 8^3 Gaussian grid -> MuPS [100000, 8, 8, 8, 80] fp32 (16.4 GB) per cloud.
One step = the whole hot path for one such cloud per GPU: index build, ball query + subsample +
normalise, 3DmFV statistics.  With N GPUs a step covers N clouds (replicated on every rank); the
query list of each cloud is split into N contiguous shards and rank r computes shard r of every
cloud into its own slab: per-GPU work is fixed (weak scaling), no collective on the data path.

value  = query points/s, device-timed (CUDA events on the launching stream, max over ranks),
         cloud already resident in HBM.
e2e    = the same metric through MuPSPipeline.features_to_host: host cloud in (pinned), MuPS
         streamed back into pinned host memory, all copies inside the timed region.
roofline: the statistics kernel against the FP32 issue roof (46 FP32 op-slots + 1
         exp per unmasked (point, Gaussian) pair), its duration measured live with CUDA events.
--impl reference: the reference's CPU path on the host cores (scipy cKDTree + the oracle's
         OpenMP C port of get_3dmfv_n_est; TensorFlow 1.12 cannot be installed) on a bounded
         sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mups_query_points_per_s"
UNIT = "query points/s"
N_POINTS = 100000
RADIUS = [0.01, 0.03, 0.05, 0.07]
P = 512
RES = 8
VARIANCE = 0.0156            # the reference's command line for an 8^3 grid (train_n_est_w_experts.py:20)
SEED = 3627473
N_CLOUDS = 4                 # distinct synthetic clouds cycled through the steps
OPS_PER_PAIR = 46            # algorithmic FP32 op-slots per unmasked (point, Gaussian) pair
WORKLOAD = ("configs[1]: Nesti-Net MoE default - 4 scales (0.01/0.03/0.05/0.07), 512-pt patches, 8^3 grid, "
            "all 100k points of a synthetic PCPNet-shape cloud per GPU per step")


def config(n_gpus):
    return {"workload": WORKLOAD, "cloud_points": N_POINTS, "queries_per_gpu_per_step": N_POINTS,
            "clouds_per_step": n_gpus, "scales": RADIUS, "points_per_patch": P, "grid": "%dx%dx%d" % (RES, RES, RES),
            "gmm_variance": VARIANCE, "seed": SEED,
            "partitioning": "query points sharded across %d GPU(s), cloud replicated, per-rank slabs, no collective" % n_gpus,
            "schedule": "half 1 of cloud i+1 (side stream) overlaps half 2 of cloud i (main stream); MUPS_BENCH_PIPELINE=0 serialises",
            "l2": "each step writes 16.4 GB of MuPS + 2.5 GB of patches per GPU (>> 126 MB L2) and cycles "
                  "through %d clouds; no explicit flush needed" % N_CLOUDS}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler(object):
    """SM clock / power / throttle reasons sampled through NVML (what nvidia-smi reads) every 50 ms
    from a background thread.  NVML is initialised in __init__, i.e. before the warm-up steps, so
    that no driver start-up cost lands in the timed region; samples are kept only between
    start() and stop().  Falls back to an `nvidia-smi -lms` child process when pynvml is missing."""
    SMI_FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                  "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []          # (sm_mhz, max_mhz, power_w, reasons set)
        self.recording = False
        self.alive = True
        self.nvml = None
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: NVML enumerates physical GPUs
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            threading.Thread(target=self._poll_nvml, daemon=True).start()
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.SMI_FIELDS,
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                threading.Thread(target=self._poll_smi, daemon=True).start()
            except Exception:
                self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while self.alive:
            if self.recording:
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                    pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                    mask = int(get_reasons(self.handle))
                    self.samples.append((sm, self.max_mhz, pw, {k for k, b in bits.items() if mask & b}))
                except Exception:
                    pass
            time.sleep(0.05)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if not self.recording or len(f) < 7:
                continue
            try:
                self.samples.append((float(f[0]), float(f[1]), float(f[2]),
                                     {n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")}))
            except ValueError:
                pass

    def start(self):
        self.samples = []
        self.recording = True

    def stop(self):
        self.recording = False
        self.alive = False
        if self.proc is not None:
            self.proc.terminate()
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = [x[0] for x in self.samples]
        reasons = set()
        for x in self.samples:
            reasons |= x[3]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(x[1] for x in self.samples) if sm else None,
                "power_w_max": max(x[2] for x in self.samples) if sm else None, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi", "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm (the only places bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------

def cpu_reference_pass(pts, kdtree, query_idx, gmm_feed):
    """One bounded pass of the reference's CPU path: cKDTree ball queries on all cores (the
    reference's own dependency and predicate), the shared seeded selection, gather/centre/normalise
    (oracle restatement of pcpnet_dataset.py:286-343) and the OpenMP C port of get_3dmfv_n_est +
    MuPS assembly.  Returns the number of query points."""
    from oracle import c_oracle
    from oracle import mups_oracle as orc
    w, mu, sigma = gmm_feed
    rads = orc.absolute_radii(pts, RADIUS)
    S = len(rads)
    B = len(query_idx)
    patches = np.zeros((B, S * P, 3), np.float32)
    n_eff = np.zeros((B, S), np.int32)
    centres = pts[query_idx]
    for s, rad in enumerate(rads):
        lists = kdtree.query_ball_point(centres, rad, workers=-1)
        for b, inds in enumerate(lists):
            inds = np.asarray(inds, np.int64)
            n_eff[b, s] = min(P, len(inds))
            inds = orc.select_subset(inds, P, SEED, int(query_idx[b]), s)
            patches[b, s * P: s * P + len(inds)] = (pts[inds] - centres[b]) / np.float32(rad)
    c_oracle.mups(patches, n_eff, w, mu, sigma, S)
    return B


def cpu_sample_queries(step, n):
    return (np.arange(n, dtype=np.int64) * (N_POINTS // n) + step) % N_POINTS


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    from oracle import mups_oracle as orc
    c_oracle.build()
    c_oracle.set_num_threads(os.cpu_count())        # torchrun exports OMP_NUM_THREADS=1
    sample = 512
    clouds = [orc.synthetic_cloud(N_POINTS, cloud_id=i) for i in range(N_CLOUDS)]
    trees = [orc.build_kdtree(p) for p in clouds]
    feed = orc.gmm_feed(*orc.get_3d_grid_gmm([RES] * 3, VARIANCE))
    for i in range(args.warmup):
        cpu_reference_pass(clouds[i % N_CLOUDS], trees[i % N_CLOUDS], cpu_sample_queries(i, sample), feed)
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        done += cpu_reference_pass(clouds[i % N_CLOUDS], trees[i % N_CLOUDS], cpu_sample_queries(i, sample), feed)
    dt = time.perf_counter() - t0
    value = done / dt
    cores = os.cpu_count()
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config(args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads_half2": c_oracle.num_threads(),
                            "kind": "port",
                            "sample": "%d strided query points of the 100k per step (kd-tree build excluded); "
                                      "cKDTree.query_ball_point(workers=-1) + oracle C port (OpenMP) of get_3dmfv_n_est" % sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------

def fp32_micro_peak():
    """Measured FP32 FMA-pipe issue rate (T op-slots/s) from profiles/microbench, else None."""
    exe = os.path.join(ROOT, "profiles", "microbench")
    if not os.path.exists(exe):
        return None, None
    try:
        txt = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None, None
    ffma, mufu = [], []
    for line in txt.splitlines():
        try:
            d = json.loads(line)
        except ValueError:
            continue
        if d.get("test") == "ffma":
            ffma.append(d["glane_op_per_s"] / 1e3)
        if d.get("test") == "mufu_ex2":
            mufu.append(d["glane_op_per_s"] / 1e3)
    return (max(ffma) if ffma else None), (max(mufu) if mufu else None)


def run_own(args):
    import torch
    import torch.distributed as dist
    import nesti_net_b200 as mb
    from nesti_net_b200 import _lib
    from oracle import mups_oracle as orc       # synthetic clouds + the cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the MuPS path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    _lib.load()

    g = mb.get_3d_grid_gmm([RES] * 3, VARIANCE)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    S, G = len(RADIUS), gmm.G
    clouds_host = [orc.synthetic_cloud(N_POINTS, cloud_id=i) for i in range(max(N_CLOUDS, n_gpus))]
    clouds_dev = [torch.from_numpy(c).to(dev) for c in clouds_host]
    bounds = mb.dist.shard_bounds(N_POINTS, n_gpus)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    q_shard = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    per_cloud = hi - lo
    rows = per_cloud * n_gpus                                   # query points this rank computes per step
    feats = torch.empty((rows, RES, RES, RES, 20 * S), dtype=torch.float32, device=dev)
    patches = torch.empty((rows, S * P, 3), dtype=torch.float32, device=dev)
    n_eff = torch.empty((rows, S), dtype=torch.int32, device=dev)
    total = torch.empty((rows, S), dtype=torch.int32, device=dev)
    L = _lib.load()
    import ctypes
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)

    stat_events = []

    # Software pipeline across clouds.  Half 1 of the NEXT cloud (index build, the one small host read of the path --
    # bbox -> radii, evaluated with the reference's numpy expression -- and the ball-query kernel, which is latency /
    # barrier bound) runs on a high-priority side stream while half 2 (the FP32-bound statistics kernel) of the CURRENT
    # cloud runs on the main stream; patches are double buffered.  Every step still performs one index build, one ball
    # query and one statistics launch per cloud inside the timed region.  MUPS_BENCH_PIPELINE=0 serialises them.
    pipelined = os.environ.get("MUPS_BENCH_PIPELINE", "1") != "0"
    side = torch.cuda.Stream(dev, priority=-1)
    side_ptr = ctypes.c_void_p(side.cuda_stream)
    patches2 = [patches, torch.empty_like(patches) if pipelined else patches]
    n_eff2 = [n_eff, torch.empty_like(n_eff) if pipelined else n_eff]
    total2 = [total, torch.empty_like(total) if pipelined else total]
    stats_done = [None, None]

    def half1(i, c, slot):
        xyz = clouds_dev[(i * n_gpus + c) % len(clouds_dev)]
        with torch.cuda.stream(side):
            index = mb.PointIndex(xyz, cell_frac=max(RADIUS))
        radii = np.ascontiguousarray(index.absolute_radii(RADIUS), dtype=np.float64)
        if not pipelined:
            return index, radii, None
        sl = slice(c * per_cloud, (c + 1) * per_cloud)
        if stats_done[slot] is not None:
            side.wait_event(stats_done[slot])            # the statistics kernel that last read this patch buffer
        _lib.check(L.mups_ball_query(index.handle, ctypes.c_void_p(q_shard.data_ptr()), per_cloud,
                                     radii.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), S, P, SEED, None,
                                     ctypes.c_void_p(total2[slot][sl].data_ptr()), ctypes.c_void_p(patches2[slot][sl].data_ptr()),
                                     ctypes.c_void_p(n_eff2[slot][sl].data_ptr()), side_ptr))
        ev = torch.cuda.Event()
        ev.record(side)
        return index, radii, ev

    prefetched = {"next": [half1(0, c, 0) for c in range(n_gpus)]}

    def step(i, timed):
        """One pass of the hot path: n_gpus clouds, this rank's query shard of each."""
        slot = (i & 1) if pipelined else 0
        cur, nxt = prefetched["next"], []
        for c in range(n_gpus):
            index, radii, ev = cur[c]
            sl = slice(c * per_cloud, (c + 1) * per_cloud)
            if pipelined:
                stream.wait_event(ev)
            else:
                _lib.check(L.mups_ball_query(index.handle, ctypes.c_void_p(q_shard.data_ptr()), per_cloud,
                                             radii.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), S, P, SEED, None,
                                             ctypes.c_void_p(total[sl].data_ptr()), ctypes.c_void_p(patches[sl].data_ptr()),
                                             ctypes.c_void_p(n_eff[sl].data_ptr()), sptr))
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.mups_3dmfv(gmm.handle, ctypes.c_void_p(patches2[slot][sl].data_ptr()),
                                    ctypes.c_void_p(n_eff2[slot][sl].data_ptr()), per_cloud, S, P, _lib.FLAG_MASKED,
                                    ctypes.c_void_p(feats[sl].data_ptr()), sptr))
            e1.record(stream)
            if timed:
                stat_events.append((e0, e1))
        done = torch.cuda.Event()
        done.record(stream)
        stats_done[slot] = done
        for c in range(n_gpus):
            nxt.append(half1(i + 1, c, slot ^ 1))
        prefetched["next"] = nxt

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp32_peak_measured, mufu_peak_measured = (None, None)
    if rank == 0:
        fp32_peak_measured, mufu_peak_measured = fp32_micro_peak()

    sampler = ClockSampler(local_rank) if rank == 0 else None     # NVML comes up before the warm-up, not inside the timed region
    for i in range(args.warmup):
        step(i, False)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, True)
    t1.record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = rows * n_gpus * args.steps / (ms_total * 1e-3)

    # the dominant kernel: statistics.  Algorithmic work of the last step's launches on this rank.
    ne = n_eff2[(args.warmup + args.steps - 1) & 1 if pipelined else 0].cpu().numpy()
    m_unmasked = np.where(ne >= P - 1, P, ne + 1).astype(np.int64)
    pairs_step = float(m_unmasked.sum()) * G
    stat_ms = float(np.mean([a.elapsed_time(b) for a, b in stat_events])) * n_gpus      # per step (n_gpus launches)
    stats_tflops = OPS_PER_PAIR * pairs_step / (stat_ms * 1e-3) / 1e12
    sm_max_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_peak_nominal = 148 * 128 * sm_max_mhz * 1e6 / 1e12        # T op-slots/s (FMA counted once)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    out_bytes_step = float(rows) * G * 20 * S * 4
    roofline = {
        "kernel": "stats_separable_kernel (K5, 3DmFV statistics)",
        "bound": "fp32",
        "achieved": stats_tflops, "peak": fp32_peak_nominal, "unit": "TFLOP/s", "frac": stats_tflops / fp32_peak_nominal,
        "peak_is": "148 SMs x 128 FP32 lanes x %.0f MHz op-slots/s (FMA counted once, SURVEY.md 8d); MEASURED_PEAKS.json "
                   "holds only HBM/bf16 peaks" % sm_max_mhz,
        "peak_measured_ffma": fp32_peak_measured, "peak_measured_mufu_ex2": mufu_peak_measured,
        "algorithmic_ops_per_pair": OPS_PER_PAIR, "pairs_per_launch": pairs_step / n_gpus,
        "pairs_per_query": pairs_step / rows, "kernel_ms_per_launch": stat_ms / n_gpus,
        "kernel_share_of_step": stat_ms / (ms_total / args.steps),
        "mufu_exp_per_s_algorithmic": pairs_step / (stat_ms * 1e-3),
        "hbm": {"achieved": out_bytes_step / (stat_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": out_bytes_step / (stat_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_is": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
        "traffic": None,
    }
    # the second kernel of the step, timed alone (inside the loop it runs on the side stream under the statistics
    # kernel): algorithmic bytes = 16 B per true neighbour visited + the patches and counts written;
    # the 2.9 MB index of a 100 k-point cloud is L2-resident, so this is an L2/latency figure, not an HBM one
    torch.cuda.synchronize()
    bq_index = mb.PointIndex(clouds_dev[0], cell_frac=max(RADIUS))
    bq_radii = np.ascontiguousarray(bq_index.absolute_radii(RADIUS), dtype=np.float64)

    def bq_launch():
        _lib.check(L.mups_ball_query(bq_index.handle, ctypes.c_void_p(q_shard.data_ptr()), per_cloud,
                                     bq_radii.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), S, P, SEED, None,
                                     ctypes.c_void_p(total[:per_cloud].data_ptr()), ctypes.c_void_p(patches[:per_cloud].data_ptr()),
                                     ctypes.c_void_p(n_eff[:per_cloud].data_ptr()), sptr))
    for _ in range(2):
        bq_launch()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for _ in range(5):
        bq_launch()
    b1.record(stream)
    torch.cuda.synchronize()
    bq_ms = b0.elapsed_time(b1) / 5
    nbrs = float(total[:per_cloud].sum().item())
    bq_bytes = 16.0 * nbrs + per_cloud * (12.0 * S * P + 8.0 * S)
    roofline["ball_query"] = {"kernel": "ball_query_kernel (K3+K4), timed alone", "ms_per_launch": bq_ms,
                              "queries_per_launch": per_cloud, "mean_neighbours_per_query": nbrs / per_cloud,
                              "algorithmic_GBps": bq_bytes / (bq_ms * 1e-3) / 1e9, "hbm_peak_GBps": hbm_peak,
                              "frac_of_hbm_peak": bq_bytes / (bq_ms * 1e-3) / 1e9 / hbm_peak,
                              "note": "index is L2-resident at this cloud size; latency/barrier bound (profiles/r01_ball_query.md)"}
    del bq_index
    prof = os.path.join(ROOT, "profiles", "stats_kernel_traffic.json")
    if os.path.exists(prof):
        try:
            t = json.load(open(prof))
            roofline["traffic"] = t["dram_bytes_per_query"] * rows / n_gpus
            roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    skip = set(os.environ.get("MUPS_BENCH_SKIP", "").split(","))      # profiling runs only: "e2e,cpu"
    # ---- end to end through the public API with host buffers ------------------------------------------
    # chunk: query points per pipeline stage.  Small enough that the first device->host copy starts ~1.5 ms after the
    # call (each cloud's call fills and drains the pipeline), large enough (336 MB per copy) for full PCIe rate.
    chunk = int(os.environ.get("MUPS_BENCH_CHUNK", "2048"))
    if world > 1:
        mb.dist.bind_to_gpu_numa_node(local_rank)     # pinned staging buffers on the GPU's NUMA node (no-op where sysfs has none)
    pipe = mb.MuPSPipeline(gmm, RADIUS, P, seed=SEED, chunk=chunk if "e2e" not in skip else 64)
    hosts = [torch.from_numpy(c).pin_memory() for c in clouds_host]
    q_host = torch.arange(lo, hi, dtype=torch.int64).pin_memory()
    e2e_steps = max(1, min(args.steps, 3))
    if "e2e" in skip:
        q_host = q_host[:128]

    def e2e_step(i):
        n = 0
        for c in range(n_gpus):
            n += pipe.features_to_host(hosts[(i * n_gpus + c) % len(hosts)], q_host)
        return n

    e2e_step(0)
    barrier()
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    w0 = time.perf_counter()
    n_done = 0
    for i in range(e2e_steps):
        n_done += e2e_step(1 + i)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    barrier()
    e2e = {"value": n_done * n_gpus / float(e2e_s.item()), "unit": UNIT, "steps": e2e_steps,
           "h2d_bytes_per_step": pipe.h2d_bytes // e2e_steps, "d2h_bytes_per_step": pipe.d2h_bytes // e2e_steps,
           "chunk_queries": pipe.chunk,
           "api": "MuPSPipeline.features_to_host (pinned host cloud in, MuPS rows streamed to pinned host memory)"}

    # the same call with the consumer on the device (what the reference's inference loop does with MuPS: it feeds the
    # 3D-CNN and never visits the host): host cloud in, features reduced to one checksum on the GPU, 8 bytes read back
    if "e2e" not in skip:
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        pipe_dev = mb.MuPSPipeline(gmm, RADIUS, P, seed=SEED, chunk=16384)     # no PCIe stage to feed: larger chunks

        def on_device(lo_, hi_, rows_):
            acc.add_(rows_.sum())

        def dev_step(i):
            n = 0
            for c in range(n_gpus):
                n += pipe_dev.features_to_consumer(hosts[(i * n_gpus + c) % len(hosts)], q_host, on_device)
            return n

        dev_step(0)
        barrier()
        pipe_dev.h2d_bytes = 0
        w0 = time.perf_counter()
        n_dev = 0
        for i in range(e2e_steps):
            n_dev += dev_step(1 + i)
        checksum = float(acc.item())                       # the device->host read of the step's result
        dc_s = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dc_s, op=dist.ReduceOp.MAX)
        barrier()
        e2e["device_consumer"] = {"value": n_dev * n_gpus / float(dc_s.item()), "unit": UNIT,
                                  "h2d_bytes_per_step": pipe_dev.h2d_bytes // e2e_steps, "d2h_bytes_per_step": 8,
                                  "chunk_queries": pipe_dev.chunk,
                                  "checksum_finite": bool(np.isfinite(checksum)),
                                  "api": "MuPSPipeline.features_to_consumer (host cloud in, MuPS reduced on the GPU)"}

    cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": "measured at N=1 only (see the N=1 line / --impl reference)"}
    if rank == 0 and n_gpus == 1:
        # ---- CPU baseline on a bounded sample (the oracle is the thing timed here, never the product) ----
        from oracle import c_oracle
        c_oracle.build()
        c_oracle.set_num_threads(os.cpu_count())
        feed = orc.gmm_feed(*orc.get_3d_grid_gmm([RES] * 3, VARIANCE))
        tree = orc.build_kdtree(clouds_host[0])
        sample = 1024 if "cpu" not in skip else 16
        cpu_reference_pass(clouds_host[0], tree, cpu_sample_queries(0, 64 if "cpu" not in skip else 8), feed)
        c0 = time.perf_counter()
        done = cpu_reference_pass(clouds_host[0], tree, cpu_sample_queries(1, sample), feed)
        cdt = time.perf_counter() - c0
        # the reference as its scripts run it: one process, one kd-tree query per patch and scale,
        # the un-fused op chain of get_3dmfv_n_est (numpy transliteration) -- a small sample, reported next to the
        # all-core figure
        shipped = None
        if "cpu" not in skip:
            qs = cpu_sample_queries(2, 16)
            s0 = time.perf_counter()
            rads_ = orc.absolute_radii(clouds_host[0], RADIUS)
            pp = np.zeros((len(qs), S * P, 3), np.float32)
            ne_ = np.zeros((len(qs), S), np.int32)
            for b_, c_ in enumerate(qs):
                for s_, rad_ in enumerate(rads_):
                    inds_ = np.asarray(tree.query_ball_point(clouds_host[0][c_], rad_), np.int64)
                    ne_[b_, s_] = min(P, len(inds_))
                    inds_ = orc.select_subset(inds_, P, SEED, int(c_), s_)
                    pp[b_, s_ * P: s_ * P + len(inds_)] = (clouds_host[0][inds_] - clouds_host[0][c_]) / np.float32(rad_)
            orc.mups_assemble(pp, feed[0], feed[1], feed[2], ne_, S)
            sdt = time.perf_counter() - s0
            shipped = {"value": len(qs) / sdt, "unit": UNIT, "cores": 1, "sample": "%d query points, one process" % len(qs)}
        cpu_baseline = {"value": done / cdt, "unit": UNIT, "cores": os.cpu_count(), "threads_half2": c_oracle.num_threads(),
                        "kind": "port", "as_shipped_one_process": shipped,
                        "sample": "%d strided query points of cloud 0 (kd-tree build excluded); cKDTree.query_ball_point("
                                  "workers=-1) + oracle C port (OpenMP) of get_3dmfv_n_est" % sample}
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config(n_gpus), "roofline": roofline,
               "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(obj):
    """The ONE JSON line of the contract, on the real stdout."""
    f = _JSON_OUT or sys.stdout
    f.write(json.dumps(obj) + "\n")
    f.flush()


def main():
    global _JSON_OUT
    # Library chatter (e.g. NCCL's version banner) is written to file descriptor 1 by C code: keep a private
    # handle on the real stdout for the JSON line and point fd 1 at stderr for everything else.
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
