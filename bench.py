#!/usr/bin/env python
"""MuPS benchmark (the driver's contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1|c2|c4|c5] [--gather]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload (BASELINE.json configs[1], "Nesti-Net MoE default", `--config c2`): every one of the 100 000 points
of a synthetic PCPNet-shape cloud is a query; 4 scales (0.01/0.03/0.05/0.07 of the bbox diagonal), 512-point patches,
8^3 Gaussian grid -> MuPS [100000, 8, 8, 8, 80] fp32 (16.4 GB) per cloud.  One step = the whole hot path for one such
cloud per GPU: index build, ball query + subsample + normalise, 3DmFV statistics.  With N GPUs a step covers N clouds
(replicated on every rank); the query list of each cloud is split into N contiguous shards and rank r computes shard r
of every cloud into its own slab: per-GPU work is fixed (weak scaling), no collective on the data path.

The other BASELINE configs are parity-test cases, not the headline; `--config` runs them through the same code:
  c1  configs[0]: the same cloud with 3 scales (0.01/0.03/0.07)                                   (weak scaling)
  c4  configs[3]: 2 M-point scan-shape cloud (non-uniform density), all 2 M points are queries, the query list (in
      scanner-sweep order) sharded across the GPUs in contiguous ranges balanced by estimated work  (strong scaling)
  c5  configs[4]: 10 M-point cloud, 1 048 576 strided queries sharded the same way; grid 8^3 / 16^3 and
      P = 256 / 512 / 1024 as `variants` of the line                                                (strong scaling)
`--gather` (with N > 1): one cloud per step, queries sharded, every rank's slab delivered to rank 0 -- by the statistics
kernel's own stores over NVLink peer memory (dist.PeerSlabGather) and, for comparison, by NCCL all-gather.

value  = query points/s, device-timed (CUDA events on the launching stream, max over ranks), cloud resident in HBM.
e2e    = the same metric through the public API with HOST buffers: MuPSPipeline.features_to_host (pinned host cloud
         in, MuPS streamed back into pinned host memory, all copies inside the timed region).
roofline: the statistics kernel against the FP32 issue roof (46 FP32 op-slots + 1 exp per unmasked
         (point, Gaussian) pair), duration measured live with CUDA events; `frac` is that ALGORITHMIC ratio (it exceeds
         1 because the lattice fast path executes far fewer instructions than the count credits), `frac_executed` is
         the machine utilisation (issued warp instructions / issue slots), `general_kernel` the same ratio for the
         kernel that really executes the 46-op count; `ball_query` and `index_build` carry their HBM / L2 figures.
cpu_baseline / --impl reference: the reference's CPU path on the host cores (scipy cKDTree on all cores + a C port of
         get_3dmfv_n_est; TensorFlow 1.12 cannot be installed): the LITERAL port (one divide, powf and expf per pair,
         like the TF op chain) and a TUNED implementation (hoisted constants, vectorised, register-blocked) are both
         reported; `value` is the tuned one -- the stronger denominator.
"""
import argparse
import ctypes
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
SEED = 3627473
N_CLOUDS = 4                 # distinct synthetic clouds cycled through the steps (weak-scaling configs)
OPS_PER_PAIR = 46            # algorithmic FP32 op-slots per unmasked (point, Gaussian) pair

CONFIGS = {
    "c1": dict(workload="configs[0]: 3 scales (0.01/0.03/0.07), 512-pt patches, 8^3 grid, all 100k points of a synthetic "
                        "PCPNet-shape cloud per GPU per step", mode="clouds", n_points=100000, kind="pcpnet",
               radius=[0.01, 0.03, 0.07], P=512, res=8, variance=0.0156),
    "c2": dict(workload="configs[1]: Nesti-Net MoE default - 4 scales (0.01/0.03/0.05/0.07), 512-pt patches, 8^3 grid, "
                        "all 100k points of a synthetic PCPNet-shape cloud per GPU per step", mode="clouds", n_points=100000,
               kind="pcpnet", radius=[0.01, 0.03, 0.05, 0.07], P=512, res=8, variance=0.0156),
    "c4": dict(workload="configs[3]: dense scan-shape cloud (2M points, non-uniform density), 4 scales, 512-pt patches, 8^3 "
                        "grid, all 2M points are queries, sharded across the GPUs", mode="dense", n_points=2000000,
               kind="scan", cloud_id=1, radius=[0.01, 0.03, 0.05, 0.07], P=512, res=8, variance=0.0156, queries=None,
               cpu_sample=256, variants=[]),
    "c5": dict(workload="configs[4]: 10M-point cloud, 4 scales, 512-pt patches, 8^3 grid, 1 048 576 strided queries per step, "
                        "sharded across the GPUs; variants: 16^3 grid, 256/1024-pt patches", mode="dense", n_points=10000000,
               kind="pcpnet", cloud_id=2, radius=[0.01, 0.03, 0.05, 0.07], P=512, res=8, variance=0.0156, queries=1 << 20,
               cpu_sample=128, variants=[dict(res=8, P=256), dict(res=8, P=1024), dict(res=16, P=512)]),
}


def grid_variance(res):
    return 0.0156 if res == 8 else (1.0 / res) ** 2     # the reference's command line for 8^3; (1/res)^2 otherwise


def config(cfg, n_gpus, extra=None):
    out = {"workload": cfg["workload"], "cloud_points": cfg["n_points"], "scales": cfg["radius"],
           "points_per_patch": cfg["P"], "grid": "%dx%dx%d" % ((cfg["res"],) * 3), "gmm_variance": cfg["variance"], "seed": SEED}
    if cfg["mode"] == "clouds":
        out.update({"queries_per_gpu_per_step": cfg["n_points"], "clouds_per_step": n_gpus,
                    "partitioning": "query points sharded across %d GPU(s), cloud replicated, per-rank slabs, no collective" % n_gpus,
                    "schedule": "half 1 of the next step's clouds (side stream; index builds, one bbox read, ball queries) overlaps "
                                "half 2 of the current step (main stream: ONE statistics launch over all slabs); "
                                "MUPS_BENCH_PIPELINE=0 serialises",
                    "l2": "each step writes 16.4 GB of MuPS + 2.5 GB of patches per GPU (>> 126 MB L2) and cycles "
                          "through %d clouds; no explicit flush needed" % N_CLOUDS})
    else:
        nq = cfg["queries"] or cfg["n_points"]
        out.update({"queries_per_step": nq,
                    "partitioning": "one cloud, replicated; the query list in scanner-sweep (z) order is cut into %d contiguous "
                                    "ranges balanced by estimated work (neighbour count of every 64th query); per-rank "
                                    "slabs streamed in chunks and consumed on the device; no collective" % n_gpus,
                    "l2": "the index alone (%d MB) exceeds or rivals the 126 MB L2 and every chunk writes > 2 GB of "
                          "features; no explicit flush needed" % (cfg["n_points"] * 28 // 1000000)})
    if extra:
        out.update(extra)
    return out


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

def make_cloud(cfg, cloud_id):
    from nesti_net_b200.synthetic import synthetic_cloud
    return synthetic_cloud(cfg["n_points"], cloud_id=cfg.get("cloud_id", cloud_id), kind=cfg["kind"])


def cpu_half1(cfg, pts, kdtree, query_idx):
    """Half 1 of the reference's CPU path on all cores: cKDTree ball queries (the reference's own dependency and
    predicate, workers=-1), then the shared seeded selection and gather/centre/normalise of pcpnet_dataset.py:310-343 in
    C/OpenMP (oracle_half1_gather; bit-identical to oracle/mups_oracle.py::gather_patches)."""
    from oracle import c_oracle
    from oracle import mups_oracle as orc
    P, radius = cfg["P"], cfg["radius"]
    rads = orc.absolute_radii(pts, radius)
    S, B = len(rads), len(query_idx)
    patches = np.zeros((B, S * P, 3), np.float32)
    n_eff = np.zeros((B, S), np.int32)
    centres = pts[query_idx]
    for s, rad in enumerate(rads):
        lists = kdtree.query_ball_point(centres, rad, workers=-1)
        c_oracle.half1_gather(pts, query_idx, rad, lists, P, S, s, SEED, patches, n_eff)
    return patches, n_eff


def cpu_reference_pass(cfg, pts, kdtree, query_idx, gmm_feed, tuned=True):
    """One bounded pass of the reference's CPU path (both halves).  Returns (queries, seconds of half 1, of half 2)."""
    from oracle import c_oracle
    w, mu, sigma = gmm_feed
    t0 = time.perf_counter()
    patches, n_eff = cpu_half1(cfg, pts, kdtree, query_idx)
    t1 = time.perf_counter()
    (c_oracle.mups_tuned if tuned else c_oracle.mups)(patches, n_eff, w, mu, sigma, len(cfg["radius"]))
    t2 = time.perf_counter()
    return len(query_idx), t1 - t0, t2 - t1


def cpu_sample_queries(cfg, step, n):
    return (np.arange(n, dtype=np.int64) * (cfg["n_points"] // n) + step) % cfg["n_points"]


def cpu_baseline_entry(cfg, pts, sample, with_shipped):
    """cpu_baseline of the own arm (rank 0, N = 1): literal and tuned port on a bounded strided sample."""
    from oracle import c_oracle
    from oracle import mups_oracle as orc
    c_oracle.build()
    march = c_oracle.build_tuned_native()
    c_oracle.set_num_threads(os.cpu_count())
    feed = orc.gmm_feed(*orc.get_3d_grid_gmm([cfg["res"]] * 3, cfg["variance"]))
    tree = orc.build_kdtree(pts)
    cpu_reference_pass(cfg, pts, tree, cpu_sample_queries(cfg, 0, max(8, sample // 16)), feed)                 # warm
    n, h1, h2t = cpu_reference_pass(cfg, pts, tree, cpu_sample_queries(cfg, 1, sample), feed, tuned=True)
    _, h1b, h2l = cpu_reference_pass(cfg, pts, tree, cpu_sample_queries(cfg, 1, sample), feed, tuned=False)
    shipped = None
    if with_shipped:
        # the reference as its scripts run it: one process, one kd-tree query per patch and scale, the
        # un-fused op chain of get_3dmfv_n_est (numpy transliteration) -- a small sample, reported next to the all-core figures
        P, radius = cfg["P"], cfg["radius"]
        S = len(radius)
        qs = cpu_sample_queries(cfg, 2, 16)
        s0 = time.perf_counter()
        rads_ = orc.absolute_radii(pts, radius)
        pp = np.zeros((len(qs), S * P, 3), np.float32)
        ne_ = np.zeros((len(qs), S), np.int32)
        for b_, c_ in enumerate(qs):
            for s_, rad_ in enumerate(rads_):
                inds_ = np.asarray(tree.query_ball_point(pts[c_], rad_), np.int64)
                ne_[b_, s_] = min(P, len(inds_))
                inds_ = orc.select_subset(inds_, P, SEED, int(c_), s_)
                pp[b_, s_ * P: s_ * P + len(inds_)] = (pts[inds_] - pts[c_]) / np.float32(rad_)
        orc.mups_assemble(pp, feed[0], feed[1], feed[2], ne_, S)
        shipped = {"value": len(qs) / (time.perf_counter() - s0), "unit": UNIT, "cores": 1,
                   "sample": "%d query points, one process" % len(qs)}
    return {"value": n / (h1 + h2t), "unit": UNIT, "cores": os.cpu_count(), "threads_half2": c_oracle.num_threads(), "kind": "port",
            "what": "TUNED port: cKDTree.query_ball_point(workers=-1) + selection on a thread pool + oracle/mups_oracle_tuned.c "
                    "(hoisted constants, vectorised, register-blocked, -O3 -march=%s -ffast-math, OpenMP)" % march,
            "literal_port": {"value": n / (h1b + h2l), "unit": UNIT,
                             "what": "the same half 1 + oracle/mups_oracle.c: the TF op chain transliterated (divide, powf, "
                                     "expf per pair), -O3 -march=x86-64-v3, no fast-math, OpenMP"},
            "seconds": {"half1": h1, "half2_tuned": h2t, "half2_literal": h2l},
            "as_shipped_one_process": shipped,
            "sample": "%d strided query points of cloud 0 (kd-tree build excluded)" % sample}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_oracle
    from oracle import mups_oracle as orc
    c_oracle.build()
    march = c_oracle.build_tuned_native()
    c_oracle.set_num_threads(os.cpu_count())        # torchrun exports OMP_NUM_THREADS=1
    dense = cfg["mode"] == "dense"
    sample = cfg["cpu_sample"] if dense else 512
    n_clouds = 1 if dense else N_CLOUDS
    clouds = [make_cloud(cfg, i) for i in range(n_clouds)]
    trees = [orc.build_kdtree(p) for p in clouds]
    feed = orc.gmm_feed(*orc.get_3d_grid_gmm([cfg["res"]] * 3, cfg["variance"]))
    for i in range(args.warmup):
        cpu_reference_pass(cfg, clouds[i % n_clouds], trees[i % n_clouds], cpu_sample_queries(cfg, i, sample), feed)
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        done += cpu_reference_pass(cfg, clouds[i % n_clouds], trees[i % n_clouds], cpu_sample_queries(cfg, i, sample), feed)[0]
    dt = time.perf_counter() - t0
    value = done / dt
    # the literal port on one such sample, reported beside the tuned figure
    l0 = time.perf_counter()
    nl = cpu_reference_pass(cfg, clouds[0], trees[0], cpu_sample_queries(cfg, 0, sample), feed, tuned=False)[0]
    literal = nl / (time.perf_counter() - l0)
    cores = os.cpu_count()
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
           "scaling": "strong" if dense else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config(cfg, args.gpus, {"cpu_sample_per_step": "%d strided query points of one cloud (the rate, not the "
                                                                     "step, is the metric)" % sample}),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "threads_half2": c_oracle.num_threads(),
                            "kind": "port",
                            "what": "TUNED port (the stronger denominator): cKDTree.query_ball_point(workers=-1) + selection "
                                    "on a thread pool + oracle/mups_oracle_tuned.c (-O3 -march=%s -ffast-math, OpenMP)" % march,
                            "literal_port": {"value": literal, "unit": UNIT,
                                             "what": "oracle/mups_oracle.c, the TF op chain transliterated (divide, powf, expf per pair)"},
                            "sample": "%d strided query points per step (kd-tree build excluded)" % sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ------------------------------------------------------------------------------------------------
# own arm: helpers
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


def load_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def unmasked_pairs(n_eff_np, P, G):
    m = np.where(n_eff_np >= P - 1, P, n_eff_np + 1).astype(np.int64)       # tf_util.py:696: slot r is masked iff r > n_eff
    return float(m.sum()) * G


def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the MuPS path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def make_barrier(world):
    import torch
    import torch.distributed as dist

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return barrier


def max_over_ranks(value, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def stats_roofline(pairs, stat_ms, out_bytes, clocks, peaks, n_launches=1):
    """The statistics kernel against the FP32 issue roof: algorithmic ratio + executed utilisation."""
    sm_max_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = 148 * 128 * sm_max_mhz * 1e6 / 1e12                       # T op-slots/s (FMA counted once)
    tops = OPS_PER_PAIR * pairs / (stat_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    r = {"kernel": "stats_separable_kernel (K5, 3DmFV statistics)", "bound": "fp32",
         "achieved": tops, "peak": peak, "unit": "TFLOP/s", "frac": tops / peak,
         "frac_is": "ALGORITHMIC ratio: 46 op-slots x unmasked pairs / time / issue peak.  It exceeds 1 because the lattice fast "
                    "path executes ~22 warp instructions per 32 pairs and 24 instead of 512 exps per point; the machine "
                    "utilisation is frac_executed, and general_kernel is the kernel for which the 46-op count is physical",
         "peak_is": "148 SMs x 128 FP32 lanes x %.0f MHz op-slots/s (FMA counted once, SURVEY.md 8d); MEASURED_PEAKS.json "
                    "holds only HBM/bf16 peaks" % sm_max_mhz,
         "algorithmic_ops_per_pair": OPS_PER_PAIR, "pairs_per_launch": pairs / n_launches,
         "kernel_ms_per_launch": stat_ms / n_launches, "T_pairs_per_s": pairs / (stat_ms * 1e-3) / 1e12,
         "mufu_exp_per_s_algorithmic": pairs / (stat_ms * 1e-3),
         "hbm": {"achieved": out_bytes / (stat_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": out_bytes / (stat_ms * 1e-3) / 1e9 / hbm_peak,
                 "peak_is": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"},
         "traffic": None}
    ex = load_json("stats_kernel_executed.json")
    if ex:
        warp_inst = ex["warp_inst_per_pair"] * pairs                  # issued warp instructions of the timed launches
        slots = 148 * 4 * sm_max_mhz * 1e6 * (stat_ms * 1e-3)         # 4 schedulers per SM, 1 warp instruction / clk each
        r["frac_executed"] = warp_inst / slots
        r["frac_executed_is"] = ("issued warp instructions / issue slots (148 SMs x 4 schedulers x clk): smsp__inst_executed.sum per "
                                 "unmasked pair from the kept ncu capture (%s) x the pairs of this run / the live kernel time"
                                 % ex.get("source", "profiles/"))
        for k in ("issue_active_pct", "pipe_fma_pct", "pipe_alu_pct", "pipe_xu_pct"):
            if k in ex:
                r["ncu_" + k] = ex[k]
    return r


# ------------------------------------------------------------------------------------------------
# own arm: weak-scaling configs (c1, c2)
# ------------------------------------------------------------------------------------------------

def run_clouds(args, cfg):
    import torch
    import torch.distributed as dist
    import nesti_net_b200 as mb
    from nesti_net_b200 import _lib

    world, rank, local_rank, dev = setup_dist()
    n_gpus = world
    L = _lib.load()
    RADIUS, P, RES = cfg["radius"], cfg["P"], cfg["res"]
    N_POINTS = cfg["n_points"]

    g = mb.get_3d_grid_gmm([RES] * 3, cfg["variance"])
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    S, G = len(RADIUS), gmm.G
    clouds_host = [make_cloud(cfg, i) for i in range(max(N_CLOUDS, n_gpus))]
    clouds_dev = [torch.from_numpy(c).to(dev) for c in clouds_host]
    bounds = mb.dist.shard_bounds(N_POINTS, n_gpus)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    q_shard = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    per_cloud = hi - lo
    rows = per_cloud * n_gpus                                   # query points this rank computes per step
    feats = torch.empty((rows, RES, RES, RES, 20 * S), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    sptr = ctypes.c_void_p(stream.cuda_stream)
    dptr = ctypes.POINTER(ctypes.c_double)
    stat_events = []

    # Software pipeline across steps.  Half 1 of the NEXT step (the index builds of its clouds, the one small host read of
    # the path -- the bounding boxes, read after ONE wait for all builds; the radii are the reference's own numpy
    # expression on them -- and the ball-query launches, which are latency / barrier bound) runs on a high-priority side
    # stream while half 2 of the CURRENT step -- ONE statistics launch over the slabs of all its clouds -- runs on the
    # main stream; patches are double buffered.  Every step still performs one index build and one ball query per cloud
    # and the statistics of all its query points inside the timed region.  MUPS_BENCH_PIPELINE=0 serialises them.
    pipelined = os.environ.get("MUPS_BENCH_PIPELINE", "1") != "0"
    side = torch.cuda.Stream(dev, priority=-1) if pipelined else stream
    side_ptr = ctypes.c_void_p(side.cuda_stream)
    nbuf = 2 if pipelined else 1
    patches2 = [torch.empty((rows, S * P, 3), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    n_eff2 = [torch.empty((rows, S), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    total2 = [torch.empty((rows, S), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    stats_done = [None, None]

    def half1(i, slot):
        """Index builds + ball queries of step i's clouds into patch buffer `slot`; returns the event that follows them."""
        with torch.cuda.stream(side):
            indices = [mb.PointIndex(clouds_dev[(i * n_gpus + c) % len(clouds_dev)], cell_frac=max(RADIUS)) for c in range(n_gpus)]
        radii = [np.ascontiguousarray(ix.absolute_radii(RADIUS), dtype=np.float64) for ix in indices]     # waits once
        if stats_done[slot] is not None:
            side.wait_event(stats_done[slot])            # the statistics launch that last read this patch buffer
        for c, (index, r) in enumerate(zip(indices, radii)):
            sl = slice(c * per_cloud, (c + 1) * per_cloud)
            _lib.check(L.mups_ball_query(index.handle, ctypes.c_void_p(q_shard.data_ptr()), per_cloud, r.ctypes.data_as(dptr),
                                         S, P, SEED, None, ctypes.c_void_p(total2[slot][sl].data_ptr()),
                                         ctypes.c_void_p(patches2[slot][sl].data_ptr()),
                                         ctypes.c_void_p(n_eff2[slot][sl].data_ptr()), side_ptr))
        ev = torch.cuda.Event()
        ev.record(side)
        return indices, ev

    prefetched = {"next": half1(0, 0) if pipelined else None}

    def step(i, timed):
        """One pass of the hot path: n_gpus clouds, this rank's query shard of each."""
        slot = (i & 1) if pipelined else 0
        indices, ev = prefetched["next"] if pipelined else half1(i, 0)
        stream.wait_event(ev)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.mups_3dmfv(gmm.handle, ctypes.c_void_p(patches2[slot].data_ptr()), ctypes.c_void_p(n_eff2[slot].data_ptr()),
                                rows, S, P, _lib.FLAG_MASKED, ctypes.c_void_p(feats.data_ptr()), sptr))
        e1.record(stream)
        if timed:
            stat_events.append((e0, e1))
        done = torch.cuda.Event()
        done.record(stream)
        stats_done[slot] = done
        if pipelined:
            prefetched["next"] = half1(i + 1, slot ^ 1)

    barrier = make_barrier(world)
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
    ms_total = max_over_ranks(t0.elapsed_time(t1), dev, world)
    value = rows * n_gpus * args.steps / (ms_total * 1e-3)
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # the dominant kernel: statistics.  Algorithmic work of the last step's launch on this rank.
    last_slot = ((args.warmup + args.steps - 1) & 1) if pipelined else 0
    ne = n_eff2[last_slot].cpu().numpy()
    pairs_step = unmasked_pairs(ne, P, G)
    stat_ms = float(np.mean([a.elapsed_time(b) for a, b in stat_events]))                 # per step (one launch)
    roofline = stats_roofline(pairs_step, stat_ms, float(rows) * G * 20 * S * 4, clocks, peaks)
    roofline["peak_measured_ffma"] = fp32_peak_measured
    roofline["peak_measured_mufu_ex2"] = mufu_peak_measured
    roofline["pairs_per_query"] = pairs_step / rows
    roofline["kernel_share_of_step"] = stat_ms / (ms_total / args.steps)
    skip = set(os.environ.get("MUPS_BENCH_SKIP", "").split(","))      # profiling runs only: "e2e,cpu,aux"
    torch.cuda.synchronize()

    if "aux" not in skip:
        # ---- the general kernel: the one that executes the 46-op / 1-exp algorithmic count ---------------------------
        nq_gen = min(4096, rows)
        pg, ng = patches2[last_slot][:nq_gen], n_eff2[last_slot][:nq_gen]
        og = feats[:nq_gen]
        flags = _lib.FLAG_MASKED | _lib.FLAG_NO_FASTPATH

        def gen_launch():
            _lib.check(L.mups_3dmfv(gmm.handle, ctypes.c_void_p(pg.data_ptr()), ctypes.c_void_p(ng.data_ptr()), nq_gen, S, P,
                                    flags, ctypes.c_void_p(og.data_ptr()), sptr))
        gen_launch()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(3):
            gen_launch()
        g1.record(stream)
        torch.cuda.synchronize()
        gen_ms = g0.elapsed_time(g1) / 3
        gen_pairs = unmasked_pairs(ne[:nq_gen], P, G)
        gen_tops = OPS_PER_PAIR * gen_pairs / (gen_ms * 1e-3) / 1e12
        roofline["general_kernel"] = {"kernel": "stats_general_kernel (MUPS_FLAG_NO_FASTPATH), %d queries of the same cloud" % nq_gen,
                                      "ms_per_launch": gen_ms, "T_pairs_per_s": gen_pairs / (gen_ms * 1e-3) / 1e12,
                                      "achieved": gen_tops, "peak": roofline["peak"], "unit": "TFLOP/s", "frac": gen_tops / roofline["peak"],
                                      "mufu_exp_per_s": 2 * gen_pairs / (gen_ms * 1e-3),
                                      "note": "executes the algorithmic count (46 op-slots per pair; 2 exps per pair: one for the "
                                              "posterior's normaliser, one in the reduction pass)"}

        # ---- the second kernel of the step, timed alone (inside the loop it runs on the side stream under the statistics
        # kernel): algorithmic bytes = 16 B per true neighbour visited + the patches and counts written; the
        # 2.9 MB index of a 100 k-point cloud is L2-resident, so the L2 (lts) figure is the relevant one, not HBM
        bq_index = mb.PointIndex(clouds_dev[0], cell_frac=max(RADIUS))
        bq_radii = np.ascontiguousarray(bq_index.absolute_radii(RADIUS), dtype=np.float64)
        pb, nb, tb = patches2[0][:per_cloud], n_eff2[0][:per_cloud], total2[0][:per_cloud]

        def bq_launch():
            _lib.check(L.mups_ball_query(bq_index.handle, ctypes.c_void_p(q_shard.data_ptr()), per_cloud,
                                         bq_radii.ctypes.data_as(dptr), S, P, SEED, None, ctypes.c_void_p(tb.data_ptr()),
                                         ctypes.c_void_p(pb.data_ptr()), ctypes.c_void_p(nb.data_ptr()), sptr))
        for _ in range(2):
            bq_launch()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record(stream)
        for _ in range(5):
            bq_launch()
        b1.record(stream)
        torch.cuda.synchronize()
        bq_ms = b0.elapsed_time(b1) / 5
        nbrs = float(tb.sum().item())
        bq_bytes = 16.0 * nbrs + per_cloud * (12.0 * S * P + 8.0 * S)
        roofline["ball_query"] = {"kernel": "ball_query_kernel (K3+K4, flat cell scan), timed alone", "ms_per_launch": bq_ms,
                                  "queries_per_launch": per_cloud, "mean_neighbours_per_query": nbrs / per_cloud,
                                  "algorithmic_GBps": bq_bytes / (bq_ms * 1e-3) / 1e9, "hbm_peak_GBps": hbm_peak,
                                  "frac_of_hbm_peak": bq_bytes / (bq_ms * 1e-3) / 1e9 / hbm_peak,
                                  "note": "index is L2-resident at this cloud size; latency/barrier bound (profiles/r02_ball_query.md)"}
        lts = load_json("ball_query_traffic.json")
        if lts:
            roofline["ball_query"].update({"lts_bytes_per_query": lts.get("lts_bytes_per_query"),
                                           "lts_GBps": lts.get("lts_bytes_per_query", 0) * per_cloud / (bq_ms * 1e-3) / 1e9,
                                           "dram_bytes_per_query": lts.get("dram_bytes_per_query"), "traffic_source": lts.get("source")})
        del bq_index

        # ---- index build (K1 + K2) on a 10 M-point cloud: the size at which it is an HBM kernel -------
        if rank == 0 and n_gpus == 1:
            big = torch.from_numpy(make_cloud(CONFIGS["c5"], 2)).to(dev)
            n_big = big.shape[0]
            for _ in range(2):
                ix = mb.PointIndex(big, cell_frac=max(RADIUS))
                del ix
            i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            i0.record(stream)
            for _ in range(5):
                ix = mb.PointIndex(big, cell_frac=max(RADIUS))
                del ix
            i1.record(stream)
            torch.cuda.synchronize()
            ib_ms = i0.elapsed_time(i1) / 5
            alg = 68.0 * n_big                                   # 12 read + 16 + 4 (idx) written, code / rank / position 36
            ib = {"kernels": "bbox, cell_code, 3-kernel scan, scatter (K1 + K2)", "points": n_big, "ms_per_build": ib_ms,
                  "algorithmic_bytes_per_point": 68, "algorithmic_GBps": alg / (ib_ms * 1e-3) / 1e9, "hbm_peak_GBps": hbm_peak,
                  "frac_of_hbm_peak": alg / (ib_ms * 1e-3) / 1e9 / hbm_peak,
                  "cell_table_entries": "8^bits Morton codes, memset + scanned every build (bits = 8 for the fine grid of a 10 M-point cloud: 67 MB)"}
            tr = load_json("index_build_traffic.json")
            if tr:
                ib.update({"dram_bytes_per_build_ncu": tr.get("dram_bytes"), "lts_bytes_per_build_ncu": tr.get("lts_bytes"),
                           "dram_GBps_at_live_time": tr.get("dram_bytes", 0) / (ib_ms * 1e-3) / 1e9, "traffic_source": tr.get("source")})
            roofline["index_build"] = ib
            del big
    prof = load_json("stats_kernel_traffic.json")
    if prof:
        roofline["traffic"] = prof["dram_bytes_per_query"] * rows
        roofline["traffic_source"] = prof.get("source")

    # ---- end to end through the public API with host buffers ------------------------------------------
    # chunk: query points per pipeline stage.  Small enough that the first device->host copy starts ~1.5 ms after the
    # call (each cloud's call fills and drains the pipeline), large enough (336 MB per copy) for full PCIe rate.
    chunk = int(os.environ.get("MUPS_BENCH_CHUNK", "2048"))
    if world > 1:
        mb.dist.bind_to_gpu_numa_node(local_rank)     # pinned staging buffers on the GPU's NUMA node (no-op where sysfs has none)
    del patches2[1:], n_eff2[1:], total2[1:]
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
    e2e_s = max_over_ranks(time.perf_counter() - w0, dev, world)
    barrier()
    e2e = {"value": n_done * n_gpus / e2e_s, "unit": UNIT, "steps": e2e_steps,
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
        dc_s = max_over_ranks(time.perf_counter() - w0, dev, world)
        barrier()
        e2e["device_consumer"] = {"value": n_dev * n_gpus / dc_s, "unit": UNIT,
                                  "h2d_bytes_per_step": pipe_dev.h2d_bytes // e2e_steps, "d2h_bytes_per_step": 8,
                                  "chunk_queries": pipe_dev.chunk,
                                  "checksum_finite": bool(np.isfinite(checksum)),
                                  "api": "MuPSPipeline.features_to_consumer (host cloud in, MuPS reduced on the GPU)"}

    # ---- the reference's user-visible path (test_n_est_w_experts.py:129-197): host cloud in -> MuPS -> Mixture-of-Experts ->
    # normals out, with the consumer on the tensor cores (moe_engine.TensorCoreExperts: hand-written tcgen05 conv3d, random-init
    # network) fed chunk by chunk through MuPSPipeline.features_to_consumer, so MuPS never leaves the device
    if "e2e" not in skip and "normals" not in skip and RES == 8:        # every rank: its own cloud, its own consumer (data parallel)
        from nesti_net_b200.experts_net import ExpertsNormalEstimator
        from nesti_net_b200.moe_engine import TensorCoreExperts
        torch.manual_seed(1234)
        tc = TensorCoreExperts(ExpertsNormalEstimator(n_rads=S, n_gaussians=G, n_experts=7).eval().to(dev))
        from nesti_net_b200.inference import CloudNormalEstimator
        nq_n = int(os.environ.get("MUPS_BENCH_NORMALS_QUERIES", "16384"))
        est = CloudNormalEstimator(tc, gmm, RADIUS, P, seed=SEED, chunk=2048)
        pipe_n = est.pipe
        qn_host = (torch.arange(nq_n, dtype=torch.int64) * (N_POINTS // nq_n)).pin_memory()
        est(hosts[0], qn_host[:2048])                                   # warm-up (allocations, tensor maps)
        torch.cuda.synchronize()
        barrier()
        pipe_n.h2d_bytes = 0
        w0 = time.perf_counter()
        nrm_host, exp_host, prob_host = est(hosts[(1 + rank) % len(hosts)], qn_host)     # synchronises before it returns
        nt = max_over_ranks(time.perf_counter() - w0, dev, world)
        barrier()
        n_n = int(nrm_host.shape[0])
        e2e["normals"] = {"value": n_n * n_gpus / nt, "unit": UNIT, "queries": n_n * n_gpus, "h2d_bytes_per_step": pipe_n.h2d_bytes,
                          "d2h_bytes_per_step": n_n * (12 + 8 + 4 * int(prob_host.shape[1])), "chunk_queries": pipe_n.chunk,
                          "finite": bool(np.isfinite(nrm_host).all()),
                          "api": "inference.CloudNormalEstimator: host cloud in -> MuPSPipeline.features_to_consumer -> "
                                 "moe_engine.TensorCoreExperts.predict (tcgen05 conv3d, bf16 x bf16 -> fp32) -> normals, experts, "
                                 "probabilities to pinned host memory; random-init 7-expert network; every rank its own cloud and consumer",
                          "cudnn_strict_fp32_queries_per_s": 774, "cudnn_source": "profiles/r02_moe.jsonl"}
        del tc, pipe_n, est
        if "x3" not in skip:
            # the same path with the consumer in bf16x3 mode (hi / lo bf16 pairs: a_hi w_hi + a_lo w_hi + a_hi w_lo in fp32 on the
            # same tcgen05 kernels): normals within ~1e-3 degrees of the fp32 network (tests/test_gpu.py) at 3 x the tensor work
            torch.manual_seed(1234)
            tc3 = TensorCoreExperts(ExpertsNormalEstimator(n_rads=S, n_gaussians=G, n_experts=7).eval().to(dev), precision="bf16x3")
            est3 = CloudNormalEstimator(tc3, gmm, RADIUS, P, seed=SEED, chunk=1024)
            nq3 = max(1024, nq_n // 2)
            est3(hosts[0], qn_host[:1024])
            torch.cuda.synchronize()
            barrier()
            w0 = time.perf_counter()
            nrm3, exp3, _ = est3(hosts[(1 + rank) % len(hosts)], qn_host[:nq3])
            nt3 = max_over_ranks(time.perf_counter() - w0, dev, world)
            barrier()
            same3 = exp3 == exp_host[:nq3]
            cosang = np.clip(np.abs((nrm3[same3] * nrm_host[:nq3][same3]).sum(1)) /
                             np.maximum(np.linalg.norm(nrm3[same3], axis=1) * np.linalg.norm(nrm_host[:nq3][same3], axis=1), 1e-30), 0.0, 1.0)
            e2e["normals"]["bf16x3"] = {
                "value": int(nrm3.shape[0]) * n_gpus / nt3, "unit": UNIT, "queries": int(nrm3.shape[0]) * n_gpus, "chunk_queries": est3.pipe.chunk,
                "finite": bool(np.isfinite(nrm3).all()),
                "bf16_vs_bf16x3_rms_deg": float(np.degrees(np.sqrt(np.mean(np.arccos(cosang) ** 2)))) if same3.any() else None,
                "same_expert_as_bf16": "%d/%d" % (int(same3.sum()), int(nq3)),
                "api": "the same call with moe_engine.TensorCoreExperts(precision='bf16x3') (csrc/moe_split.cu)"}
            del tc3, est3

    cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": "measured at N=1 only (see the N=1 line / --impl reference)"}
    if rank == 0 and n_gpus == 1:
        # ---- CPU baseline on a bounded sample (the oracle is the thing timed here, never the product) ----
        cpu_baseline = cpu_baseline_entry(cfg, clouds_host[0], 1024 if "cpu" not in skip else 16, "cpu" not in skip)
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config(cfg, n_gpus), "roofline": roofline,
               "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# own arm: strong-scaling configs on one dense cloud (c4, c5)
# ------------------------------------------------------------------------------------------------

def run_dense(args, cfg):
    import torch
    import torch.distributed as dist
    import nesti_net_b200 as mb
    from nesti_net_b200 import _lib

    world, rank, local_rank, dev = setup_dist()
    n_gpus = world
    _lib.load()
    RADIUS, N_POINTS = cfg["radius"], cfg["n_points"]
    S = len(RADIUS)
    stream = torch.cuda.current_stream(dev)
    barrier = make_barrier(world)
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    pts_host = make_cloud(cfg, 0)
    xyz = torch.from_numpy(pts_host).to(dev)
    nq = cfg["queries"] or N_POINTS
    q_all = (np.arange(nq, dtype=np.int64) * (N_POINTS // nq)) if nq < N_POINTS else np.arange(N_POINTS, dtype=np.int64)
    q_all = q_all[np.argsort(pts_host[q_all, 2], kind="stable")]          # scanner-sweep order: contiguous ranges differ in density

    # ---- partition: contiguous ranges balanced by estimated work, computed once per cloud -----------
    index = mb.PointIndex(xyz, cell_frac=max(RADIUS))
    radii = index.absolute_radii(RADIUS)
    torch.cuda.synchronize()
    p0 = time.perf_counter()
    stride = 64
    _, _, tot = index.ball_query(torch.from_numpy(q_all[::stride].copy()).to(dev), radii, cfg["P"], seed=SEED, return_patches=False)
    tot = tot.double().cpu().numpy()
    # cost model: the ball query visits every neighbour of the larger radii, the statistics kernel min(P, n) + 1 points per scale
    est = 0.004 * tot.sum(1) + np.minimum(tot, cfg["P"]).sum(1) + 200.0
    weights = np.repeat(est, stride)[:nq]
    bounds = mb.dist.shard_bounds(nq, n_gpus, weights)
    partition_ms = 1e3 * (time.perf_counter() - p0)
    even = mb.dist.shard_bounds(nq, n_gpus)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    q_dev = torch.from_numpy(q_all[lo:hi].copy()).to(dev)
    del index

    def run_variant(res, P, q, steps, warmup, timed_clocks):
        """`steps` passes over the query shard `q` (index build + chunked ball query / statistics into a reused feature
        buffer: at 164 KB per query the full tensor of these configs is never stored; the e2e leg adds a device consumer); returns timing and algorithmic work of this rank."""
        g = mb.get_3d_grid_gmm([res] * 3, grid_variance(res))
        gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
        G = gmm.G
        chunk = 16384 if res <= 8 else 2048
        B = int(q.shape[0])
        feats = torch.empty((min(chunk, max(B, 1)), res, res, res, 20 * S), dtype=torch.float32, device=dev)
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        ev = {"bq": [], "st": [], "ib": []}
        work = torch.zeros(2, dtype=torch.float64, device=dev)       # neighbours visited, unmasked patch slots (last step)

        def one_step(timed):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            ix = mb.PointIndex(xyz, cell_frac=max(RADIUS))
            a1.record(stream)
            r = ix.absolute_radii(RADIUS)
            for c0 in range(0, B, chunk):
                m = min(chunk, B - c0)
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record(stream)
                patches, n_eff, total = ix.ball_query(q[c0:c0 + m], r, P, seed=SEED)
                e1.record(stream)
                mb.stats_3dmfv(patches, n_eff, gmm, S, out=feats[:m])
                e2.record(stream)
                if timed:
                    ev["bq"].append((e0, e1))
                    ev["st"].append((e1, e2))
                    if timed == "last":        # algorithmic work of the step, accumulated on the device (no host sync here)
                        work[0] += total.sum(dtype=torch.float64)
                        work[1] += torch.where(n_eff >= P - 1, P, n_eff + 1).sum(dtype=torch.float64)
            if timed:
                ev["ib"].append((a0, a1))
        for _ in range(warmup):
            one_step(False)
        barrier()
        if timed_clocks is not None:
            timed_clocks.start()
        launches0 = _lib.launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for i in range(steps):
            one_step("last" if i == steps - 1 else True)
        t1.record(stream)
        barrier()
        launches = _lib.launch_count() - launches0
        ms_rank = t0.elapsed_time(t1)
        ms = max_over_ranks(ms_rank, dev, world)
        s = lambda k: float(sum(a.elapsed_time(b) for a, b in ev[k])) / steps
        acc.add_(feats[:min(chunk, max(B, 1))].sum(dtype=torch.float64))      # the last chunk's features, after the timed region
        wk = work.cpu().numpy()
        return {"ms_per_step": ms / steps, "ms_per_step_this_rank": ms_rank / steps, "bq_ms": s("bq"), "st_ms": s("st"),
                "ib_ms": s("ib"), "nbrs": float(wk[0]), "pairs": float(wk[1]) * G, "launches": int(launches), "G": G, "queries": B,
                "checksum_finite": bool(np.isfinite(float(acc.item())))}

    sampler = ClockSampler(local_rank) if rank == 0 else None
    main = run_variant(cfg["res"], cfg["P"], q_dev, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if rank == 0 else None
    value = nq / (main["ms_per_step"] * 1e-3)
    # per-rank times of the main line (the limiter of strong scaling is the slowest rank)
    import torch.distributed as tdist
    mine = torch.tensor([main["ms_per_step_this_rank"], float(hi - lo), main["bq_ms"], main["st_ms"]], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        tdist.all_gather(allr, mine)
    else:
        allr = [mine]
    allr = np.array([t.cpu().numpy() for t in allr])

    P, G = cfg["P"], main["G"]
    bq_bytes = 16.0 * main["nbrs"] + (hi - lo) * (12.0 * S * P + 8.0 * S)
    roofline = {"kernel": "ball_query_hier_kernel (K3+K4: octree descent over the Morton cells, whole-cell acceptance, one-pass key "
                          "threshold), summed over the chunks of one step on rank 0",
                "bound": "hbm", "achieved": bq_bytes / (main["bq_ms"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bq_bytes / (main["bq_ms"] * 1e-3) / 1e9 / hbm_peak,
                "frac_is": "ALGORITHMIC bytes (16 B per true neighbour + the patches and counts written, SURVEY.md 8d) / time / measured "
                           "HBM copy peak.  Cells wholly inside a ball are accepted from cell_start and only their 4-byte indices "
                           "are read, so the executed traffic is below the algorithmic count (profiles/r02_ball_query_dense.md)",
                "peak_is": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                "kernel_ms_per_step": main["bq_ms"], "queries_per_s_alone": (hi - lo) / (main["bq_ms"] * 1e-3),
                "mean_neighbours_per_query": main["nbrs"] / max(1, hi - lo),
                "kernel_share_of_step": main["bq_ms"] / main["ms_per_step_this_rank"], "traffic": None}
    tr = load_json("ball_query_dense_traffic.json")
    if tr and cfg["n_points"] >= 10000000:
        roofline["traffic"] = tr["dram_bytes_per_query"] * (hi - lo)
        roofline["lts_bytes_per_query"] = tr.get("lts_bytes_per_query")
        roofline["traffic_source"] = tr.get("source")
    roofline["stats"] = stats_roofline(main["pairs"], main["st_ms"], float(hi - lo) * G * 20 * S * 4, clocks, peaks,
                                       n_launches=max(1, (hi - lo + 16383) // 16384))
    roofline["stats"]["queries_per_s_alone"] = (hi - lo) / (main["st_ms"] * 1e-3)
    roofline["index_build"] = {"ms_per_build": main["ib_ms"], "points": N_POINTS,
                               "algorithmic_GBps": 68.0 * N_POINTS / (main["ib_ms"] * 1e-3) / 1e9,
                               "frac_of_hbm_peak": 68.0 * N_POINTS / (main["ib_ms"] * 1e-3) / 1e9 / hbm_peak}
    sharding = {"queries_per_rank": [int(x) for x in allr[:, 1]], "ms_per_rank": [round(float(x), 2) for x in allr[:, 0]],
                "imbalance_max_over_mean": float(allr[:, 0].max() / allr[:, 0].mean()),
                "even_split_would_be": [int(even[r + 1] - even[r]) for r in range(n_gpus)],
                "partition_ms_once_per_cloud": partition_ms,
                "limiter": "no collective: the step ends when the slowest rank ends; residual imbalance of the work estimate + "
                           "one index build per rank per step (not sharded)"}

    skip = set(os.environ.get("MUPS_BENCH_SKIP", "").split(","))
    variants = []
    if "variants" not in skip:
        nv = 131072
        sub = q_dev[:: max(1, int(q_dev.shape[0]) * n_gpus // nv)].contiguous()
        for v in cfg["variants"]:
            r = run_variant(v["res"], v["P"], sub, 2, 1, None)
            tot_q = max_over_ranks(float(sub.shape[0]), dev, world) * n_gpus
            variants.append({"grid": "%d^3" % v["res"], "points_per_patch": v["P"], "queries_per_step": int(tot_q),
                             "value": tot_q / (r["ms_per_step"] * 1e-3), "unit": UNIT,
                             "ball_query_q_per_s_per_gpu": int(sub.shape[0]) / (r["bq_ms"] * 1e-3),
                             "stats_q_per_s_per_gpu": int(sub.shape[0]) / (r["st_ms"] * 1e-3),
                             "stats_T_pairs_per_s": r["pairs"] / (r["st_ms"] * 1e-3) / 1e12})

    # ---- end to end: host cloud and host query list in, MuPS consumed on the device chunk by chunk (the reference's
    # flow: MuPS feeds the CNN; at 164 KB per query the full tensor of these configs -- 0.17 / 0.33 TB -- is never stored)
    e2e = None
    if "e2e" not in skip:
        g = mb.get_3d_grid_gmm([cfg["res"]] * 3, cfg["variance"])
        gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
        pipe = mb.MuPSPipeline(gmm, RADIUS, P, seed=SEED, chunk=16384)
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        host_pts = torch.from_numpy(pts_host).pin_memory()
        host_q = torch.from_numpy(q_all[lo:hi].copy()).pin_memory()

        def on_device(lo_, hi_, rows_):
            acc.add_(rows_.sum())
        pipe.features_to_consumer(host_pts, host_q[:16384], on_device)
        barrier()
        pipe.h2d_bytes = 0
        w0 = time.perf_counter()
        n = pipe.features_to_consumer(host_pts, host_q, on_device)
        checksum = float(acc.item())
        dt = max_over_ranks(time.perf_counter() - w0, dev, world)
        barrier()
        e2e = {"value": nq / dt, "unit": UNIT, "steps": 1, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": 8,
               "chunk_queries": pipe.chunk, "checksum_finite": bool(np.isfinite(checksum)), "queries_this_rank": int(n),
               "api": "MuPSPipeline.features_to_consumer (pinned host cloud + query list in, index build, both halves per chunk, MuPS "
                      "reduced on the GPU, 8 bytes read back)"}

    cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                    "sample": "measured at N=1 only (see the N=1 line / --impl reference)"}
    if rank == 0 and n_gpus == 1 and "cpu" not in skip:
        cpu_baseline = cpu_baseline_entry(cfg, pts_host, cfg["cpu_sample"], False)
    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": config(cfg, n_gpus), "roofline": roofline, "sharding": sharding,
               "variants": variants, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": main["launches"], "clocks": clocks}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# own arm: the optional gather for a single-rank consumer (--gather, N > 1)
# ------------------------------------------------------------------------------------------------

def run_gather(args, cfg):
    """One cloud per step, its queries sharded over the ranks, every slab delivered to rank 0: (a) by the statistics
    kernel's own stores into rank 0's memory over NVLink (dist.PeerSlabGather + MUPS_FLAG_WIDE_STORES: the kernel IS the
    collective), (b) local slabs + NCCL all-gather (dist.gather_slabs), (c) local slabs only (no gather), same data."""
    import torch
    import torch.distributed as dist
    import nesti_net_b200 as mb
    from nesti_net_b200 import _lib

    world, rank, local_rank, dev = setup_dist()
    _lib.load()
    RADIUS, P, RES, N_POINTS = cfg["radius"], cfg["P"], cfg["res"], cfg["n_points"]
    S = len(RADIUS)
    g = mb.get_3d_grid_gmm([RES] * 3, cfg["variance"])
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    nq = min(N_POINTS, int(os.environ.get("MUPS_GATHER_QUERIES", "65536")))         # 10.7 GB gathered per step at 64k
    clouds = [torch.from_numpy(make_cloud(cfg, i)).to(dev) for i in range(N_CLOUDS)]
    bounds = mb.dist.shard_bounds(nq, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    q = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)
    barrier = make_barrier(world)
    row = (RES, RES, RES, 20 * S)
    local = torch.empty((hi - lo,) + row, dtype=torch.float32, device=dev)
    gather = mb.dist.PeerSlabGather(nq, row, dst=0) if world > 1 else None

    def half1(i):
        ix = mb.PointIndex(clouds[i % N_CLOUDS], cell_frac=max(RADIUS))
        return ix.ball_query(q, ix.absolute_radii(RADIUS), P, seed=SEED)

    def mode_local(i):
        patches, n_eff, _ = half1(i)
        mb.stats_3dmfv(patches, n_eff, gmm, S, out=local)

    def mode_peer(i):
        patches, n_eff, _ = half1(i)
        mb.stats_3dmfv(patches, n_eff, gmm, S, out=gather.target(lo, hi), wide_stores=True)
        gather.finish()

    def mode_nccl(i):
        patches, n_eff, _ = half1(i)
        mb.stats_3dmfv(patches, n_eff, gmm, S, out=local)
        return mb.dist.gather_slabs(local)

    results = {}
    full = None
    modes = [("local_slabs_no_gather", mode_local)]
    if world > 1:
        modes += [("peer_store_fused", mode_peer), ("nccl_all_gather", mode_nccl)]
    for name, fn in modes:
        for i in range(args.warmup):
            fn(i)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(stream)
        for i in range(args.steps):
            out = fn(args.warmup + i)
        t1.record(stream)
        barrier()
        ms = max_over_ranks(t0.elapsed_time(t1), dev, world) / args.steps
        results[name] = {"ms_per_step": ms, "value": nq / (ms * 1e-3), "unit": UNIT}
        if name == "nccl_all_gather":
            full = out
    identical = None
    if world > 1 and rank == 0:
        identical = bool(torch.equal(gather.result(), full))           # same last cloud in both modes
    if rank == 0:
        head = results.get("peer_store_fused", results["local_slabs_no_gather"])
        emit({"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
              "data": "synthetic", "mode": "gather",
              "config": config(cfg, world, {"queries_per_step": nq, "gathered_bytes_per_step": nq * RES ** 3 * 20 * S * 4,
                                            "partitioning": "one cloud per step, %d queries sharded over %d ranks, all slabs "
                                                            "delivered to rank 0" % (nq, world)}),
              "gather": results, "gathered_tensor_bit_identical_peer_vs_nccl": identical,
              "roofline": {"bound": "nvlink", "note": "rank 0 receives (N-1)/N of the tensor: ingress bound %.1f GB at the measured "
                                                      "770 GB/s peer-copy rate = %.2f ms"
                                                      % (nq * RES ** 3 * 80.0 * S * (world - 1) / world / 1e9,
                                                         nq * RES ** 3 * 80.0 * S * (world - 1) / world / 770e9 * 1e3)}})
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--gather", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.gather:
        run_gather(args, CONFIGS["c2"])
    elif cfg["mode"] == "clouds":
        run_clouds(args, cfg)
    else:
        run_dense(args, cfg)


if __name__ == "__main__":
    main()
