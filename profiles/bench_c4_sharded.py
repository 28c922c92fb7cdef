#!/usr/bin/env python
"""BASELINE.json configs[3]: dense scan-shape cloud (2 M points, density spread > 20x), MuPS at 4 scales, the query
list sharded across the GPUs of one box: cloud and index replicated, contiguous query ranges, per-rank
slabs, no collective on the data path.  Compares an even split with a split balanced by estimated work (neighbour
count at the largest radius of every 64th query, one cheap extra ball-query launch).  torchrun, one rank per GPU;
rank 0 prints one JSON line per mode.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 profiles/bench_c4_sharded.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from oracle import mups_oracle as orc  # noqa: E402   (synthetic cloud only)

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]
P = 512
CHUNK = 8192


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(os.environ.get("N", 2000000))
    nq = int(os.environ.get("NQ", 262144))
    pts = orc.synthetic_cloud(n, cloud_id=1, kind="scan")
    g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    S = len(RADIUS)
    index = mb.PointIndex(torch.from_numpy(pts).to(dev), cell_frac=max(RADIUS))
    radii = index.absolute_radii(RADIUS)
    # the query list in the order of the cloud file (the scan cloud is sorted by nothing: density varies along it
    # only through the sampling), sorted by z so that contiguous ranges differ in density like a scanner sweep
    q_all = np.sort(np.random.RandomState(3).choice(n, nq, replace=False))
    q_all = q_all[np.argsort(pts[q_all, 2], kind="stable")]
    feats = torch.empty((CHUNK, 8, 8, 8, 20 * S), dtype=torch.float32, device=dev)

    # work estimate: neighbours at the largest radius of every 64th query (+ a constant for the statistics kernel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, tot = index.ball_query(torch.from_numpy(q_all[::64]).to(dev), radii, P, seed=SEED, return_patches=False)
    e1.record()
    torch.cuda.synchronize()
    est_ms = e0.elapsed_time(e1)
    sample = tot[:, -1].double().cpu().numpy()
    weights = np.repeat(sample, 64)[:nq] + 35000.0

    def run(bounds):
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        q = torch.from_numpy(q_all[lo:hi]).to(dev)
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for c in range(0, hi - lo, CHUNK):
            m = min(CHUNK, hi - lo - c)
            mb.mups_features(index, gmm, q[c:c + m], radii, P, seed=SEED, out=feats[:m])
            acc.add_(feats[:m].sum(dtype=torch.float64))
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b), float(hi - lo), float(acc.item())], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(out, t)
        else:
            out = [t]
        return np.array([o.cpu().numpy() for o in out])

    run(mb.dist.shard_bounds(min(nq, 4096 * world), world))      # warm-up
    results = {}
    for mode, bounds in (("even", mb.dist.shard_bounds(nq, world)), ("balanced", mb.dist.shard_bounds(nq, world, weights))):
        r = run(bounds)
        results[mode] = r
        if rank == 0:
            ms = r[:, 0]
            print(json.dumps({"config": "C4 sharded", "mode": mode, "n_gpus": world, "points": n, "queries": nq,
                              "queries_per_rank": [int(x) for x in r[:, 1]], "ms_per_rank": [round(float(x), 1) for x in ms],
                              "ms_max": round(float(ms.max()), 1), "imbalance_max_over_mean": round(float(ms.max() / ms.mean()), 3),
                              "kq_per_s": round(nq / ms.max(), 2), "work_estimate_ms": round(est_ms, 2),
                              "checksum": float(r[:, 2].sum())}), flush=True)
    if rank == 0:
        a, b = results["even"][:, 2].sum(), results["balanced"][:, 2].sum()
        print(json.dumps({"checksums_agree": bool(abs(a - b) <= 1e-9 * abs(a))}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
