#!/usr/bin/env python
"""The other BASELINE.json configs (parity-test cases, not bench lines) on one B200:
  C1  100k PCPNet-shape cloud, 3 scales (0.01/0.03/0.07), P=512, 8^3
  C4  2M-point scan-shape cloud (non-uniform density), 4 scales, P=512, 8^3
  C5  10M-point cloud, grid 8^3 / 16^3, P = 256 / 512 / 1024, 4 scales
For each: index build time, ball-query and statistics throughput on a strided query sample
(CUDA events, after warm-up), and an oracle spot check (cKDTree neighbour counts bit-exact, patches
bit-exact, features within tolerance) on a few of the sampled queries.  One JSON object per line.

    python profiles/bench_configs.py [C1 C4 C5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from oracle import c_oracle  # noqa: E402
from oracle import mups_oracle as orc  # noqa: E402

SEED = 3627473


def timed(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def run(name, pts, radius, P, res, n_queries, n_check, kdtree=None):
    n = len(pts)
    var = 0.0156 if res == 8 else (1.0 / res) ** 2
    g = mb.get_3d_grid_gmm([res] * 3, var)
    w, mu, sg = np.asarray(g.weights_, np.float32), np.asarray(g.means_, np.float32), np.sqrt(g.covariances_).astype(np.float32)
    gmm = mb.gmm_handle(w, mu, sg)
    S = len(radius)
    xyz = torch.from_numpy(pts).cuda()
    t_build, index = timed(lambda: mb.PointIndex(xyz, cell_frac=max(radius), cell_scale=float(os.environ['CELL_SCALE']) if os.environ.get('CELL_SCALE') else None))
    radii = index.absolute_radii(radius)
    q = (np.arange(n_queries, dtype=np.int64) * (n // n_queries) + 17) % n
    qd = torch.from_numpy(q).cuda()
    t_query, (patches, n_eff, total) = timed(lambda: index.ball_query(qd, radii, P, seed=SEED), iters=2)
    feats = torch.empty((n_queries, res, res, res, 20 * S), dtype=torch.float32, device="cuda")
    t_stats, _ = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=feats), iters=2)
    ne = n_eff.cpu().numpy()
    m = np.where(ne >= P - 1, P, ne + 1)
    pairs = float(m.sum()) * gmm.G
    rec = {"config": name, "points": n, "scales": radius, "P": P, "grid": res, "queries_timed": n_queries,
           "index_build_ms": round(t_build, 3), "index_build_GBps_64B_per_point": round(64.0 * n / t_build / 1e6, 1),
           "ball_query_ms": round(t_query, 3), "ball_query_kq_per_s": round(n_queries / t_query, 2),
           "mean_neighbours": [round(float(x), 1) for x in total.float().mean(0).tolist()],
           "max_neighbours": int(total.max().item()),
           "neighbour_visit_GBps_16B": round(16.0 * float(total.sum().item()) / t_query / 1e6, 1),
           "stats_ms": round(t_stats, 3), "stats_Tpairs_per_s": round(pairs / t_stats / 1e9, 4),
           "stats_algorithmic_Tflops": round(46 * pairs / t_stats / 1e9, 2),
           "mups_kq_per_s": round(n_queries / (t_query + t_stats), 2)}
    if n_check:
        t0 = time.time()
        kd = kdtree if kdtree is not None else orc.build_kdtree(pts)
        sub = np.linspace(0, n_queries - 1, n_check).astype(np.int64)
        o_patches, o_neff, o_total = orc.gather_patches(pts, q[sub], radius, P, seed=SEED, kdtree=kd)
        rec["check_counts_exact"] = bool(np.array_equal(total.cpu().numpy()[sub], o_total))
        rec["check_patches_bit_exact"] = bool(np.array_equal(patches.cpu().numpy()[sub].view(np.uint32), o_patches.view(np.uint32)))
        ref = c_oracle.mups(o_patches, o_neff, w, mu, sg, S)
        got = feats.cpu().numpy()[sub]
        err = np.abs(got - ref)
        rec["check_features_max_err"] = float(err.max())
        rec["check_features_frac_outside_tol"] = float((err > 1e-6 + 1e-5 * np.abs(ref)).mean())
        rec["check_queries"] = int(n_check)
        rec["oracle_seconds"] = round(time.time() - t0, 1)
    print(json.dumps(rec), flush=True)
    return rec


def main():
    which = sys.argv[1:] or ["C1", "C4", "C5"]
    if "C1" in which:
        run("C1", orc.synthetic_cloud(100000, cloud_id=0), [0.01, 0.03, 0.07], 512, 8, 32768, 32)
    if "C4" in which:
        pts = orc.synthetic_cloud(2000000, cloud_id=1, kind="scan")
        run("C4", pts, [0.01, 0.03, 0.05, 0.07], 512, 8, 4096, 8)
    if "C5" in which:
        pts = orc.synthetic_cloud(10000000, cloud_id=2)
        kd = orc.build_kdtree(pts)
        for res, P in ((8, 512), (8, 256), (8, 1024), (16, 512)):
            run("C5 res=%d P=%d" % (res, P), pts, [0.01, 0.03, 0.05, 0.07], P, res, 1024, 4, kdtree=kd)


if __name__ == "__main__":
    main()
