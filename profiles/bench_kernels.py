#!/usr/bin/env python
"""Per-kernel timings (CUDA events, after warm-up) of the MuPS path on one B200: index build,
ball query, statistics kernel variants.  Development aid; bench.py is the contract benchmark."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200 import _lib  # noqa: E402
from nesti_net_b200 import synthetic as orc  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(os.environ.get("N", 100000))
    nq = int(os.environ.get("NQ", 32768))
    res = int(os.environ.get("RES", 8))
    P = int(os.environ.get("P", 512))
    radius = [0.01, 0.03, 0.05, 0.07]
    S = len(radius)
    pts = orc.synthetic_cloud(n, cloud_id=0)
    var = 0.0156 if res == 8 else (1.0 / res) ** 2
    g = mb.get_3d_grid_gmm([res] * 3, var)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    xyz = torch.from_numpy(pts).cuda()
    t_build = timed(lambda: mb.PointIndex(xyz, cell_frac=max(radius), cell_scale=float(os.environ['CELL_SCALE']) if os.environ.get('CELL_SCALE') else None))
    index = mb.PointIndex(xyz, cell_frac=max(radius), cell_scale=float(os.environ['CELL_SCALE']) if os.environ.get('CELL_SCALE') else None)
    radii = index.absolute_radii(radius)
    q = torch.from_numpy(np.random.RandomState(0).choice(n, nq, replace=nq > n)).cuda()
    out = {}
    for fuse in [int(v) for v in os.environ.get("FUSE", "").split(",") if v]:
        _lib.set_option("fuse_candidates", fuse)
        print(json.dumps({"fuse_candidates": fuse, "ball_query_ms": timed(lambda: index.ball_query(q, radii, P), iters=10, warm=3)}))
    _lib.set_option("fuse_candidates", 12288)
    t_query = timed(lambda: index.ball_query(q, radii, P))
    patches, n_eff, total = index.ball_query(q, radii, P)
    ne = n_eff.cpu().numpy()
    m = np.where(ne >= P - 1, P, ne + 1)
    pairs = float(m.sum()) * gmm.G
    feats = torch.empty((nq, res, res, res, 20 * S), dtype=torch.float32, device="cuda")
    print(json.dumps({"n": n, "queries": nq, "index_build_ms": t_build, "ball_query_ms": t_query,
                      "ball_query_Mq_per_s": nq / t_query / 1e3, "mean_total": total.float().mean(0).tolist(),
                      "mean_unmasked_points_per_query": float(m.sum()) / nq, "pairs_per_query": pairs / nq}))
    reps = int(os.environ.get("REPS", 2))
    for rep in range(reps):
        for name, variant, fast in (("general", 0, False), ("sep_default_hybrid_pairsum_serialstage", 0, True),
                                    ("sep_round1_packed_shufflestage", 1, True), ("sep_allscalar_pairsum_serialstage", 2, True),
                                    ("sep_two_items_per_cta_bulk_prefetch", 3, True)):
            if rep and name == "general":
                continue
            if os.environ.get("ONLY") and str(variant) not in os.environ["ONLY"].split(","):
                continue
            _lib.set_option("stats_variant", variant)
            t = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=feats, fastpath=fast), iters=10, warm=3)
            print(json.dumps({"stats_kernel": name, "rep": rep, "ms": round(t, 3), "Tpairs_per_s": round(pairs / t / 1e9, 4),
                              "Mq_per_s": round(nq / t / 1e3, 4), "algorithmic_Tflop_per_s": round(46 * pairs / t / 1e9, 2),
                              "out_GB_per_s": round(feats.numel() * 4 / t / 1e6, 1)}))
    _lib.set_option("stats_variant", 0)


if __name__ == "__main__":
    main()
