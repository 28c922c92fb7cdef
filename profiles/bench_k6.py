#!/usr/bin/env python
"""K6 (selection hand-off: mups_ball_query_select + mups_3dmfv_selected, patches never in HBM) against the two-launch
path through the patch tensor (mups_ball_query + mups_3dmfv), same cloud, same queries, CUDA events after a warm-up;
the features of the two paths are compared bit for bit.  One JSON object per line.

    python profiles/bench_k6.py [c2 c5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]
P, S = 512, 4


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


def main():
    which = sys.argv[1:] or ["c2", "c5"]
    g = mb.get_3d_grid_gmm([8] * 3, 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    for name in which:
        if name == "c2":
            pts, nq = synthetic_cloud(100000, cloud_id=0), 100000
            q = np.arange(nq, dtype=np.int64)
        else:
            pts, nq = synthetic_cloud(10000000, cloud_id=2), 32768
            q = np.random.RandomState(5).choice(len(pts), nq, replace=False).astype(np.int64)
        xyz = torch.from_numpy(pts).cuda()
        qd = torch.from_numpy(q).cuda()
        index = mb.PointIndex(xyz, cell_frac=max(RADIUS))
        radii = index.absolute_radii(RADIUS)
        out = torch.empty((nq, 8, 8, 8, 20 * S), dtype=torch.float32, device="cuda")
        out2 = torch.empty_like(out)

        def two_launch():
            patches, n_eff, _ = index.ball_query(qd, radii, P, seed=SEED)
            return mb.stats_3dmfv(patches, n_eff, gmm, S, out=out)

        def k6():
            return mb.mups_features(index, gmm, qd, radii, P, seed=SEED, out=out2)

        def half1_patches():
            return index.ball_query(qd, radii, P, seed=SEED)

        def half1_select():
            return index.select(qd, radii, P, seed=SEED)
        t_two, _ = timed(two_launch)
        t_k6, _ = timed(k6)
        t_h1p, (patches, n_eff, _) = timed(half1_patches)
        t_h1s, (pos, n_eff2, _) = timed(half1_select)
        t_st, _ = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=out))
        t_sts, _ = timed(lambda: mb.stats_3dmfv_selected(index, gmm, qd, radii, pos, n_eff2, out=out2))
        print(json.dumps({"config": name, "points": len(pts), "queries": nq,
                          "two_launch_ms": round(t_two, 3), "k6_ms": round(t_k6, 3), "speedup": round(t_two / t_k6, 4),
                          "half1_with_patches_ms": round(t_h1p, 3), "half1_select_only_ms": round(t_h1s, 3),
                          "stats_from_patches_ms": round(t_st, 3), "stats_from_selection_ms": round(t_sts, 3),
                          "intermediate_bytes_per_query": {"patch_tensor_write_plus_read": 2 * 12 * S * P,
                                                           "selection_write_plus_read": 2 * 4 * S * P},
                          "bit_identical": bool(torch.equal(out, out2))}), flush=True)
        del index, out, out2, patches, pos


if __name__ == "__main__":
    main()
