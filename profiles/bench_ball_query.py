#!/usr/bin/env python
"""Ball query (K3 + K4) on the dense clouds of BASELINE configs[3] / configs[4]: flat cell scan against the
hierarchical kernel at several grid resolutions and CTA orders, same cloud, same queries, CUDA events after a
warm-up launch, with an oracle check (cKDTree + shared seeded selection) on a strided subset.  One JSON object per line.

    python profiles/bench_ball_query.py [C4 C5] [NQ=65536] [CHECK=32]
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
from nesti_net_b200 import _lib  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]


def main():
    args = [a for a in sys.argv[1:] if "=" not in a] or ["C4", "C5"]
    opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
    nq = int(opts.get("NQ", 65536))
    n_check = int(opts.get("CHECK", 32))
    P = int(opts.get("P", 512))
    for name in args:
        if name == "C2":                      # configs[1]: the 100 k-point cloud of the default bench line
            pts = synthetic_cloud(100000, cloud_id=0)
            nq = min(nq, 100000)
        elif name == "C4":
            pts = synthetic_cloud(2000000, cloud_id=1, kind="scan")
        else:
            pts = synthetic_cloud(10000000, cloud_id=2)
        n = len(pts)
        xyz = torch.from_numpy(pts).cuda()
        q = np.random.RandomState(5).choice(n, nq, replace=False).astype(np.int64)
        qd = torch.from_numpy(q).cuda()
        ref = None
        if n_check:
            from oracle import mups_oracle as orc
            t0 = time.time()
            sub = np.linspace(0, nq - 1, n_check).astype(np.int64)
            ref = orc.gather_patches(pts, q[sub], RADIUS, P, seed=SEED, return_indices=True)
            oracle_s = time.time() - t0
        variants = [("flat", 1, 0.34, 1), ("hier", 2, 0.125, 1), ("hier", 2, 0.125, 2), ("hier", 2, 0.0625, 1),
                    ("hier", 2, 0.0625, 2), ("hier", 2, 0.085, 2), ("flat", 1, 0.34, 2)]
        if opts.get("VARIANTS"):
            variants = [variants[int(i)] for i in opts["VARIANTS"].split(",")]
        if opts.get("SPEC"):                  # SPEC=flat:1.0:1;hier:0.25:2 ... (kernel : cell scale : CTA order)
            variants = [(k, 1 if k == "flat" else 2, float(sc), int(o)) for k, sc, o in (v.split(":") for v in opts["SPEC"].split(";"))]
        for kname, kernel, scale, order in variants:
            nq_v = nq if (kname == "hier" or name == "C2") else min(nq, 8192)
            _lib.set_option("query_kernel", kernel)
            _lib.set_option("query_order", order)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            index = mb.PointIndex(xyz, cell_frac=max(RADIUS), cell_scale=scale)      # warm (pool growth)
            del index
            e0.record()
            index = mb.PointIndex(xyz, cell_frac=max(RADIUS), cell_scale=scale)
            e1.record()
            torch.cuda.synchronize()
            build_ms = e0.elapsed_time(e1)
            radii = index.absolute_radii(RADIUS)
            out = index.ball_query(qd[:nq_v], radii, P, seed=SEED, return_indices=True)      # warm-up
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(2):
                out = index.ball_query(qd[:nq_v], radii, P, seed=SEED, return_indices=True)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 2
            patches, n_eff, total, nbr = out
            rec = {"config": name, "points": n, "P": P, "kernel": kname, "cell_scale": scale,
                   "cta_order": "morton" if order == 2 else "caller", "queries": nq_v, "index_build_ms": round(build_ms, 3),
                   "ball_query_ms": round(ms, 3), "kq_per_s": round(nq_v / ms, 1),
                   "mean_neighbours": [round(float(x), 1) for x in total.float().mean(0).tolist()],
                   "max_neighbours": int(total.max().item()),
                   "algorithmic_GBps": round((16.0 * float(total.sum().item()) + nq_v * (12.0 * 4 * P + 32)) / ms / 1e6, 1)}
            if ref is not None:
                sel = sub[sub < nq_v]
                k = len(sel)
                rec["check_queries"] = int(k)
                rec["check_counts_exact"] = bool(np.array_equal(total.cpu().numpy()[sel], ref[2][:k]))
                rec["check_indices_exact"] = bool(np.array_equal(nbr.cpu().numpy()[sel], ref[3][:k]))
                rec["check_patches_bit_exact"] = bool(np.array_equal(patches.cpu().numpy()[sel].view(np.uint32), ref[0][:k].view(np.uint32)))
                rec["oracle_seconds"] = round(oracle_s, 1)
            print(json.dumps(rec), flush=True)
            del index, out, patches, n_eff, total, nbr
        _lib.set_option("query_kernel", 0)
        _lib.set_option("query_order", 0)


if __name__ == "__main__":
    main()
