#!/usr/bin/env python
"""Statistics kernel time per (query, scale) item as a function of the unmasked point count m:
T(m) = a + b*m.  `a` is the fixed cost of an item (setup, epilogue, channel norms, stores), `b` the
per-point cost (factor staging + 512 pairs of the main loop).  Development aid."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200 import _lib  # noqa: E402


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
    B, S, P, res = int(os.environ.get("NQ", 16384)), 4, 512, 8
    g = mb.get_3d_grid_gmm([res] * 3, 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    rng = np.random.RandomState(0)
    x = rng.normal(size=(B, S * P, 3)).astype(np.float32) * 0.4
    x /= np.maximum(1.0, np.linalg.norm(x, axis=2, keepdims=True))
    patches = torch.from_numpy(x).cuda()
    feats = torch.empty((B, res, res, res, 20 * S), dtype=torch.float32, device="cuda")
    variants = [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]
    for variant in variants:
        _lib.set_option("stats_variant", variant)
        rows = []
        for m in (2, 16, 32, 64, 128, 129, 256, 384, 512):
            n_eff = torch.full((B, S), m - 1 if m < P else P, dtype=torch.int32, device="cuda")
            t = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=feats), iters=5, warm=2)
            rows.append((m, t))
            print(json.dumps({"variant": variant, "m": m, "ms": round(t, 4), "ns_per_item": round(t * 1e6 / (B * S), 1),
                              "Tpairs_per_s": round(B * S * m * 512 / t / 1e9, 4)}))
        ms = np.array([r[0] for r in rows], float)
        ts = np.array([r[1] for r in rows], float) * 1e6 / (B * S)
        b, a = np.polyfit(ms, ts, 1)
        print(json.dumps({"variant": variant, "fit_ns_per_item": {"fixed_a": round(a, 1), "per_point_b": round(b, 3),
                                                                  "fixed_in_points": round(a / b, 1)}}))
    _lib.set_option("stats_variant", 0)


if __name__ == "__main__":
    main()
