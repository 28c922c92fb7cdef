#!/usr/bin/env python
"""Where the time of the consumer's bf16x3 mode goes: the three device entry points it is built from, each call bracketed by
CUDA events (serialised: the sum is what a forward pass costs, launch gaps excluded).  One JSON line."""
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200 import moe_engine as me  # noqa: E402
from nesti_net_b200.experts_net import ExpertsNormalEstimator  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402


def main():
    B = int(os.environ.get("B", 1024))
    pts = synthetic_cloud(100000, cloud_id=0, noise=0.001)
    g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    radius = [0.01, 0.03, 0.05, 0.07]
    index = mb.PointIndex(pts, cell_frac=max(radius))
    q = np.random.RandomState(0).choice(100000, B, replace=False)
    mups = mb.mups_features(index, gmm, q, index.absolute_radii(radius), 512, seed=3627473)
    torch.manual_seed(1234)
    net = ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval().cuda()
    tc = me.TensorCoreExperts(net, precision=os.environ.get("PRECISION", "bf16x3"))
    tc.predict(mups)
    torch.cuda.synchronize()
    acc, cnt = defaultdict(float), defaultdict(int)

    def timed(name, fn, key=lambda a: ""):
        def wrapper(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            e1.synchronize()
            acc[name + key(a)] += e0.elapsed_time(e1)
            cnt[name + key(a)] += 1
            return r
        return wrapper
    me.conv3d_bn_relu = timed("conv", me.conv3d_bn_relu, lambda a: " D=%d k=%d" % (a[0].shape[1] if a[0].ndim == 5 else 1, a[3].k))
    me.split_x3 = timed("split", me.split_x3)
    me.pool3d_x3 = timed("pool_x3", me.pool3d_x3, lambda a: " D=%d k=%d max=%d" % (a[0].shape[1], a[3], int(a[4])))
    for name in ("pool3d", "conv1_split", "avgpool_bn_relu", "avgpool_f32_x3"):
        setattr(me, name, timed(name, getattr(me, name)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tc.predict(mups)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"precision": tc.precision, "batch": B, "forward_ms_serialised": round(e0.elapsed_time(e1), 2),
                      "ms": {k: round(v, 2) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])}, "calls": dict(cnt)}))


if __name__ == "__main__":
    main()
