#!/usr/bin/env python
"""Workload for the ncu captures of the dense-cloud kernels (run under `ncu -k regex:...`): index build (K1 + K2) of the
10 M-point cloud of BASELINE configs[4] and the hierarchical ball query on NQ of its points.

    ncu --set full --clock-control none --import-source on -k regex:"bbox|cell_code|scan_|scatter|ball_query_hier|order_" \
        -o gpurun_out/r02_dense python profiles/prof_dense.py [NQ=4440] [N=10000000] [ORDER=2]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200 import _lib  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
n = int(opts.get("N", 10000000))
nq = int(opts.get("NQ", 4440))            # 10 waves of 148 x 3 CTAs
radius = [0.01, 0.03, 0.05, 0.07]
pts = synthetic_cloud(n, cloud_id=2 if n >= 5000000 else 1, kind="pcpnet" if n >= 5000000 else "scan")
xyz = torch.from_numpy(pts).cuda()
_lib.set_option("query_order", int(opts.get("ORDER", 2)))
index = mb.PointIndex(xyz, cell_frac=max(radius))
radii = index.absolute_radii(radius)
q = torch.from_numpy(np.random.RandomState(5).choice(n, nq, replace=False).astype(np.int64)).cuda()
out = index.ball_query(q, radii, 512, seed=3627473)
torch.cuda.synchronize()
print("neighbours per query (mean):", out[2].float().mean(0).tolist())
