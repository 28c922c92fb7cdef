#!/usr/bin/env python
"""Workload for the ncu captures of the tensor-core consumer: one forward of moe_engine.TensorCoreExperts on B queries of a
100 k-point cloud (random-init network).   ncu ... python profiles/prof_moe.py [B=256]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200.experts_net import ExpertsNormalEstimator  # noqa: E402
from nesti_net_b200.moe_engine import TensorCoreExperts  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
B = int(opts.get("B", 256))
radius = [0.01, 0.03, 0.05, 0.07]
pts = synthetic_cloud(100000, cloud_id=0, noise=0.001)
g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
index = mb.PointIndex(pts, cell_frac=max(radius))
q = np.random.RandomState(0).choice(100000, B, replace=False)
mups = mb.mups_features(index, gmm, q, index.absolute_radii(radius), 512, seed=3627473)
torch.manual_seed(1234)
tc = TensorCoreExperts(ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval().cuda())
for _ in range(int(opts.get("REPS", 1))):
    out = tc.predict(mups)
torch.cuda.synchronize()
print("ok", tuple(out[0].shape))
