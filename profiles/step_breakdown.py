#!/usr/bin/env python
"""Where one bench step spends its device time: CUDA events around each phase (+ host wall time)."""
import ctypes, os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nesti_net_b200 as mb
from nesti_net_b200 import _lib
from oracle import mups_oracle as orc
N = int(os.environ.get("N", 100000)); P = 512; RADIUS = [0.01, 0.03, 0.05, 0.07]; S = 4
ORDER = os.environ.get("ORDER", "natural")
pts = orc.synthetic_cloud(N, cloud_id=0)
xyz = torch.from_numpy(pts).cuda()
g = mb.get_3d_grid_gmm([8] * 3, 0.0156); gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
q = torch.arange(N, dtype=torch.int64, device="cuda")
feats = torch.empty((N, 8, 8, 8, 80), dtype=torch.float32, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(4):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e = [ev() for _ in range(5)]
    e[0].record(); index = mb.PointIndex(xyz, cell_frac=0.07); w1 = time.perf_counter()
    e[1].record(); radii = index.absolute_radii(RADIUS); w2 = time.perf_counter()
    e[2].record(); patches, n_eff, total = index.ball_query(q, radii, P); w3 = time.perf_counter()
    e[3].record(); mb.stats_3dmfv(patches, n_eff, gmm, S, out=feats); w4 = time.perf_counter()
    e[4].record(); torch.cuda.synchronize(); w5 = time.perf_counter()
    print(json.dumps({"iter": it, "dev_ms": {"index_create": e[0].elapsed_time(e[1]), "bbox": e[1].elapsed_time(e[2]),
          "ball_query": e[2].elapsed_time(e[3]), "stats": e[3].elapsed_time(e[4]), "total": e[0].elapsed_time(e[4])},
          "host_ms": {"index_create": (w1-w0)*1e3, "bbox": (w2-w1)*1e3, "ball_query_enqueue": (w3-w2)*1e3, "stats_enqueue": (w4-w3)*1e3, "sync": (w5-w4)*1e3}}))
    del index
