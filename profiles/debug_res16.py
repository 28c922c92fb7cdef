import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import nesti_net_b200 as mb
from oracle import mups_oracle as orc, c_oracle
res, P, S, var = 16, 128, 2, 0.00390625
w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([res] * 3, var))
rng = np.random.RandomState(res * 1000 + P)
edge = [1, 2, 3, P // 2, P - 2, P - 1, P]
B = 12
ne = np.array([[edge[(b + s) % len(edge)] for s in range(S)] for b in range(B)], np.int32)
ne[-1] = rng.randint(1, P + 1, S)
pts = np.zeros((B, S * P, 3), np.float32)
for b in range(B):
    for s in range(S):
        x = rng.normal(size=(ne[b, s], 3)) * rng.uniform(0.1, 0.6)
        x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
        x[0] = 0
        pts[b, s * P: s * P + ne[b, s]] = x
gmm = mb.gmm_handle(w, mu, sg)
fast = mb.stats_3dmfv(pts, ne, gmm, S, fastpath=True).cpu().numpy().reshape(B, -1, S, 20)
slow = mb.stats_3dmfv(pts, ne, gmm, S, fastpath=False).cpu().numpy().reshape(B, -1, S, 20)
ref = c_oracle.mups(pts, ne, w, mu, sg, S).reshape(B, -1, S, 20)
f64 = np.stack([orc.get_3dmfv_n_est_f64(pts[:, s * P:(s + 1) * P], w, mu, sg, ne[:, s]) for s in range(S)], 0)  # [S,B,20,G]
f64 = f64.transpose(1, 3, 0, 2)
for name, x in (("fast", fast), ("general", slow), ("c_oracle", ref)):
    e = np.abs(x - f64)
    print(name, "vs f64: max err", e.max(), "argmax (b,g,s,c)", np.unravel_index(e.argmax(), e.shape), "n_eff there", ne[np.unravel_index(e.argmax(), e.shape)[0]])
    print("   per-channel max err", np.round(e.max(axis=(0, 1, 2)) * 1e6, 2))
    print("   per-(b,s) max err", np.round(e.max(axis=(1, 3)) * 1e6, 2).tolist())
