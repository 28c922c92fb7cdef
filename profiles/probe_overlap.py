#!/usr/bin/env python
"""Does the ball query of the NEXT cloud really run under the statistics kernel of the current one?

bench.py launches half 1 (index build + ball query) of step i + 1 on a high-priority side stream while half 2 (ONE statistics
launch, 400 k CTAs that own every register of every SM) of step i runs on the main stream, and gains < 1 ms of the 7.5 ms.
This probe times, on the configs[1] cloud (100 k queries, 4 scales, P = 512, 8^3):

  alone        statistics kernel alone, ball query alone
  side_prio    statistics on stream A, ball query launched right behind it on a high-priority stream B: when does B finish
               (measured from the start of the statistics kernel), how long does the statistics kernel take
  chunks=N     the same with the statistics work issued as N launches (kernel boundaries for the block scheduler)

MAIN=default|pool selects the legacy default stream or a pool stream for A.  One JSON line per case.

Result (profiles/r02_overlap_probe.jsonl): the priority works -- the ball query finishes 7.7 ms after the start of the
statistics kernel (7.5 alone) -- but the statistics kernel then takes 68.7 instead of 61.1 ms: the sum.  A higher-priority
grid takes every slot the big grid frees until all of its CTAs are dispatched, so the two kernels run one after the other.
A variant of the flat kernel with a fixed number of RESIDENT CTAs looping over the queries (148 / 296 / 444 / 592 CTAs,
high or equal priority, launched before or after the statistics kernel) was built to force real sharing of the SMs and
measured with this probe: both kernels were done after 73.7-87.2 ms in every configuration, i.e. never sooner than one after
the other (68.7); removed again.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200 import _lib  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]
P, S = 512, 4


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = _lib.load()
    n = 100000
    g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    pts = torch.from_numpy(synthetic_cloud(n, cloud_id=0)).to(dev)
    q = torch.arange(n, dtype=torch.int64, device=dev)
    index = mb.PointIndex(pts, cell_frac=max(RADIUS))
    radii = np.ascontiguousarray(index.absolute_radii(RADIUS), dtype=np.float64)
    dptr = ctypes.POINTER(ctypes.c_double)
    patches = [torch.empty((n, S * P, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    n_eff = [torch.empty((n, S), dtype=torch.int32, device=dev) for _ in range(2)]
    total = torch.empty((n, S), dtype=torch.int32, device=dev)
    feats = torch.empty((n, 8, 8, 8, 20 * S), dtype=torch.float32, device=dev)

    main_s = torch.cuda.current_stream(dev) if os.environ.get("MAIN", "default") == "default" else torch.cuda.Stream(dev)
    side = torch.cuda.Stream(dev, priority=-1)

    def ball_query(slot, st):
        _lib.check(L.mups_ball_query(index.handle, ctypes.c_void_p(q.data_ptr()), n, radii.ctypes.data_as(dptr), S, P, SEED, None,
                                     ctypes.c_void_p(total.data_ptr()), ctypes.c_void_p(patches[slot].data_ptr()),
                                     ctypes.c_void_p(n_eff[slot].data_ptr()), ctypes.c_void_p(st.cuda_stream)))

    def stats(slot, st, chunks=1):
        per = n // chunks
        for c in range(chunks):
            lo, hi = c * per, (n if c == chunks - 1 else (c + 1) * per)
            _lib.check(L.mups_3dmfv(gmm.handle, ctypes.c_void_p(patches[slot][lo:hi].data_ptr()), ctypes.c_void_p(n_eff[slot][lo:hi].data_ptr()),
                                    hi - lo, S, P, _lib.FLAG_MASKED, ctypes.c_void_p(feats[lo:hi].data_ptr()), ctypes.c_void_p(st.cuda_stream)))

    ev = lambda: torch.cuda.Event(enable_timing=True)
    ball_query(0, main_s); ball_query(1, main_s); stats(0, main_s)
    torch.cuda.synchronize()

    a, b, c = ev(), ev(), ev()
    a.record(main_s); stats(0, main_s); b.record(main_s); ball_query(1, main_s); c.record(main_s)
    torch.cuda.synchronize()
    print(json.dumps({"case": "alone", "main": os.environ.get("MAIN", "default"), "stats_ms": round(a.elapsed_time(b), 2),
                      "ball_query_ms": round(b.elapsed_time(c), 2)}), flush=True)

    for chunks in (1, 4, 16):
        for rep in range(2):
            s0, s1, q1 = ev(), ev(), ev()
            s0.record(main_s)
            side.wait_event(s0)                    # B may start as soon as the statistics launch may
            stats(0, main_s, chunks)
            s1.record(main_s)
            ball_query(1, side)
            q1.record(side)
            torch.cuda.synchronize()
            print(json.dumps({"case": "side_prio", "chunks": chunks, "stats_ms": round(s0.elapsed_time(s1), 2),
                              "ball_query_done_after_ms": round(s0.elapsed_time(q1), 2)}), flush=True)


if __name__ == "__main__":
    main()
