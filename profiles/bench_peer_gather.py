#!/usr/bin/env python
"""The optional gather of SURVEY.md 8e done by the statistics kernel itself: every rank computes its shard of
the query list and stores the features straight into rank 0's buffer through peer memory (PeerSlabGather), against
(a) local slabs only and (b) local slabs followed by the NCCL all-gather of dist.gather_slabs.  Checks that rank 0's
tensor equals its own single-GPU result bit for bit.  torchrun, one rank per GPU; rank 0 prints JSON lines.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 profiles/bench_peer_gather.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from oracle import mups_oracle as orc  # noqa: E402   (synthetic cloud only)

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]
P = 512


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 100000
    nq = int(os.environ.get("NQ", 32768))
    pts = orc.synthetic_cloud(n, cloud_id=0)
    g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    S = len(RADIUS)
    index = mb.PointIndex(torch.from_numpy(pts).to(dev), cell_frac=max(RADIUS))
    radii = index.absolute_radii(RADIUS)
    q_all = torch.arange(nq, dtype=torch.int64, device=dev) * (n // nq)
    mine, lo, hi = mb.dist.shard_queries(q_all, rank, world)
    patches, n_eff, _ = index.ball_query(mine, radii, P, seed=SEED)
    row = (8, 8, 8, 20 * S)
    gather = mb.dist.PeerSlabGather(nq, row, dst=0)
    slab = torch.empty((hi - lo,) + row, dtype=torch.float32, device=dev)

    def timed(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_local = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=slab))
    t_peer = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=gather.target(lo, hi)))
    t_peer_wide = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=gather.target(lo, hi), wide_stores=True))
    t_local_wide = timed(lambda: mb.stats_3dmfv(patches, n_eff, gmm, S, out=slab, wide_stores=True))
    t_nccl = timed(lambda: mb.dist.gather_slabs(mb.stats_3dmfv(patches, n_eff, gmm, S, out=slab)))
    full_dst = torch.empty((nq,) + row, dtype=torch.float32, device=dev) if rank == 0 else None

    def send_to_rank0():
        mb.stats_3dmfv(patches, n_eff, gmm, S, out=slab)
        if world == 1:
            full_dst[lo:hi].copy_(slab)
        elif rank == 0:
            full_dst[lo:hi].copy_(slab)
            reqs = []
            for r in range(1, world):
                rlo, rhi = mb.dist.shard_bounds(nq, world)[r], mb.dist.shard_bounds(nq, world)[r + 1]
                reqs.append(dist.irecv(full_dst[int(rlo):int(rhi)], src=r))
            for w in reqs:
                w.wait()
        else:
            dist.send(slab, dst=0)
    t_send = timed(send_to_rank0)
    # correctness: rank 0's gathered tensor against its own single-GPU computation of the whole list
    gather.result().zero_()
    torch.cuda.synchronize()
    dist.barrier()
    mb.stats_3dmfv(patches, n_eff, gmm, S, out=gather.target(lo, hi), wide_stores=True)
    gather.finish()
    ok = True
    if rank == 0:
        full = mb.mups_features(index, gmm, q_all, radii, P, seed=SEED)
        ok = bool(torch.equal(gather.result(), full))
        gb = nq * 512 * 80 * 4 / 1e9
        print(json.dumps({"n_gpus": world, "queries": nq, "output_GB": round(gb, 2),
                          "stats_local_slabs_ms": round(t_local, 3),
                          "stats_storing_into_rank0_peer_memory_ms": round(t_peer, 3),
                          "stats_wide_stores_into_rank0_peer_memory_ms": round(t_peer_wide, 3),
                          "stats_wide_stores_local_ms": round(t_local_wide, 3),
                          "stats_then_nccl_send_to_rank0_ms": round(t_send, 3),
                          "stats_then_nccl_all_gather_ms": round(t_nccl, 3),
                          "peer_store_overhead_pct": round(100.0 * (t_peer_wide / t_local - 1.0), 1),
                          "gathered_equals_single_gpu_bit_for_bit": ok}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
