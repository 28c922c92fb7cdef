#!/usr/bin/env python
"""SURVEY.md 8(f) rank 1, the consumer of MuPS: forward throughput of the (random-init) Mixture-of-Experts on one B200 --
the hand-written tcgen05 engine (moe_engine.TensorCoreExperts: bf16 products, fp32 accumulation in tensor memory) against
torch / cuDNN library kernels in strict fp32, TF32 and bf16 autocast -- and what each does to the normals relative to strict
fp32 on the host (the checker of the fourth parity gate).  MuPS inputs come from the GPU
path on a synthetic cloud.  One JSON line per mode."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nesti_net_b200 as mb  # noqa: E402
from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg  # noqa: E402
from nesti_net_b200.moe_engine import TensorCoreExperts  # noqa: E402
from nesti_net_b200.synthetic import synthetic_cloud  # noqa: E402

SEED = 3627473
RADIUS = [0.01, 0.03, 0.05, 0.07]


def main():
    B = int(os.environ.get("B", 256))
    n_ref = 32
    pts = synthetic_cloud(100000, cloud_id=0, noise=0.001)
    g = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    gmm = mb.gmm_handle(g.weights_, g.means_, np.sqrt(g.covariances_))
    index = mb.PointIndex(pts, cell_frac=max(RADIUS))
    q = np.random.RandomState(0).choice(100000, B, replace=False)
    mups = mb.mups_features(index, gmm, q, index.absolute_radii(RADIUS), 512, seed=SEED)          # [B,8,8,8,80] cuda
    torch.manual_seed(1234)
    net = ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        t0 = time.time()
        ref_n, ref_e, _ = net.predict(mups[:n_ref].cpu())
        cpu_s = time.time() - t0
    print(json.dumps({"mode": "host fp32 (checker)", "queries": n_ref, "queries_per_s": round(n_ref / cpu_s, 2),
                      "cores": os.cpu_count()}), flush=True)
    net = net.cuda()
    modes = [("gpu fp32 strict", False, None), ("gpu tf32", True, None), ("gpu bf16 autocast", True, torch.bfloat16)]
    if os.environ.get("ONLY_TC"):            # skip the cuDNN modes (A/B runs of the engine)
        modes = []
    for name, tf32, amp in modes:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

        def run(x):
            with torch.no_grad():
                if amp is not None:
                    with torch.autocast("cuda", dtype=amp):
                        return net.predict(x)
                return net.predict(x)
        try:
            small = mups[:n_ref]
            t0 = time.time()
            n_s, e_s, _ = run(small)
            torch.cuda.synchronize()
            first_s = time.time() - t0
            if first_s > 60 and not tf32:
                print(json.dumps({"mode": name, "skipped_timing": "first %d queries took %.1f s" % (n_ref, first_s)}), flush=True)
                iters, ms = 0, None
            else:
                run(mups)
                torch.cuda.synchronize()
                iters = 3
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    run(mups)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
            same = (e_s.cpu() == ref_e)
            rms = float(angular_rms_deg(n_s.float().cpu()[same], ref_n[same])) if bool(same.any()) else None
            print(json.dumps({"mode": name, "batch": B, "ms_per_batch": None if ms is None else round(ms, 2),
                              "queries_per_s": None if ms is None else round(B / ms * 1e3, 1),
                              "normals_rms_deg_vs_host_fp32": rms, "same_expert": "%d/%d" % (int(same.sum()), n_ref)}), flush=True)
        except Exception as e:      # a mode the library cannot run is a finding, not a failure of the script
            print(json.dumps({"mode": name, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}), flush=True)


    # ---- the hand-written tensor-core engine --------------------------------------------------------------------
    from nesti_net_b200 import moe_engine
    moe_engine.FUSE_POOL_BRANCH = os.environ.get("FUSE_POOL", "1") != "0"      # A/B: pool branch's convolution fused with `one`
    tc = TensorCoreExperts(net)
    # TC_VARIANTS="pool,conv;..." : mups_set_option pool_variant / conv_variant pairs to A/B (default: the shipped policy)
    variants = [tuple(int(v) for v in pair.split(",")) for pair in os.environ.get("TC_VARIANTS", "0,0").split(";")]
    for pool_v, conv_v, bsz in [(p, c, b) for (p, c) in variants for b in sorted({B, int(os.environ.get("B_TC", 1024))})]:
        mb._lib.set_option("pool_variant", pool_v)
        mb._lib.set_option("conv_variant", conv_v)
        qq = np.random.RandomState(1).choice(100000, bsz, replace=False)
        x = mb.mups_features(index, gmm, qq, index.absolute_radii(RADIUS), 512, seed=SEED)
        tc.predict(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = tc.predict(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        n_s, e_s, _ = tc.predict(mups[:n_ref])
        same = (e_s.cpu() == ref_e)
        rms = float(angular_rms_deg(n_s.float().cpu()[same], ref_n[same])) if bool(same.any()) else None
        flops = 6.0e10 * bsz
        print(json.dumps({"mode": "tcgen05 engine (bf16 x bf16 -> fp32 in TMEM)", "pool_variant": pool_v, "conv_variant": conv_v, "fuse_pool_branch": moe_engine.FUSE_POOL_BRANCH,
                          "batch": bsz, "ms_per_batch": round(ms, 2),
                          "queries_per_s": round(bsz / ms * 1e3, 1), "approx_TFLOPs": round(flops / ms / 1e9, 1),
                          "normals_rms_deg_vs_host_fp32": rms, "same_expert": "%d/%d" % (int(same.sum()), n_ref)}), flush=True)

    # ---- the same kernels in bf16x3 mode (hi / lo pairs, [w_hi | w_hi | w_lo] weights: 3 x the tensor work, fp32-grade) ----
    if os.environ.get("X3", "1") != "0":
        mb._lib.set_option("pool_variant", 0)
        mb._lib.set_option("conv_variant", 0)
        tc3 = TensorCoreExperts(net, precision="bf16x3")
        for bsz in sorted({B, int(os.environ.get("B_X3", 512))}):
            qq = np.random.RandomState(1).choice(100000, bsz, replace=False)
            x = mb.mups_features(index, gmm, qq, index.absolute_radii(RADIUS), 512, seed=SEED)
            tc3.predict(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                tc3.predict(x)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            n_s, e_s, _ = tc3.predict(mups[:n_ref])
            same = (e_s.cpu() == ref_e)
            rms = float(angular_rms_deg(n_s.float().cpu()[same], ref_n[same])) if bool(same.any()) else None
            print(json.dumps({"mode": "tcgen05 engine, bf16x3 (a_hi w_hi + a_lo w_hi + a_hi w_lo -> fp32 in TMEM)", "batch": bsz,
                              "ms_per_batch": round(ms, 2), "queries_per_s": round(bsz / ms * 1e3, 1),
                              "approx_TFLOPs_executed": round(3 * 6.0e10 * bsz / ms / 1e9, 1),
                              "normals_rms_deg_vs_host_fp32": rms, "same_expert": "%d/%d" % (int(same.sum()), n_ref)}), flush=True)


if __name__ == "__main__":
    main()
