"""Inference driver: a list of `.xyz` clouds -> `.normals`, `.experts`, `.experts_probs` files.

SURVEY.md 8(f) rank 2 ("next" row): the user-visible loop of the reference's
test_n_est_w_experts.py:108-197 -- iterate every patch of every shape in 'full' order, compute
MuPS, run the Mixture-of-Experts, keep the normal of the most probable expert, write one file
triple per shape -- with the hot path on the GPU: batches come from one ball-query launch and one
statistics launch, MuPS is consumed on the device and never stored.
"""
import os

import numpy as np
import torch

from . import mups as _m
from .pipeline import MuPSPipeline
from .provider import get_data_loader


@torch.no_grad()
def estimate_normals(indir, dataset_name, outdir, model, gmm, patch_radius, points_per_patch=512, batch_size=128,
                     sparse_patches=False, seed=3627473, write=True):
    """Returns {shape name: (normals [n,3], experts [n], experts_probs [n, n_experts])} and, when
    `write`, saves them with np.savetxt exactly as test_n_est_w_experts.py:185-191 does.

    model: nesti_net_b200.experts_net.ExpertsNormalEstimator (any device; MuPS is moved to it), or a
           nesti_net_b200.moe_engine.TensorCoreExperts (the same network packed for the tcgen05 kernels: MuPS stays on the GPU)
    gmm:   GridGMM-like (weights_, means_, covariances_)"""
    loader, dataset = get_data_loader(
        dataset_name=dataset_name, batchSize=batch_size, indir=indir, patch_radius=patch_radius,
        points_per_patch=points_per_patch, outputs=[], patch_point_count_std=0, seed=seed, identical_epochs=False,
        use_pca=False, patch_center='point', point_tuple=1, cache_capacity=100, patch_sample_order='full',
        workers=0, dataset_type='test', sparse_patches=sparse_patches)
    handle = _m.gmm_handle(gmm.weights_, gmm.means_, np.sqrt(gmm.covariances_))
    packed = not hasattr(model, "parameters")
    model_device = model.device if packed else next(model.parameters()).device
    if not packed:
        model.eval()
    cudnn_benchmark = torch.backends.cudnn.benchmark
    if model_device.type == "cuda" and not packed:
        torch.backends.cudnn.benchmark = True      # the 8^3 conv3d stack is 2-3x faster with cuDNN's tuned algorithms (restored below)
    n_rads = len(patch_radius)
    normals, experts, probs = [], [], []
    try:
        for data in loader:
            points, n_eff = data[0], data[-1]
            mups = _m.stats_3dmfv(points, n_eff, handle, n_rads, masked=True, layout="mups")
            normal, expert, prob = model.predict(mups.to(model_device))
            normals.append(normal.float().cpu().numpy())
            experts.append(expert.cpu().numpy())
            probs.append(prob.float().cpu().numpy())
    finally:
        torch.backends.cudnn.benchmark = cudnn_benchmark
    normals, experts, probs = np.concatenate(normals), np.concatenate(experts), np.concatenate(probs)
    out, offset = {}, 0
    if write and not os.path.exists(outdir):
        os.makedirs(outdir)
    for name, count in zip(dataset.shape_names, dataset.shape_patch_count):
        sl = slice(offset, offset + count)
        out[name] = (normals[sl], experts[sl], probs[sl])
        if write:
            np.savetxt(os.path.join(outdir, name + '.normals'), normals[sl])
            np.savetxt(os.path.join(outdir, name + '.experts'), experts[sl].astype(int), fmt='%i')
            np.savetxt(os.path.join(outdir, name + '.experts_probs'), probs[sl])
        offset += count
    return out


class CloudNormalEstimator(object):
    """Whole-cloud inference without the dataset machinery: host cloud in -> index build -> per chunk of query points the ball
    query hands its selection straight to the statistics kernel (no patch tensor) -> MuPS stays on the device and feeds the
    Mixture-of-Experts -> normals / experts / probabilities stream into pinned host memory.  The same numbers as
    ``estimate_normals`` (same shared seeded selection, same features, same network), at the speed of the device path:
    what ``bench.py`` reports as ``e2e.normals``.

    model: ``moe_engine.TensorCoreExperts`` (the tensor-core engine; ``precision="bf16x3"`` for fp32-grade normals at a third of
    the throughput) or a CUDA ``ExpertsNormalEstimator``."""

    def __init__(self, model, gmm, patch_radius, points_per_patch=512, seed=3627473, chunk=2048):
        self.model = model
        self.pipe = MuPSPipeline(gmm, patch_radius, points_per_patch, seed=seed, chunk=chunk)
        self.n_rads = len(patch_radius)

    @torch.no_grad()
    def __call__(self, points, query_idx=None):
        """points: [N, 3] float32 host array / tensor; query_idx: int64 indices of the centres (default: every point).
        Returns (normals [B, 3] float32, experts [B] int64, experts_probs [B, n_experts] float32) as numpy arrays."""
        pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)) if not torch.is_tensor(points) else points
        B = int(pts.shape[0]) if query_idx is None else int(len(query_idx))
        res = self.pipe.feat_shape[0]
        out = {}

        def consume(lo, hi, rows):
            normal, expert, prob = self.model.predict(rows.view(hi - lo, res, res, res, 20 * self.n_rads))
            if not out:
                out["n"] = torch.empty((B, 3), dtype=torch.float32).pin_memory()
                out["e"] = torch.empty((B,), dtype=torch.int64).pin_memory()
                out["p"] = torch.empty((B, int(prob.shape[1])), dtype=torch.float32).pin_memory()
            out["n"][lo:hi].copy_(normal.float(), non_blocking=True)
            out["e"][lo:hi].copy_(expert, non_blocking=True)
            out["p"][lo:hi].copy_(prob.float(), non_blocking=True)
        self.pipe.features_to_consumer(pts, query_idx, consume)
        torch.cuda.synchronize(self.pipe.device)
        if not out:
            return np.zeros((0, 3), np.float32), np.zeros((0,), np.int64), np.zeros((0, 0), np.float32)
        return out["n"].numpy(), out["e"].numpy(), out["p"].numpy()
