"""Drop-in for ``provider.get_data_loader`` of the reference (utils/provider.py:319-429): same
arguments and return value ``(dataloader, dataset)``; iterating the loader yields
``[points [B,S*P,3], *targets, trans [B,3,3], n_eff [B,S]]`` exactly like a collated reference
batch, but a batch is produced by one ball-query launch per distinct shape in it instead of B kd-tree
queries on one Python thread (``workers`` is accepted and ignored: there is no host loop to parallelise).
"""
import math

from .pcpnet_dataset import (PointcloudPatchDataset, RandomPointcloudPatchSampler,
                             SequentialPointcloudPatchSampler, SequentialShapeRandomPointcloudPatchSampler)


class PatchBatchLoader(object):
    """Iterable with ``len()`` = number of batches, like torch.utils.data.DataLoader."""

    def __init__(self, dataset, sampler, batch_size):
        self.dataset = dataset
        self.sampler = sampler
        self.batch_size = int(batch_size)
        self._started = False

    def __len__(self):
        return int(math.ceil(len(self.sampler) / float(self.batch_size)))

    def __iter__(self):
        batch = []
        if self._started and hasattr(self.dataset, "begin_epoch"):
            self.dataset.begin_epoch()          # a fresh subsample per epoch unless identical_epochs
        self._started = True
        for idx in self.sampler:
            batch.append(int(idx))
            if len(batch) == self.batch_size:
                yield self.dataset.get_batch(batch)
                batch = []
        if batch:
            yield self.dataset.get_batch(batch)


def get_data_loader(dataset_name='trainingset_temp.txt', batchSize=128, indir='./pclouds', patch_radius=[0.05],
                    points_per_patch=500, outputs=['unoriented_normals'], patch_point_count_std=0,
                    seed=3627473, identical_epochs=False, use_pca=False, patch_center='point',
                    point_tuple=1, cache_capacity=100, patches_per_shape=1000, patch_sample_order='random',
                    workers=0, dataset_type='training', sparse_patches=False):
    target_features = []
    for o in outputs:
        if o in ('unoriented_normals', 'oriented_normals'):
            if 'normal' not in target_features:
                target_features.append('normal')
        elif o in ('max_curvature', 'min_curvature'):
            if o not in target_features:
                target_features.append(o)
        elif o == 'noise':
            target_features.append(o)
        else:
            raise ValueError('Unknown output: %s' % (o))

    dataset = PointcloudPatchDataset(
        root=indir, shape_list_filename=dataset_name, patch_radius=patch_radius,
        points_per_patch=points_per_patch, patch_features=target_features,
        point_count_std=patch_point_count_std, seed=seed, identical_epochs=identical_epochs,
        use_pca=use_pca, center=patch_center, point_tuple=point_tuple, cache_capacity=cache_capacity,
        sparse_patches=sparse_patches)

    if patch_sample_order == 'random':
        sampler = RandomPointcloudPatchSampler(dataset, patches_per_shape=patches_per_shape, seed=seed,
                                               identical_epochs=identical_epochs)
    elif patch_sample_order == 'random_shape_consecutive':
        sampler = SequentialShapeRandomPointcloudPatchSampler(dataset, patches_per_shape=patches_per_shape,
                                                              seed=seed, identical_epochs=identical_epochs)
    elif patch_sample_order == 'full':
        sampler = SequentialPointcloudPatchSampler(dataset)
    else:
        raise ValueError('Unknown patch sampling order: %s' % (patch_sample_order))

    dataloader = PatchBatchLoader(dataset, sampler, batchSize)
    print(dataset_type + ' set: %d patches (in %d batches))' % (len(sampler), len(dataloader)))
    return (dataloader, dataset)


# --------------------------------------------------------------------------------------------------
# training-time rotation augmentation (train_n_est_w_experts.py:262-274), between half 1 and half 2
# --------------------------------------------------------------------------------------------------

def euler2mat(z=0.0, y=0.0, x=0.0):
    """Rotation matrix Rx(x) . Ry(y) . Rz(z) with the reference's conventions (utils/eulerangles.py:
    the factors for z, y, x are collected in that order, zero angles skipped, and multiplied in reverse)."""
    import numpy as np
    cz, sz, cy, sy, cx, sx = math.cos(z), math.sin(z), math.cos(y), math.sin(y), math.cos(x), math.sin(x)
    rz = np.array([[cz, -sz, 0.0], [sz, cz, 0.0], [0.0, 0.0, 1.0]])
    ry = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cx, -sx], [0.0, sx, cx]])
    ms = [f for f, angle in ((rz, z), (ry, y), (rx, x)) if angle]
    if not ms:
        return np.eye(3)
    m = ms[-1]                      # reduce(np.dot, Ms[::-1]): ((Rx . Ry) . Rz), same association -> same bits
    for f in ms[-2::-1]:
        m = np.dot(m, f)
    return m


def rotation_augmentation(points, normals, rng=None, angles=None):
    """One random rotation per batch applied to the patches and the target normals, as the training loop
    does when ``--insert_rotation_augmentation`` is set: ``angles = 2 pi randn(3)``,
    ``R = euler2mat(z, y, x).T``, ``points[k] @ R`` and ``normals[k] @ R`` evaluated in float64 and
    stored as float32.  ``points`` [B, S*P, 3] / ``normals`` [B, 3] may live on the GPU (the patches
    of a batch stay on the device between the ball query and the statistics kernel).
    Returns (rotated points, rotated normals, R)."""
    import numpy as np
    import torch
    if angles is None:
        rng = np.random if rng is None else rng
        angles = 2 * np.pi * rng.randn(3)
    R = np.transpose(euler2mat(z=angles[0], y=angles[1], x=angles[2]))
    points = torch.as_tensor(points)
    normals = torch.as_tensor(normals)
    Rt = torch.as_tensor(R, dtype=torch.float64, device=points.device)
    rp = torch.matmul(points.to(torch.float64), Rt).to(torch.float32)
    rn = torch.matmul(normals.to(torch.float64), Rt.to(normals.device)).to(torch.float32)
    return rp, rn, R
