"""Drop-in for ``provider.get_data_loader`` of the reference (utils/provider.py:319-429): same
arguments and return value ``(dataloader, dataset)``; iterating the loader yields
``[points [B,S*P,3], *targets, trans [B,3,3], n_eff [B,S]]`` exactly like a collated reference
batch, but a batch is produced by one ball-query launch on the GPU instead of B kd-tree queries
on one Python thread (``workers`` is accepted and ignored: there is no host loop to parallelise).
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

    def __len__(self):
        return int(math.ceil(len(self.sampler) / float(self.batch_size)))

    def __iter__(self):
        batch = []
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
