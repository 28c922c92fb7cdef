"""GPU-backed drop-in for the patch loader of the reference's utils/pcpnet_dataset.py.

``PointcloudPatchDataset`` keeps the reference constructor signature
(utils/pcpnet_dataset.py:182-183), item layout (:419) and the attributes its callers read
(``shape_names``, ``shape_patch_count``, ``patch_radius_absolute``).  The per-item kd-tree query,
subsample, centring and normalisation (:303-343) run in the ball-query kernel of
libmups_b200.so; ``get_batch`` answers a whole batch of patch indices with one launch per distinct
shape of the batch and is what ``provider.get_data_loader`` iterates.

Only the options every reference script uses are on the hot path (``use_pca=False``,
``center='point'``, ``point_tuple=1``, ``point_count_std=0``); the others raise.
The subsample is the shared seeded selection (a pure function of seed, centre, scale) instead
of the reference's stateful ``rng.choice`` stream -- see DESIGN.md.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from .mups import PointIndex


class Shape(object):
    """One cloud: host points (float32 [N,3]) + its GPU spatial index (replaces the kd-tree)."""

    def __init__(self, pts, index, normals=None, curv=None, pidx=None, noise_level=None):
        self.pts = pts
        self.index = index
        self.normals = normals
        self.curv = curv
        self.pidx = pidx
        self.noise_level = noise_level


def _load_table(path, dtype):
    """Text table (``.xyz`` / ``.normals`` / ``.curv`` / ``.pidx``) -> array.  A ``.npy`` sibling written by the
    reference (pcpnet_dataset.py:251) is used only while it is at least as new as the text file: the reference
    regenerates it from the text in every constructor (:250-251), so an edited text file must win over a stale ``.npy``."""
    npy = path + ".npy"
    if os.path.exists(npy) and (not os.path.exists(path) or os.path.getmtime(npy) >= os.path.getmtime(path)):
        return np.load(npy).astype(dtype)
    return np.loadtxt(path).astype(dtype)


class ShapeCache(object):
    """LRU cache of loaded shapes (the role of reference Cache, pcpnet_dataset.py:151-176)."""

    def __init__(self, capacity, loader):
        self.capacity = max(1, int(capacity))
        self.loader = loader
        self.items = OrderedDict()

    def get(self, shape_ind):
        if shape_ind in self.items:
            self.items.move_to_end(shape_ind)
            return self.items[shape_ind]
        while len(self.items) >= self.capacity:
            self.items.popitem(last=False)
        shape = self.items[shape_ind] = self.loader(shape_ind)
        return shape


class SequentialPointcloudPatchSampler(torch.utils.data.Sampler):
    """All patches of all shapes in order (reference :41-55; the 'full' order of inference)."""

    def __init__(self, data_source):
        self.data_source = data_source
        self.total_patch_count = int(sum(data_source.shape_patch_count))

    def __iter__(self):
        return iter(range(self.total_patch_count))

    def __len__(self):
        return self.total_patch_count


class RandomPointcloudPatchSampler(torch.utils.data.Sampler):
    """Random subset of min(patches_per_shape, count) patches per shape worth of indices, drawn
    without replacement over the whole dataset (reference :112-138)."""

    def __init__(self, data_source, patches_per_shape, seed=None, identical_epochs=False):
        self.data_source = data_source
        self.patches_per_shape = patches_per_shape
        self.identical_epochs = identical_epochs
        self.seed = int(np.random.randint(0, 2 ** 32 - 1)) if seed is None else seed
        self.rng = np.random.RandomState(self.seed)
        self.total_patch_count = int(sum(min(patches_per_shape, c) for c in data_source.shape_patch_count))

    def __iter__(self):
        if self.identical_epochs:
            self.rng.seed(self.seed)
        return iter(self.rng.choice(int(sum(self.data_source.shape_patch_count)), size=self.total_patch_count,
                                    replace=False))

    def __len__(self):
        return self.total_patch_count


class SequentialShapeRandomPointcloudPatchSampler(torch.utils.data.Sampler):
    """Shapes in (optionally permuted) order, a random patch subset inside each shape, patches of
    one shape adjacent (reference :56-110)."""

    def __init__(self, data_source, patches_per_shape, seed=None, sequential_shapes=False, identical_epochs=False):
        self.data_source = data_source
        self.patches_per_shape = patches_per_shape
        self.sequential_shapes = sequential_shapes
        self.identical_epochs = identical_epochs
        self.seed = int(np.random.randint(0, 2 ** 32 - 1)) if seed is None else seed
        self.rng = np.random.RandomState(self.seed)
        self.shape_patch_inds = None
        self.total_patch_count = int(sum(min(patches_per_shape, c) for c in data_source.shape_patch_count))

    def __iter__(self):
        if self.identical_epochs:
            self.rng.seed(self.seed)
        counts = list(self.data_source.shape_patch_count)
        offsets = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        order = np.arange(len(counts))
        if not self.sequential_shapes:
            order = self.rng.permutation(order)
        self.shape_patch_inds = [[] for _ in counts]
        out = []
        for shape_ind in order:
            local = self.rng.choice(counts[shape_ind], size=min(self.patches_per_shape, counts[shape_ind]), replace=False)
            self.shape_patch_inds[shape_ind] = local
            out.extend((local + offsets[shape_ind]).tolist())
        return iter(out)

    def __len__(self):
        return self.total_patch_count


class PointcloudPatchDataset(torch.utils.data.Dataset):

    # patch radius as fraction of the bounding box diagonal of a shape
    def __init__(self, root, shape_list_filename, patch_radius, points_per_patch, patch_features,
                 seed=None, identical_epochs=False, use_pca=True, center='point', point_tuple=1, cache_capacity=1,
                 point_count_std=0.0, sparse_patches=False):
        self.root = root
        self.shape_list_filename = shape_list_filename
        self.patch_features = list(patch_features)
        self.patch_radius = [float(r) for r in patch_radius]      # Python floats, as argparse hands them to the reference
        self.points_per_patch = int(points_per_patch)
        self.identical_epochs = identical_epochs
        self.use_pca = use_pca
        self.sparse_patches = sparse_patches
        self.center = center
        self.point_tuple = point_tuple
        self.point_count_std = point_count_std
        self.seed = int(np.random.randint(0, 2 ** 32 - 1)) if seed is None else int(seed)
        self.epoch = 0

        if center not in ('point', 'mean', 'none'):
            raise ValueError('Unknown patch centering option: %s' % (center))
        if use_pca or center != 'point' or point_tuple != 1 or point_count_std > 0:
            raise NotImplementedError(
                "only use_pca=False, center='point', point_tuple=1, point_count_std=0 (what every Nesti-Net "
                "script passes) are on the accelerated path")

        self.include_normals = False
        self.include_curvatures = False
        self.include_noise = False
        for pfeat in self.patch_features:
            if pfeat == 'normal':
                self.include_normals = True
            elif pfeat in ('max_curvature', 'min_curvature'):
                self.include_curvatures = True
            elif pfeat == 'noise':
                self.include_noise = True
            else:
                raise ValueError('Unknown patch feature: %s' % (pfeat))

        with open(os.path.join(root, shape_list_filename)) as f:
            self.shape_names = [line.strip() for line in f if line.strip()]
        noise_file = os.path.join(root, shape_list_filename[:-4] + '_noise_levels.txt')
        if os.path.exists(noise_file):
            with open(noise_file) as f:
                self.noise_levels = [float(line.strip()) for line in f if line.strip()]
        else:
            self.noise_levels = [0.0] * len(self.shape_names)

        self.shape_cache = ShapeCache(cache_capacity, self.load_shape_by_index)
        self.shape_patch_count = []
        self.patch_radius_absolute = []
        for shape_ind in range(len(self.shape_names)):
            shape = self.shape_cache.get(shape_ind)
            self.shape_patch_count.append(shape.pts.shape[0] if shape.pidx is None else len(shape.pidx))
            # bbdiag from the device bbox, evaluated with the reference's own numpy expression
            self.patch_radius_absolute.append(shape.index.absolute_radii(self.patch_radius))
        self._offsets = np.concatenate([[0], np.cumsum(self.shape_patch_count)]).astype(np.int64)

    # ---- subsample seed per epoch -----------------------------------------------------------------
    def begin_epoch(self):
        """Called by the batch loader at the start of every pass.  The reference draws its subsamples from one
        stateful stream, so a patch gets a different 512-point subset in every epoch unless ``identical_epochs``
        reseeds it per item (pcpnet_dataset.py:306-308); here the selection is a pure function of (seed, centre,
        scale), and the epoch enters through the seed instead."""
        if not self.identical_epochs:
            self.epoch += 1

    def selection_seed(self):
        """64-bit seed of the shared seeded selection for the current epoch (epoch 0 and every epoch of an
        ``identical_epochs`` dataset: the constructor's seed)."""
        return (self.seed + self.epoch * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)

    # ---- reference-compatible per-item access (:286-419) ---------------------------------------
    def __getitem__(self, index):
        batch = self.get_batch([int(index)])
        patch_pts = batch[0][0].cpu()
        feats = tuple(t[0].cpu() for t in batch[1:-2])
        trans = batch[-2][0].cpu()
        effective_points_num = batch[-1][0].cpu().numpy().astype(np.float64)
        return (patch_pts,) + feats + (trans,) + (effective_points_num,)

    def __len__(self):
        return int(self._offsets[-1])

    def shape_index(self, index):
        """global patch index -> (shape index, patch index inside the shape) (:427-436)."""
        if index < 0 or index >= self._offsets[-1]:
            raise IndexError(index)
        shape_ind = int(np.searchsorted(self._offsets, index, side='right') - 1)
        return shape_ind, int(index - self._offsets[shape_ind])

    def load_shape_by_index(self, shape_ind):
        base = os.path.join(self.root, self.shape_names[shape_ind])
        pts = np.ascontiguousarray(_load_table(base + '.xyz', 'float32'))
        normals = _load_table(base + '.normals', 'float32') if self.include_normals else None
        curv = _load_table(base + '.curv', 'float32') if self.include_curvatures else None
        pidx = _load_table(base + '.pidx', 'int64') if self.sparse_patches else None
        index = PointIndex(pts, cell_frac=max(self.patch_radius))
        return Shape(pts, index, normals=normals, curv=curv, pidx=pidx, noise_level=self.noise_levels[shape_ind])

    # ---- batched access: one ball-query launch per distinct shape of the batch ------------------------------
    def get_batch(self, indices):
        """[points [B,S*P,3] f32 (CUDA), *targets, trans [B,3,3], n_eff [B,S] (float64 like the
        collated reference item, CUDA)] -- the list a DataLoader batch of the reference holds
        (provider.py:319-429).  The batch is grouped by shape (stable, so the order inside a shape is the caller's):
        ONE ball-query launch per distinct shape whatever the sampling order -- the reference's training order
        ('random', train_n_est_w_experts.py:232) interleaves the shapes patch by patch -- and each shape is fetched from
        the cache once per batch."""
        indices = np.asarray(indices, dtype=np.int64).reshape(-1)
        B, S, P = len(indices), len(self.patch_radius), self.points_per_patch
        if B and (indices.min() < 0 or indices.max() >= self._offsets[-1]):
            raise IndexError("patch index out of range")
        shape_inds = np.searchsorted(self._offsets, indices, side='right') - 1
        dev = torch.device('cuda', torch.cuda.current_device())
        points = torch.empty((B, S * P, 3), dtype=torch.float32, device=dev)
        n_eff = torch.empty((B, S), dtype=torch.int32, device=dev)
        feats = {}
        for pfeat in self.patch_features:
            width = {'normal': 3, 'max_curvature': 1, 'min_curvature': 1}.get(pfeat)
            feats[pfeat] = np.empty((B,) if width is None else (B, width), dtype=np.float64 if width is None else np.float32)
        order = np.argsort(shape_inds, kind='stable')
        cuts = np.flatnonzero(np.diff(shape_inds[order])) + 1
        for rows in np.split(order, cuts):
            if not len(rows):
                continue
            shape_ind = int(shape_inds[rows[0]])
            shape = self.shape_cache.get(shape_ind)
            local = indices[rows] - self._offsets[shape_ind]
            c = local if shape.pidx is None else shape.pidx[local]
            radii = self.patch_radius_absolute[shape_ind]
            p, ne, _ = shape.index.ball_query(c, radii, P, seed=self.selection_seed())
            if len(rows) == B:                              # one shape: the launch wrote the batch in order
                points, n_eff = p, ne
            else:
                at = torch.as_tensor(rows, device=dev)
                points.index_copy_(0, at, p)
                n_eff.index_copy_(0, at, ne)
            for pfeat in self.patch_features:
                if pfeat == 'normal':
                    feats[pfeat][rows] = shape.normals[c, :]
                elif pfeat == 'max_curvature':
                    feats[pfeat][rows] = shape.curv[c, 0:1] * radii[0]
                elif pfeat == 'min_curvature':
                    feats[pfeat][rows] = shape.curv[c, 1:2] * radii[0]
                elif pfeat == 'noise':
                    feats[pfeat][rows] = shape.noise_level
        targets = [torch.as_tensor(feats[pfeat]) for pfeat in self.patch_features]
        trans = torch.eye(3, dtype=torch.float32).repeat(B, 1, 1)
        return [points] + targets + [trans, n_eff.to(torch.float64)]
