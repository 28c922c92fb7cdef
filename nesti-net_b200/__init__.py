"""nesti-net_b200: B200-native (sm_100a) MuPS hot path of sitzikbs/Nesti-Net.

Half 1: multi-radius ball query + seeded 512-point subsample + per-scale normalisation
(reference utils/pcpnet_dataset.py).  Half 2: 3DmFV statistics over a Gaussian grid in the
[B, res, res, res, 20*S] layout (reference utils/tf_util.py get_3dmfv_n_est +
models/experts_n_est.py).  Everything is computed by hand-written CUDA kernels behind the C ABI of
include/mups.h; there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .mups import (GMMHandle, GridGMM, PointIndex, get_3d_grid_gmm, gmm_handle, mups_features,  # noqa: F401
                   stats_3dmfv, stats_3dmfv_selected)
from . import tf_util, experts_n_est, experts_net, pcpnet_dataset, provider, dist, pipeline  # noqa: F401
from . import evaluate, inference, moe_engine  # noqa: F401
from .pipeline import MuPSPipeline  # noqa: F401

__all__ = ["GMMHandle", "GridGMM", "PointIndex", "get_3d_grid_gmm", "gmm_handle", "mups_features",
           "stats_3dmfv", "stats_3dmfv_selected", "tf_util", "experts_n_est", "pcpnet_dataset", "provider", "dist", "pipeline", "MuPSPipeline"]
