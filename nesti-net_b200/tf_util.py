"""Drop-in for the two 3DmFV entry points of the reference's utils/tf_util.py.

Same names, argument meaning and output layout as ``tf_util.get_3dmfv_n_est``
(utils/tf_util.py:655-753, the one every model calls) and ``tf_util.get_3dmfv``
(utils/tf_util.py:578-652), computed by the fused sm_100a kernel instead of a TF1 op chain.
Inputs may be numpy arrays or torch tensors (host inputs are copied to the current CUDA
device, like a feed_dict); the result is a CUDA torch tensor.
"""
from . import mups as _m


def get_3dmfv_n_est(points, w, mu, sigma, flatten=True, n_original_points=None):
    """points [B,P,3], w [G], mu [G,3], sigma [G,3] (std-dev), n_original_points [B]
    -> [B, 20*G] (flatten) or [B, 20, G].  Channel order: pi_max, pi_sum, mu_max xyz,
    mu_min xyz, mu_sum xyz, sigma_max xyz, sigma_min xyz, sigma_sum xyz."""
    if n_original_points is None:
        # tf.cast(None, tf.int32) raises in the reference (tf_util.py:665)
        raise ValueError("n_original_points is required")
    gmm = _m.gmm_handle(w, mu, sigma)
    out = _m.stats_3dmfv(points, n_original_points, gmm, 1, masked=True, layout="channel")
    B = out.shape[0]
    return out.reshape(B, 20 * gmm.G) if flatten else out.reshape(B, 20, gmm.G)


def get_3dmfv(points, w, mu, sigma, flatten=True):
    """The unmasked variant (static point count, MultivariateNormalDiag pdf)."""
    gmm = _m.gmm_handle(w, mu, sigma)
    out = _m.stats_3dmfv(points, None, gmm, 1, masked=False, layout="channel")
    B = out.shape[0]
    return out.reshape(B, 20 * gmm.G) if flatten else out.reshape(B, 20, gmm.G)
