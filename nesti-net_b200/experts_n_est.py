"""Drop-in for the MuPS assembly at the head of the reference's model graphs
(models/experts_n_est.py:59-76; same idiom in ms_norm_est.py:62-80, ms_sw_n_est.py:61-74,
ss_norm_est.py:38-50).  The CNN / Mixture-of-Experts that consumes the tensor is out of scope
(SURVEY.md section 8f)."""
from . import mups as _m


def multi_scale_point_statistics(points, w, mu, sigma, radius, original_n_points):
    """points [B, n_rads*P, 3], original_n_points [B, n_rads] -> MuPS [B,res,res,res,20*n_rads]
    with MuPS[b,i,j,k,s*20+c] = statistic c of Gaussian g=(i*res+j)*res+k at scale s -- the
    tensor get_model returns third and feeds to scale_manager_net / normal_est_net."""
    n_rads = len(radius)
    gmm = _m.gmm_handle(w, mu, sigma)
    return _m.stats_3dmfv(points, original_n_points, gmm, n_rads, masked=True, layout="mups")
