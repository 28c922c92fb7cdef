"""CPU tests: the oracle against the golden fixtures (outputs of the unmodified reference for
half 1, recorded transliteration + float64 cross-check for half 2) and against itself (numpy
restatement vs C port vs float64)."""
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import mups_oracle as orc

TOL_ABS, TOL_REL = 1e-6, 1e-5          # BASELINE.json north_star: 1e-5 relative, 1e-6 absolute in fp32


def frac_outside(got, ref):
    return float((np.abs(got - ref) > TOL_ABS + TOL_REL * np.abs(ref)).mean())


@pytest.fixture(scope="module")
def half1(golden_dir):
    return np.load(os.path.join(golden_dir, "half1_reference.npz"))


@pytest.fixture(scope="module")
def half2(golden_dir):
    return np.load(os.path.join(golden_dir, "half2_oracle.npz"))


# ---- half 1 against the reference's own outputs --------------------------------------------------

@pytest.mark.parametrize("case", ["A", "B"])
def test_half1_matches_reference_fixture(half1, case):
    g = {k[len(case) + 1:]: half1[k] for k in half1.files if k.startswith(case + "_")}
    pts, q, P = g["pts"], g["query_idx"], int(g["P"])
    radius = list(g["patch_radius"])
    S = len(radius)
    # radii: bbdiag * rad, bit-exact (pcpnet_dataset.py:281-282)
    assert abs(orc.bbdiag_of(pts) - float(g["bbdiag"])) < 1e-12   # stored as radii_abs[0] / radius[0]
    assert orc.absolute_radii(pts, radius) == list(g["radii_abs"])
    patches, n_eff, total, nbr = orc.gather_patches(pts, q, radius, P, seed=3627473, return_indices=True)
    assert np.array_equal(n_eff, g["n_eff"])
    off = g["nbr_off"]
    for b in range(len(q)):
        for s in range(S):
            ref_set = g["nbr_flat"][off[b * S + s]:off[b * S + s + 1]]
            assert total[b, s] == len(ref_set)
            # the float64 leaf predicate, independently of the tree
            assert np.array_equal(orc.ball_query_bruteforce(pts, int(q[b]), g["radii_abs"][s]), ref_set)
            mine = patches[b, s * P:(s + 1) * P]
            ref = g["patches"][b, s * P:(s + 1) * P]
            if not g["subsampled"][b, s]:
                assert np.array_equal(nbr[b, s, :len(ref_set)], ref_set)
                assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32)), "patch not bit-exact"
            else:
                # both selections are P-subsets of the same neighbourhood, normalised identically
                cand = ((pts[ref_set] - pts[q[b]]) / np.float32(g["radii_abs"][s])).astype(np.float32)
                cand_rows = {r.tobytes() for r in cand}
                assert all(r.tobytes() in cand_rows for r in ref)
                assert all(r.tobytes() in cand_rows for r in mine)
                sel = nbr[b, s]
                assert len(sel) == P and np.all(np.diff(sel) > 0) and np.isin(sel, ref_set).all()
            assert np.all(mine[n_eff[b, s]:] == 0)


def test_selection_rule_properties():
    nbr = np.random.RandomState(0).choice(100000, 3000, replace=False)
    sel = orc.select_subset(nbr, 512, 3627473, 17, 2)
    assert len(sel) == 512 and np.all(np.diff(sel) > 0) and np.isin(sel, nbr).all()
    assert np.array_equal(sel, orc.select_subset(nbr[::-1], 512, 3627473, 17, 2))     # order independent
    assert not np.array_equal(sel, orc.select_subset(nbr, 512, 3627473, 17, 3))        # scale enters the key
    assert not np.array_equal(sel, orc.select_subset(nbr, 512, 3627474, 17, 2))        # seed enters the key
    keys = orc.selection_keys(3627473, 17, 2, np.sort(nbr)).astype(np.int64)
    kth = np.sort(keys)[511]
    assert np.all(keys[np.isin(np.sort(nbr), sel)] <= kth)
    small = np.arange(40)
    assert np.array_equal(orc.select_subset(small[::-1], 64, 1, 2, 0), small)          # no subsample needed


def test_selection_rule_is_a_uniform_sample():
    """The shared seeded selection draws uniform P-subsets: over 1 500 centres and one fixed neighbour set (random and
    contiguous indices) the per-point selection counts and the pair co-selection counts sit at their hypergeometric
    expectations (chi-square within 5 standard deviations of its degrees of freedom), and the subsets of two scales of the
    same centre overlap like independent draws."""
    n, P, T = 1500, 256, 1500
    for nbr in (np.sort(np.random.RandomState(1).choice(100000, n, replace=False)), np.arange(70000, 70000 + n)):
        cnt, pair, overlap = np.zeros(n), np.zeros((40, 40)), []
        for c in range(T):
            sel = np.zeros(n, bool)
            sel[np.argsort(orc.selection_keys(3627473, c, 2, nbr).astype(np.int64), kind="stable")[:P]] = True
            sel2 = np.zeros(n, bool)
            sel2[np.argsort(orc.selection_keys(3627473, c, 3, nbr).astype(np.int64), kind="stable")[:P]] = True
            cnt += sel
            pair += np.outer(sel[:40], sel[:40])
            overlap.append(int((sel & sel2).sum()))
        p, pp = P / n, P * (P - 1) / (n * (n - 1))
        chi = ((cnt - T * p) ** 2 / (T * p * (1 - p))).sum()
        assert abs(chi - n) < 5 * np.sqrt(2 * n), chi
        off = pair[np.triu_indices(40, 1)]
        chi2 = ((off - T * pp) ** 2 / (T * pp * (1 - pp))).sum()
        assert abs(chi2 - len(off)) < 5 * np.sqrt(2 * len(off)), chi2
        sd = np.sqrt(P * p * (1 - p) * (n - P) / (n - 1))                        # hypergeometric
        assert abs(np.mean(overlap) - P * p) < 5 * sd / np.sqrt(T) and 0.8 * sd < np.std(overlap) < 1.2 * sd


def test_philox_known_answers_and_c_port():
    # Random123 known-answer vectors for philox4x32-10
    assert [int(x) for x in orc.philox4x32_10(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(x) for x in orc.philox4x32_10(f, f, f, f, f, f)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in orc.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    nbr = np.arange(1000, 1200)
    for seed, c, s in [(3627473, 17, 0), (3627473, 17, 3), ((9 << 32) | 5, 99999, 6)]:
        assert np.array_equal(orc.selection_keys(seed, c, s, nbr), c_oracle.selection_keys(seed, c, s, nbr))


# ---- half 2 ------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def half2_ref(golden_dir):
    return np.load(os.path.join(golden_dir, "half2_reference_numpy.npz"))


@pytest.mark.parametrize("n", [3, 8, 16])
def test_grid_gmm_matches_the_reference_run(half2_ref, n):
    """utils/utils.py:70-95 run unmodified in the build container (make_golden.py): bit-identical arrays."""
    w, mu, cov = orc.get_3d_grid_gmm([n] * 3, float(half2_ref["grid%d_variance" % n]))
    assert np.array_equal(w, half2_ref["grid%d_weights" % n])
    assert np.array_equal(mu, half2_ref["grid%d_means" % n])
    assert np.array_equal(cov, half2_ref["grid%d_covariances" % n])


@pytest.mark.parametrize("case", ["g3", "g8", "g8p512"])
def test_half2_stages_pinned_by_reference_numpy(half2_ref, case):
    """The reference's own numpy get_3DmFV (utils/utils.py:260-330), run unmodified, differs from the TF
    path in one stage only (Q = p: no posterior normalisation, no mask).  With that stage switched off the
    transliteration must reproduce the reference's outputs: prefactor, derivative terms, max/min/sum
    reductions, 1/sqrt(w), 1/sqrt(2w), 1/n, signed square root, per-channel L2 norm, channel order."""
    pts, w, mu, sg = (half2_ref["%s_%s" % (case, k)] for k in ("points", "w", "mu", "sigma"))
    B, P, _ = pts.shape
    ref = half2_ref[case + "_fv"]                                   # [B, 20, G] float64
    got = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=np.full(B, P, np.int32),
                              _posterior="pdf")
    assert got.shape == ref.shape
    # the band, except where a sum channel is a near-complete fp32 cancellation (DESIGN.md 'Tolerance'): those
    # few elements must be explained by <= 2e-7 of noise in the unit-norm statistic before the square root
    bad = np.abs(got - ref) > TOL_ABS + TOL_REL * np.abs(ref)
    assert bad.mean() <= 1e-4, (int(bad.sum()), np.abs(got - ref).max())
    du = np.abs(got * np.abs(got) - ref * np.abs(ref))
    assert not bad.any() or du[bad].max() <= 2e-7
    assert np.abs(got - ref).max() < 5e-5
    # flatten=True is the reference layout flattened channel-major (tf_util.py:743-748)
    flat = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=True, n_original_points=np.full(B, P, np.int32),
                               _posterior="pdf")
    assert np.array_equal(flat, got.reshape(B, -1))
    # and the default stays the posterior of tf_util.py:700-701 (rows of Q sum to one -> different numbers)
    assert frac_outside(orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False,
                                            n_original_points=np.full(B, P, np.int32)), ref) > 0.5

@pytest.mark.parametrize("case", ["g3", "g8", "g8p512", "gen"])
def test_half2_fixture_regression_and_c_port(half2, case):
    pts, ne, w, mu, sg = (half2["%s_%s" % (case, k)] for k in ("points", "n_eff", "w", "mu", "sigma"))
    fv = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=ne)
    assert np.allclose(fv, half2[case + "_fv_n_est"], rtol=0, atol=1e-7)
    plain = orc.get_3dmfv(pts, w, mu, sg, flatten=False)
    assert np.allclose(plain, half2[case + "_fv_plain"], rtol=0, atol=1e-7)
    # float64 restatement written independently
    assert np.abs(fv - orc.get_3dmfv_n_est_f64(pts, w, mu, sg, ne, masked=True)).max() < 1e-6
    assert np.abs(plain - orc.get_3dmfv_n_est_f64(pts, w, mu, sg, None, masked=False)).max() < 1e-6
    # C port
    assert frac_outside(c_oracle.get_3dmfv(pts, w, mu, sg, ne, masked=True), fv) == 0.0
    assert frac_outside(c_oracle.get_3dmfv(pts, w, mu, sg, None, masked=False), plain) == 0.0
    # flatten=True is the channel-major flattening of flatten=False
    flat = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=True, n_original_points=ne)
    assert np.array_equal(flat, fv.reshape(len(pts), -1))
    # every channel is L2-normalised over the Gaussians (or identically zero)
    nrm = np.sqrt((fv.astype(np.float64) ** 2).sum(-1))
    assert np.all((np.abs(nrm - 1) < 1e-5) | (nrm == 0))


@pytest.mark.parametrize("case", ["post_g3", "post_g8"])
def test_half2_posterior_stage_pinned_by_reference(half2_ref, case):
    """The reference's fisher_vector_per_point (utils/utils.py:214-245, run unmodified by make_golden.py) evaluates
    the per-point terms with the posterior Q = gmm.predict_proba(x) = w p / sum_g w p (tf_util.py:700-701); reduced
    with the reference's own get_3DmFV tail it is get_3dmfv_n_est with nothing masked: the full restatement -- and
    its C port and float64 twin -- must reproduce it."""
    pts, w, mu, sg = (half2_ref["%s_%s" % (case, k)] for k in ("points", "w", "mu", "sigma"))
    B, P, _ = pts.shape
    ref = half2_ref[case + "_fv"]
    ne = np.full(B, P, np.int32)
    got = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=ne)
    assert frac_outside(got, ref) == 0.0, np.abs(got - ref).max()
    assert frac_outside(c_oracle.get_3dmfv(pts, w, mu, sg, ne, masked=True), ref) == 0.0
    # the float64 twin is fed the float32-rounded GMM the graph gets, the reference run its float64 sklearn object
    assert np.abs(orc.get_3dmfv_n_est_f64(pts, w, mu, sg, ne, masked=True) - ref).max() < 1e-6
    # n_eff = P - 1 masks nothing either (tf_util.py:696 masks r > n_eff) but divides by P - 1: same features, the
    # positive factor cancels in the channel norm
    got2 = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=ne - 1)
    assert frac_outside(got2, ref) == 0.0


def test_half2_properties():
    """Size-independent properties of the statistics the CUDA tests rely on at full size: invariance to the order of
    the unmasked points, every channel L2-normalised over the Gaussians (or identically zero), max channels >= min
    channels, and invariance of the normalised features to the positive factors that cancel in the channel norm
    (n_eff = P - 1 vs P masks nothing, tf_util.py:696)."""
    from hypothesis import given, settings, strategies as st

    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([4] * 3, 0.0625))

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(st.integers(1, 24), st.integers(0, 2 ** 31 - 1))
    def check(ne, seed):
        rng = np.random.RandomState(seed)
        P = 24
        pts = np.zeros((1, P, 3), np.float32)
        x = rng.normal(size=(ne, 3)) * rng.uniform(0.05, 0.6)
        x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
        pts[0, :ne] = x
        fv = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=[ne])[0]        # [20, G]
        perm = rng.permutation(ne)
        p2 = pts.copy()
        p2[0, :ne] = pts[0, :ne][perm]
        fv2 = orc.get_3dmfv_n_est(p2, w, mu, sg, flatten=False, n_original_points=[ne])[0]
        assert np.abs(fv - fv2).max() < 2e-6                                   # only the sums re-associate
        nrm = np.sqrt((fv.astype(np.float64) ** 2).sum(-1))
        assert np.all((np.abs(nrm - 1) < 1e-5) | (nrm == 0))
        # max >= min holds before the per-channel normalisation; after it the signs still cannot cross
        assert not np.any((fv[2:5] < 0) & (fv[5:8] > 0)) and not np.any((fv[11:14] < 0) & (fv[14:17] > 0))
        if ne < P - 1:      # masked slots feed exact zeros into max / min (tf_util.py:698,703)
            assert np.all(fv[[0, 2, 3, 4, 11, 12, 13]] >= 0) and np.all(fv[[5, 6, 7, 14, 15, 16]] <= 0)
    check()


def test_mask_off_by_one_and_padding_semantics():
    """tf_util.py:696 masks r > n_eff: slot n_eff (a zero pad) takes part, slot n_eff+1 does not;
    masked slots feed exact zeros into max/min."""
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([3] * 3, 0.11))
    rng = np.random.RandomState(3)
    P, ne = 16, 6
    pts = np.zeros((1, P, 3), np.float32)
    pts[0, :ne] = rng.uniform(-0.5, 0.5, (ne, 3))
    base = orc.get_3dmfv_n_est(pts, w, mu, sg, n_original_points=[ne])
    p2 = pts.copy(); p2[0, ne + 1] = 0.3
    assert np.array_equal(base, orc.get_3dmfv_n_est(p2, w, mu, sg, n_original_points=[ne]))
    p3 = pts.copy(); p3[0, ne] = 0.3
    assert not np.array_equal(base, orc.get_3dmfv_n_est(p3, w, mu, sg, n_original_points=[ne]))
    fv = base.reshape(20, -1)
    assert np.all(fv[[0, 2, 3, 4, 11, 12, 13]] >= 0) and np.all(fv[[5, 6, 7, 14, 15, 16]] <= 0)
    with pytest.raises(ValueError):
        orc.get_3dmfv_n_est(pts, w, mu, sg)


def test_mups_layout_and_c_port():
    """models/experts_n_est.py:71-76: MuPS[b,i,j,k,s*20+c] = fv_s[b,c,(i*res+j)*res+k]."""
    res, P, S, B = 3, 16, 3, 5
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([res] * 3, 0.11))
    rng = np.random.RandomState(5)
    pts = rng.uniform(-0.6, 0.6, (B, S * P, 3)).astype(np.float32)
    ne = rng.randint(1, P + 1, (B, S)).astype(np.int32)
    for b in range(B):
        for s in range(S):
            pts[b, s * P + ne[b, s]:(s + 1) * P] = 0
    mups = orc.mups_assemble(pts, w, mu, sg, ne, S)
    assert mups.shape == (B, res, res, res, 20 * S)
    for s in range(S):
        fv = orc.get_3dmfv_n_est(pts[:, s * P:(s + 1) * P], w, mu, sg, flatten=False, n_original_points=ne[:, s])
        for c in (0, 7, 19):
            assert np.array_equal(mups[..., s * 20 + c].reshape(B, -1), fv[:, c, :])
    assert frac_outside(c_oracle.mups(pts, ne, w, mu, sg, S), mups) == 0.0
    # the grid: x slowest (np.mgrid), cell centres, sigma = sqrt(variance)
    w8, mu8, sg8 = orc.gmm_feed(*orc.get_3d_grid_gmm([8] * 3, 0.0156))
    assert np.allclose(mu8[0], [-0.875] * 3) and np.allclose(mu8[1], [-0.875, -0.875, -0.625])
    assert np.allclose(mu8[64], [-0.625, -0.875, -0.875]) and np.allclose(sg8, 0.1249, atol=1e-4) and np.allclose(w8, 1 / 512)


def test_c_port_on_realistic_patches():
    """numpy restatement vs C port on patches of a PCPNet-shape cloud.  Sum channels whose value
    is a near-complete fp32 cancellation are ill-conditioned under the signed square root, so a
    ~1e-6 fraction of elements may leave the 1e-5/1e-6 band between ANY two fp32 evaluations
    ('Tolerance'); bound that fraction and the size of the excursions."""
    pts = orc.synthetic_cloud(30000, cloud_id=2)
    radius = [0.01, 0.03, 0.05, 0.07]
    q = np.arange(0, 30000, 2000)
    patches, n_eff, _ = orc.gather_patches(pts, q, radius, 512)
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([8] * 3, 0.0156))
    a = orc.mups_assemble(patches, w, mu, sg, n_eff, 4)
    c = c_oracle.mups(patches, n_eff, w, mu, sg, 4)
    assert frac_outside(c, a) < 2e-5
    assert np.abs(c - a).max() < 1e-4


# ---- half 2 against the reference's own TensorFlow text run on emulated ops ------------------------------------------

@pytest.fixture(scope="module")
def half2_tf(golden_dir):
    return np.load(os.path.join(golden_dir, "half2_tf_emulated.npz"))


@pytest.mark.parametrize("case", ["g3", "g8", "gen"])
def test_half2_pinned_by_reference_text_on_emulated_tf(half2_tf, case):
    """tests/golden/half2_tf_emulated.npz holds outputs of utils/tf_util.py::get_3dmfv_n_est / get_3dmfv EXECUTED FROM
    THE REFERENCE'S SOURCE TEXT with `tf` bound to a numpy emulation of the primitive ops (make_golden.py): it pins the
    n_eff mask (:691-703) and the sigma_0^3 prefactor (anisotropic 'gen' case) -- the one stage no reference numpy code
    covers -- for the transliteration, its C port and the float64 restatement, with n_eff in {1, 2, 3, P/2, P-2, P-1, P}."""
    pts, ne, w, mu, sg = (half2_tf["%s_%s" % (case, k)] for k in ("points", "n_eff", "w", "mu", "sigma"))
    ref = half2_tf[case + "_fv_n_est"]
    assert frac_outside(orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=ne), ref) == 0.0
    assert frac_outside(c_oracle.get_3dmfv(pts, w, mu, sg, ne, masked=True), ref) == 0.0
    assert frac_outside(orc.get_3dmfv_n_est_f64(pts, w, mu, sg, ne), ref) == 0.0
    plain = half2_tf[case + "_fv_plain"]
    assert frac_outside(orc.get_3dmfv(pts, w, mu, sg, flatten=False), plain) == 0.0
    assert frac_outside(c_oracle.get_3dmfv(pts, w, mu, sg, None, masked=False), plain) == 0.0
    # the mask really is `slot > n_eff`: a perturbation of slot n_eff changes the reference's output, slot n_eff + 1 does not
    B, P, _ = pts.shape
    b = int(np.argmax((ne >= 2) & (ne < P - 2)))
    moved = pts.copy()
    moved[b, ne[b] + 1] = 0.3
    assert np.array_equal(orc.get_3dmfv_n_est(moved, w, mu, sg, False, ne)[b], orc.get_3dmfv_n_est(pts, w, mu, sg, False, ne)[b])
    moved[b, ne[b]] = 0.3
    assert not np.array_equal(orc.get_3dmfv_n_est(moved, w, mu, sg, False, ne)[b], orc.get_3dmfv_n_est(pts, w, mu, sg, False, ne)[b])


def test_mups_assembly_pinned_by_reference_text(half2_tf):
    """The MuPS loop of models/experts_n_est.py::get_model (:59-76) executed from the reference's text (py2 integer
    division, reshape [B,-1,res,res,res], transpose [0,2,3,4,1], concat) against mups_assemble and the C port."""
    pts, ne, ref = half2_tf["mups_points"], half2_tf["mups_n_eff"], half2_tf["mups_out"]
    w, mu, sg = half2_tf["g3_w"], half2_tf["g3_mu"], half2_tf["g3_sigma"]
    assert frac_outside(orc.mups_assemble(pts, w, mu, sg, ne, 2), ref) == 0.0
    assert frac_outside(c_oracle.mups(pts, ne, w, mu, sg, 2), ref) == 0.0
    truth, bound = c_oracle.mups_f64(pts, ne, w, mu, sg, 2)
    assert frac_outside(truth, ref) == 0.0 and np.all(bound >= 0) and np.all(np.isfinite(bound))


def test_forward_error_bound_covers_the_fp32_oracles():
    """every element of an fp32 evaluation lies within 1e-5 rel / 1e-6 abs of the float64 value OR
    within the forward-error bound oracle_mups_f64 derives from the data (no tolerated exceptions).  Checked here for the
    two CPU fp32 evaluations (literal port, tuned port) on realistic patches; the GPU tests apply the same criterion."""
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([8] * 3, 0.0156))
    pts = orc.synthetic_cloud(30000, cloud_id=5, noise=0.001)
    q = np.random.RandomState(2).choice(30000, 48, replace=False)
    pa, ne, _ = orc.gather_patches(pts, q, [0.01, 0.03, 0.05, 0.07], 512)
    truth, bound = c_oracle.mups_f64(pa, ne, w, mu, sg, 4)
    band = TOL_ABS + TOL_REL * np.abs(truth)
    for name, got in (("literal", c_oracle.mups(pa, ne, w, mu, sg, 4)), ("tuned", c_oracle.mups_tuned(pa, ne, w, mu, sg, 4))):
        err = np.abs(got - truth)
        assert np.all(err <= np.maximum(band, bound)), name
        outside = err > band
        assert outside.mean() < 1e-4                                   # the bound is only needed for a handful
    # the bound is not a blanket allowance: for the typical element it is below the contract's band
    assert np.median(bound / band) < 1.0


def test_tuned_cpu_path_matches_the_oracle():
    """The CPU legs of bench.py: the C/OpenMP half 1 (selection + gather + normalise after cKDTree) is bit-identical to
    gather_patches, the tuned half 2 agrees with the literal port to 1e-5 -- they time the same computation."""
    pts = orc.synthetic_cloud(20000, cloud_id=6)
    tree = orc.build_kdtree(pts)
    q = np.random.RandomState(4).choice(20000, 40, replace=False).astype(np.int64)
    radius, P = [0.02, 0.05, 0.09], 96
    ref_p, ref_ne, ref_tot = orc.gather_patches(pts, q, radius, P, seed=11, kdtree=tree)
    assert (ref_tot > P).any() and (ref_tot <= P).any()
    patches = np.zeros_like(ref_p)
    n_eff = np.zeros_like(ref_ne)
    for s, rad in enumerate(orc.absolute_radii(pts, radius)):
        lists = tree.query_ball_point(pts[q], rad)
        c_oracle.half1_gather(pts, q, rad, lists, P, len(radius), s, 11, patches, n_eff)
    assert np.array_equal(n_eff, ref_ne) and np.array_equal(patches.view(np.uint32), ref_p.view(np.uint32))
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([8] * 3, 0.0156))
    a = c_oracle.mups(ref_p, ref_ne, w, mu, sg, 3)
    b = c_oracle.mups_tuned(ref_p, ref_ne, w, mu, sg, 3)
    assert np.abs(a - b).max() < 2e-5
