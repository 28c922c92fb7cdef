"""Parity tests proper (need a B200): the CUDA path, called through the C ABI, against the
oracle on the same seeded inputs, against the committed golden fixtures, and -- at
BASELINE.json's full sizes -- through size-independent properties."""
import os
import socket

import numpy as np
import pytest
import torch

import nesti_net_b200 as mb
from nesti_net_b200 import _lib
from oracle import c_oracle
from oracle import mups_oracle as orc

pytestmark = pytest.mark.gpu

TOL_ABS, TOL_REL = 1e-6, 1e-5          # BASELINE.json north_star: 3DmFV within 1e-5 rel / 1e-6 abs (fp32)
SEED = 3627473                          # the reference's constant (test_n_est_w_experts.py:113)


def check_features(got, patches, n_eff, w, mu, sg, S, what="", layout="mups", masked=True, ref32=None):
    """The feature gate, with NO tolerated exceptions: every element of `got` must lie within
    |a-b| <= 1e-6 + 1e-5|b| of the FLOAT64 evaluation of the reference formula (oracle_mups_f64), or -- where the
    formula is ill-conditioned: a sum channel that cancels almost completely in front of the signed square root --
    within the forward-error bound of an fp32 evaluation that oracle_mups_f64 derives from the same data
    (eps (c0 + c1 ss) |term| per pair, eps csum sqrt(m) sum|term| for the accumulation, propagated through /n_eff, the
    square root and the channel norm).  If an fp32 oracle result `ref32` is given, `got` must also
    agree with it within the band or twice the bound (both are fp32 evaluations).  Rows with n_eff == 0 are last-batch
    padding: NaN in the reference, skipped.

    layout: 'mups' [B, ..., 20 S]; 'channel' [B, S, 20, G]; 'fv' [B, 20, G] or [B, 20 G] (one scale).
    Returns (elements outside the band, max err / bound among them)."""
    patches = np.asarray(patches, np.float32)
    B, G = len(patches), len(w)
    truth, bound = c_oracle.mups_f64(patches, n_eff, w, mu, sg, S, masked=masked)
    truth, bound = truth.reshape(B, G, S, 20), bound.reshape(B, G, S, 20)

    def to_bgsc(x):
        x = np.asarray(x, np.float64)
        if layout == "mups":
            return x.reshape(B, G, S, 20)
        if layout == "channel":
            return x.reshape(B, S, 20, G).transpose(0, 3, 1, 2)
        assert layout == "fv" and S == 1
        return x.reshape(B, 20, G).transpose(0, 2, 1)[:, :, None, :]
    got = to_bgsc(got)
    real = np.ones((B, S), bool) if (n_eff is None or not masked) else (np.asarray(n_eff).reshape(B, S) > 0)
    rows = np.broadcast_to(real[:, None, :, None], got.shape)
    assert np.all(np.isfinite(got[rows])), what + ": non-finite feature"
    err = np.where(rows, np.abs(got - np.where(rows, truth, 0)), 0)
    band = TOL_ABS + TOL_REL * np.abs(np.where(rows, truth, 0))
    lim = np.maximum(band, np.where(rows, bound, 0))
    bad = err > lim
    assert not bad.any(), "%s: %d of %d elements outside 1e-5 rel / 1e-6 abs of float64 AND outside the fp32 error bound " \
        "(max err %.3g, worst err/bound %.3g)" % (what, int(bad.sum()), bad.size, err.max(), float((err / np.maximum(lim, 1e-300)).max()))
    outside = err > band
    worst = float((err[outside] / bound[outside]).max()) if outside.any() else 0.0
    if ref32 is not None:
        r = to_bgsc(ref32)
        e2 = np.where(rows, np.abs(got - np.where(rows, r, 0)), 0)
        assert np.all(e2 <= np.maximum(TOL_ABS + TOL_REL * np.abs(np.where(rows, r, 0)), 2 * np.where(rows, bound, 0))), \
            "%s: disagrees with the fp32 oracle beyond band and bound (max %.3g)" % (what, e2.max())
    return int(outside.sum()), worst


@pytest.fixture(scope="module", autouse=True)
def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("-m gpu tests need a CUDA device (B200); there is no CPU fallback to test instead")
    torch.cuda.set_device(0)
    _lib.load()
    yield
    torch.cuda.synchronize()


@pytest.fixture(scope="module")
def half1(golden_dir):
    return np.load(os.path.join(golden_dir, "half1_reference.npz"))


@pytest.fixture(scope="module")
def half2(golden_dir):
    return np.load(os.path.join(golden_dir, "half2_oracle.npz"))


def grid_gmm(res, var):
    return orc.gmm_feed(*orc.get_3d_grid_gmm([res] * 3, var))


@pytest.fixture(params=["auto", "hier"])
def query_kernel(request):
    """Half-1 tests run twice: automatic kernel choice (flat cell scan on coarse grids) and the hierarchical kernel
    forced (mups_set_option "query_kernel" = 2) -- the results may not differ in a single bit."""
    _lib.set_option("query_kernel", 2 if request.param == "hier" else 0)
    yield request.param
    _lib.set_option("query_kernel", 0)


# =====================================================================================================
# half 1: index + ball query + subsample + normalise
# =====================================================================================================

def run_half1(pts, q, radius, P, seed=SEED, cell_frac=None, indices=True):
    index = mb.PointIndex(pts, cell_frac=max(radius) if cell_frac is None else cell_frac)
    radii_abs = index.absolute_radii(radius)
    out = index.ball_query(q, radii_abs, P, seed=seed, return_indices=indices)
    torch.cuda.synchronize()
    return index, radii_abs, [t.cpu().numpy() for t in out]


@pytest.mark.parametrize("case", ["A", "B"])
def test_half1_against_reference_fixture(half1, case, query_kernel):
    """Against outputs of the unmodified reference (tests/golden/make_golden.py)."""
    g = {k[len(case) + 1:]: half1[k] for k in half1.files if k.startswith(case + "_")}
    pts, q, P = g["pts"], g["query_idx"], int(g["P"])
    radius = list(g["patch_radius"])
    S = len(radius)
    index, radii_abs, (patches, n_eff, total, nbr) = run_half1(pts, q, radius, P)
    mn, mx = index.bbox()
    assert np.array_equal(mn, pts.min(0)) and np.array_equal(mx, pts.max(0))
    assert radii_abs == list(g["radii_abs"])                       # bit-exact float64 radii
    assert np.array_equal(n_eff, g["n_eff"])
    off = g["nbr_off"]
    for b in range(len(q)):
        for s in range(S):
            ref_set = g["nbr_flat"][off[b * S + s]:off[b * S + s + 1]]
            assert total[b, s] == len(ref_set)
            sel = nbr[b, s, :n_eff[b, s]]
            assert np.all(nbr[b, s, n_eff[b, s]:] == -1)
            mine = patches[b, s * P:(s + 1) * P]
            if not g["subsampled"][b, s]:
                assert np.array_equal(sel, ref_set), "neighbour set differs from cKDTree"
                assert np.array_equal(mine.view(np.uint32), g["patches"][b, s * P:(s + 1) * P].view(np.uint32))
            else:
                assert np.all(np.diff(sel) > 0) and np.isin(sel, ref_set).all()
    # and the whole thing against the oracle (same shared seeded selection): bit-exact everywhere
    o_patches, o_neff, o_total, o_nbr = orc.gather_patches(pts, q, radius, P, seed=SEED, return_indices=True)
    assert np.array_equal(nbr, o_nbr)
    assert np.array_equal(patches.view(np.uint32), o_patches.view(np.uint32))
    assert np.array_equal(total, o_total) and np.array_equal(n_eff, o_neff)


@pytest.mark.parametrize("n,P,radius,kind,seed", [
    (20000, 512, [0.01, 0.03, 0.05, 0.07], "pcpnet", SEED),
    (20000, 512, [0.01, 0.03, 0.07], "pcpnet", SEED),                 # BASELINE configs[0] scale set
    (50000, 256, [0.07, 0.02, 0.05], "scan", 12345),                   # radii not ascending, non-uniform density
    (8000, 64, [0.02, 0.04, 0.06, 0.08, 0.1, 0.12, 0.14, 0.2], "pcpnet", (7 << 32) | 99),   # 8 scales, 64-bit seed
    (30000, 1024, [0.05, 0.1], "pcpnet", 1),
    (40000, 128, [0.015, 0.35, 0.2], "pcpnet", 77),                     # > 4096 neighbours: the hit list overflows, re-scan path
])
def test_half1_against_oracle(n, P, radius, kind, seed, query_kernel):
    pts = orc.synthetic_cloud(n, cloud_id=4, kind=kind, noise=0.002)
    q = np.random.RandomState(1).choice(n, 96, replace=False)
    _, _, (patches, n_eff, total, nbr) = run_half1(pts, q, radius, P, seed=seed)
    o_patches, o_neff, o_total, o_nbr = orc.gather_patches(pts, q, radius, P, seed=seed, return_indices=True)
    assert np.array_equal(total, o_total), "neighbour counts differ from cKDTree"
    assert np.array_equal(n_eff, o_neff)
    assert np.array_equal(nbr, o_nbr), "selected indices differ"
    assert np.array_equal(patches.view(np.uint32), o_patches.view(np.uint32)), "patches not bit-exact"
    assert (total > P).any() and (total <= P).any()      # both branches exercised


def test_half1_grid_cell_scale():
    """Dense clouds get fine grid cells (mb.mups.grid_cell_scale) and with them the hierarchical kernel; the former
    1/2 and 1/3 cells with the flat kernel forced are the re-scan regime of round 1: same neighbours, same patches."""
    assert mb.mups.grid_cell_scale(100000) == 1.0 and mb.mups.grid_cell_scale(450000) == 0.125
    assert mb.mups.grid_cell_scale(10000000) == 0.0625
    radius, P = [0.01, 0.03, 0.07], 256
    try:
        for n, scale, kernel in ((450000, None, 0), (450000, 0.5, 1), (60000, 0.34, 1), (60000, 0.5, 1), (60000, 0.125, 0),
                                 (60000, 0.0625, 0), (60000, 0.03, 2)):
            _lib.set_option("query_kernel", kernel)
            pts = orc.synthetic_cloud(n, cloud_id=11, kind="scan" if n > 100000 else "pcpnet", noise=0.001)
            q = np.random.RandomState(5).choice(n, 24, replace=False)
            index = mb.PointIndex(pts, cell_frac=max(radius), cell_scale=scale)
            patches, n_eff, total, nbr = index.ball_query(torch.from_numpy(q).cuda(), index.absolute_radii(radius), P,
                                                          seed=SEED, return_indices=True)
            o_patches, o_neff, o_total, o_nbr = orc.gather_patches(pts, q, radius, P, seed=SEED, return_indices=True)
            assert np.array_equal(total.cpu().numpy(), o_total), (n, scale)
            assert np.array_equal(nbr.cpu().numpy(), o_nbr), (n, scale)
            assert np.array_equal(patches.cpu().numpy().view(np.uint32), o_patches.view(np.uint32)), (n, scale)
            assert np.array_equal(n_eff.cpu().numpy(), o_neff)
    finally:
        _lib.set_option("query_kernel", 0)


@pytest.mark.parametrize("n,kind,noise,P,radius,scale,order", [
    (300000, "pcpnet", 0.0, 512, [0.01, 0.03, 0.05, 0.07], 0.125, 0),
    (300000, "scan", 0.0005, 256, [0.07, 0.02, 0.05], 0.0625, 2),            # radii not ascending, 8 Morton bits, CTAs reordered
    (200000, "pcpnet", 0.01, 1024, [0.05, 0.1], 0.125, 0),                    # thick noisy shell: many occupied cells
    (100000, "pcpnet", 0.0, 64, [0.01, 0.02, 0.04, 0.06, 0.08, 0.1, 0.12, 0.2], 0.125, 2),   # 8 scales
    (150000, "scan", 0.0, 512, [0.3], 0.0625, 0),                             # one huge ball: bulk nodes far above the leaves
    (50000, "pcpnet", 0.0, 2048, [0.05, 0.2], 0.125, 0),                      # lists too large for the hierarchical kernel: flat
])
def test_half1_hierarchical_kernel(n, kind, noise, P, radius, scale, order):
    """The hierarchical kernel (octree descent over the Morton-ordered cells, whole-cell acceptance, one-pass key
    threshold) against cKDTree + the shared seeded selection: counts, selected indices and patches bit-exact; also
    with the hand-over to the flat kernel forced for about half of the balls (hier_margin = 0)."""
    pts = orc.synthetic_cloud(n, cloud_id=14, kind=kind, noise=noise)
    q = np.random.RandomState(9).choice(n, 128, replace=False)
    ref = orc.gather_patches(pts, q, radius, P, seed=SEED, return_indices=True)
    try:
        _lib.set_option("query_kernel", 2)
        _lib.set_option("query_order", order)
        for margin in (-1, 0):
            _lib.set_option("hier_margin", margin)
            index = mb.PointIndex(pts, cell_frac=max(radius), cell_scale=scale)
            out = index.ball_query(q, index.absolute_radii(radius), P, seed=SEED, return_indices=True)
            patches, n_eff, total, nbr = [t.cpu().numpy() for t in out]
            assert np.array_equal(total, ref[2]), "neighbour counts differ from cKDTree (margin %d)" % margin
            assert np.array_equal(n_eff, ref[1])
            assert np.array_equal(nbr, ref[3]), "selected indices differ (margin %d)" % margin
            assert np.array_equal(patches.view(np.uint32), ref[0].view(np.uint32)), "patches not bit-exact"
    finally:
        _lib.set_option("query_kernel", 0)
        _lib.set_option("query_order", 0)
        _lib.set_option("hier_margin", -1)
    if P <= 1024:
        assert (ref[2] > P).any()


def test_half1_edge_cases(query_kernel):
    rng = np.random.RandomState(2)
    # single point, two identical points, tiny cloud
    for pts in (np.zeros((1, 3), np.float32) + 0.5,
                np.ones((2, 3), np.float32),
                rng.normal(size=(5, 3)).astype(np.float32)):
        n = len(pts)
        q = np.arange(n)
        index = mb.PointIndex(pts, cell_frac=0.07)
        diag = index.bbdiag()
        radii_abs = [diag * 0.07 if diag > 0 else 0.1, diag * 2.0 if diag > 0 else 1.0]
        patches, n_eff, total, nbr = [t.cpu().numpy() for t in index.ball_query(q, radii_abs, 8, return_indices=True)]
        kd = orc.build_kdtree(pts)
        for b in range(n):
            for s, r in enumerate(radii_abs):
                ref = orc.ball_query(kd, pts, b, r)
                assert total[b, s] == len(ref) and np.array_equal(nbr[b, s, :len(ref)], ref)
    # heavy duplicates + query on the bbox corners + radius covering the whole cloud + radius 0
    pts = orc.synthetic_cloud(3000, cloud_id=6)
    pts[100:400] = pts[100]
    corners = [int(pts[:, k].argmax()) for k in range(3)] + [int(pts[:, k].argmin()) for k in range(3)]
    q = np.array(corners + [100, 250, 0, 2999])
    index = mb.PointIndex(pts, cell_frac=0.05)
    diag = index.bbdiag()
    radii_abs = [0.0, diag * 0.05, diag * 0.3, diag * 1.5]
    patches, n_eff, total, nbr = [t.cpu().numpy() for t in index.ball_query(q, radii_abs, 128, seed=5, return_indices=True)]
    kd = orc.build_kdtree(pts)
    for b, c in enumerate(q):
        for s, r in enumerate(radii_abs):
            ref = orc.ball_query(kd, pts, int(c), r)
            assert total[b, s] == len(ref), (b, s)
            exp = orc.select_subset(ref, 128, 5, int(c), s)
            assert np.array_equal(nbr[b, s, :len(exp)], exp), (b, s)
    assert total[:, 3].min() == 3000 and total[6, 0] == 300       # whole cloud; 300 duplicates at radius 0
    # invalid centre index -> empty row, total -1 (no exception across the ABI)
    p2, ne2, tot2 = [t.cpu().numpy() for t in index.ball_query(np.array([5, -1, 3000]), radii_abs[1:2], 16)]
    assert ne2[0, 0] >= 1 and list(ne2[1:, 0]) == [0, 0] and list(tot2[1:, 0]) == [-1, -1] and np.all(p2[1:] == 0)
    with pytest.raises(ValueError):
        index.ball_query(q, [0.1] * 9, 16)
    with pytest.raises(ValueError):
        index.ball_query(q, [float("nan")], 16)
    with pytest.raises(ValueError):
        index.ball_query(q, [0.1], 4096)


def test_half1_cell_size_independent_and_refinement_levels(query_kernel):
    """Results do not depend on the grid resolution, on the query order, or on the size of the
    radix threshold group (boundary_cap lowered so the refinement levels run)."""
    pts = orc.synthetic_cloud(40000, cloud_id=7, noise=0.001)
    radius = [0.02, 0.06, 0.1]
    q = np.random.RandomState(3).choice(40000, 64, replace=False)
    ref = orc.gather_patches(pts, q, radius, 128, seed=SEED, return_indices=True)
    for frac in (0.1, 0.03, 0.25, 0.004):
        _, _, out = run_half1(pts, q, radius, 128, cell_frac=frac)
        for a, b in zip(out, ref):
            assert np.array_equal(a, b), "cell_frac=%g" % frac
    try:
        for cap in (1, 3, 17):
            _lib.set_option("boundary_cap", cap)
            _, _, out = run_half1(pts, q, radius, 128, cell_frac=0.1)
            for a, b in zip(out, ref):
                assert np.array_equal(a, b), "boundary_cap=%d" % cap
        # the same with balls too large for the shared-memory hit list (every pass re-scans the cells)
        big = [0.05, 0.3]
        ref_big = orc.gather_patches(pts, q[:24], big, 128, seed=SEED, return_indices=True)
        assert ref_big[2].max() > 4096
        for cap in (2, 256):
            _lib.set_option("boundary_cap", cap)
            _, _, out = run_half1(pts, q[:24], big, 128)
            for a, b in zip(out, ref_big):
                assert np.array_equal(a, b), "re-scan path, boundary_cap=%d" % cap
        # ... and with the key histograms built in the first scan for every ball / for none
        _lib.set_option("boundary_cap", 512)
        for fuse in (0, 2 ** 31 - 1):
            _lib.set_option("fuse_candidates", fuse)
            _, _, out = run_half1(pts, q, radius, 128, cell_frac=0.1)
            for a, b in zip(out, ref):
                assert np.array_equal(a, b), "fuse_candidates=%d" % fuse
            _, _, out = run_half1(pts, q[:24], big, 128)
            for a, b in zip(out, ref_big):
                assert np.array_equal(a, b), "re-scan path, fuse_candidates=%d" % fuse
    finally:
        _lib.set_option("boundary_cap", 512)
        _lib.set_option("fuse_candidates", 12288)
    perm = np.random.RandomState(4).permutation(len(q))
    _, _, out = run_half1(pts, q[perm], radius, 128, cell_frac=0.1)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b[perm])


# =====================================================================================================
# half 2: 3DmFV statistics
# =====================================================================================================

@pytest.mark.parametrize("fastpath", [True, False])
@pytest.mark.parametrize("case", ["g3", "g8", "g8p512", "gen"])
def test_half2_against_golden_fixture(half2, case, fastpath):
    pts, ne, w, mu, sg = (half2["%s_%s" % (case, k)] for k in ("points", "n_eff", "w", "mu", "sigma"))
    gmm = mb.gmm_handle(w, mu, sg)
    assert gmm.separable == (case != "gen")
    got = mb.stats_3dmfv(pts, ne, gmm, 1, masked=True, layout="channel", fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(got, pts, ne, w, mu, sg, 1, case + " n_est", layout="fv", ref32=half2[case + "_fv_n_est"])
    got = mb.stats_3dmfv(pts, None, gmm, 1, masked=False, layout="channel", fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(got, pts, None, w, mu, sg, 1, case + " plain", layout="fv", masked=False, ref32=half2[case + "_fv_plain"])


@pytest.mark.parametrize("fastpath", [True, False])
@pytest.mark.parametrize("case", ["post_g3", "post_g8"])
def test_half2_against_reference_run_fixture(golden_dir, case, fastpath):
    """CUDA kernels against outputs of the REFERENCE's own numpy code run in the build container (posterior from
    sklearn predict_proba through utils/utils.py::fisher_vector_per_point, reductions and normalisations of its
    get_3DmFV; tests/golden/make_golden.py): the path with nothing masked."""
    ref = np.load(os.path.join(golden_dir, "half2_reference_numpy.npz"))
    pts, w, mu, sg = (ref["%s_%s" % (case, k)] for k in ("points", "w", "mu", "sigma"))
    B, P, _ = pts.shape
    gmm = mb.gmm_handle(w, mu, sg)
    got = mb.stats_3dmfv(pts, np.full((B, 1), P, np.int32), gmm, 1, masked=True, layout="channel",
                         fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(got, pts, np.full((B, 1), P, np.int32), w, mu, sg, 1, case + " vs reference run", layout="fv",
                   ref32=ref[case + "_fv"])
    plain = mb.stats_3dmfv(pts, None, gmm, 1, masked=False, layout="channel", fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(plain, pts, None, w, mu, sg, 1, case + " get_3dmfv vs reference run", layout="fv", masked=False,
                   ref32=ref[case + "_fv"])                                                    # isotropic sigma: same pdf


@pytest.mark.parametrize("fastpath", [True, False])
@pytest.mark.parametrize("case", ["g3", "g8", "gen"])
def test_half2_against_reference_text_on_emulated_tf(golden_dir, case, fastpath):
    """Both CUDA kernels against tests/golden/half2_tf_emulated.npz: outputs of the reference's OWN TensorFlow source
    text (utils/tf_util.py::get_3dmfv_n_est / get_3dmfv) executed with the primitive ops emulated in numpy -- the one
    fixture that pins the n_eff mask stage and the anisotropic prefactor (n_eff in {1, 2, 3, P/2, P-2, P-1, P})."""
    ref = np.load(os.path.join(golden_dir, "half2_tf_emulated.npz"))
    pts, ne, w, mu, sg = (ref["%s_%s" % (case, k)] for k in ("points", "n_eff", "w", "mu", "sigma"))
    gmm = mb.gmm_handle(w, mu, sg)
    got = mb.stats_3dmfv(pts, ne, gmm, 1, masked=True, layout="channel", fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(got, pts, ne, w, mu, sg, 1, case + " n_est vs reference text", layout="fv", ref32=ref[case + "_fv_n_est"])
    got = mb.stats_3dmfv(pts, None, gmm, 1, masked=False, layout="channel", fastpath=fastpath).cpu().numpy()[:, 0]
    check_features(got, pts, None, w, mu, sg, 1, case + " plain vs reference text", layout="fv", masked=False,
                   ref32=ref[case + "_fv_plain"])
    if case == "g3":       # MuPS assembly (experts_n_est.py:59-76, also run from the reference's text)
        p2, n2 = ref["mups_points"], ref["mups_n_eff"]
        got = mb.stats_3dmfv(p2, n2, gmm, 2, fastpath=fastpath).cpu().numpy()
        assert got.shape == ref["mups_out"].shape
        check_features(got, p2, n2, w, mu, sg, 2, "MuPS vs reference text", ref32=ref["mups_out"])


def test_half2_reference_signatures(half2):
    """tf_util.get_3dmfv_n_est / get_3dmfv drop-ins: names, argument meaning, flatten, errors."""
    pts, ne, w, mu, sg = (half2["g8_%s" % k] for k in ("points", "n_eff", "w", "mu", "sigma"))
    B, G = len(pts), len(w)
    flat = mb.tf_util.get_3dmfv_n_est(pts, w, mu, sg, flatten=True, n_original_points=ne)
    assert flat.shape == (B, 20 * G) and flat.is_cuda and flat.dtype == torch.float32
    cube = mb.tf_util.get_3dmfv_n_est(torch.from_numpy(pts).cuda(), w, mu, sg, flatten=False,
                                      n_original_points=torch.from_numpy(ne.astype(np.uint16).astype(np.int32)))
    assert cube.shape == (B, 20, G) and torch.equal(cube.reshape(B, -1), flat)
    check_features(flat.cpu().numpy(), pts, ne, w, mu, sg, 1, "flatten=True", layout="fv",
                   ref32=orc.get_3dmfv_n_est(pts, w, mu, sg, True, ne))
    plain = mb.tf_util.get_3dmfv(pts, w, mu, sg, flatten=False)
    check_features(plain.cpu().numpy(), pts, None, w, mu, sg, 1, "get_3dmfv", layout="fv", masked=False,
                   ref32=orc.get_3dmfv(pts, w, mu, sg, flatten=False))
    with pytest.raises(ValueError):
        mb.tf_util.get_3dmfv_n_est(pts, w, mu, sg)                   # the reference fails on None too
    with pytest.raises(ValueError):
        mb.tf_util.get_3dmfv(pts[:, :, :2], w, mu, sg)
    with pytest.raises(ValueError):
        mb.gmm_handle(w, mu, -sg)


@pytest.mark.parametrize("fastpath", [True, False])
@pytest.mark.parametrize("res,P,S,var", [(8, 512, 4, 0.0156), (8, 512, 3, 0.0156), (3, 64, 2, 0.11), (5, 100, 1, 0.04),
                                         (8, 256, 4, 0.0156), (8, 1024, 2, 0.0156), (16, 128, 2, 0.00390625),
                                         (4, 96, 2, 0.0625), (8, 130, 1, 0.0156), (8, 129, 2, 0.0156), (32, 32, 1, 0.0009765625)])
def test_half2_mups_layout_against_oracle(res, P, S, var, fastpath):
    """Edge cases of n_eff (1, 2, P-2, P-1, P, all-zero channel) in the [B,res,res,res,20*S] layout."""
    w, mu, sg = grid_gmm(res, var)
    rng = np.random.RandomState(res * 1000 + P)
    edge = [1, 2, 3, P // 2, P - 2, P - 1, P]
    B = 12
    ne = np.array([[edge[(b + s) % len(edge)] for s in range(S)] for b in range(B)], np.int32)
    ne[-1] = rng.randint(1, P + 1, S)
    pts = np.zeros((B, S * P, 3), np.float32)
    for b in range(B):
        for s in range(S):
            x = rng.normal(size=(ne[b, s], 3)) * rng.uniform(0.1, 0.6)
            x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
            x[0] = 0
            pts[b, s * P: s * P + ne[b, s]] = x
    got = mb.experts_n_est.multi_scale_point_statistics(pts, w, mu, sg, [0.1] * S, ne) if fastpath else \
        mb.stats_3dmfv(pts, ne, mb.gmm_handle(w, mu, sg), S, fastpath=False)
    assert tuple(got.shape) == (B, res, res, res, 20 * S)
    ref = c_oracle.mups(pts, ne, w, mu, sg, S) if res >= 8 else orc.mups_assemble(pts, w, mu, sg, ne, S)
    check_features(got.cpu().numpy(), pts, ne, w, mu, sg, S, "mups res=%d P=%d S=%d" % (res, P, S), ref32=ref)
    # channel layout is the same numbers transposed
    ch = mb.stats_3dmfv(pts, ne, mb.gmm_handle(w, mu, sg), S, layout="channel", fastpath=fastpath).cpu().numpy()
    assert np.array_equal(ch.transpose(0, 3, 1, 2).reshape(B, res, res, res, 20 * S), got.cpu().numpy())


@pytest.mark.parametrize("variant,res,var", [(1, 8, 0.0156), (2, 8, 0.0156), (3, 8, 0.0156), (1, 4, 0.0625), (8, 16, 0.00390625),
                                             (1, 16, 0.00390625)])
def test_half2_kernel_variants(variant, res, var):
    """The non-default statistics kernels kept for A/B measurements (mups_set_option "stats_variant": 1 = round-1
    loop and staging, 2 = all-scalar loop, 3 = two items per CTA with cp.async.bulk + mbarrier patch prefetch, 8 = 16^3
    without the cluster) compute the same features."""
    w, mu, sg = grid_gmm(res, var)
    rng = np.random.RandomState(variant * 100 + res)
    P, S, B = 200, 2, 10
    ne = rng.randint(1, P + 1, (B, S)).astype(np.int32)
    ne[0] = [P, 1]
    pts = np.zeros((B, S * P, 3), np.float32)
    for b in range(B):
        for s in range(S):
            x = rng.normal(size=(ne[b, s], 3)) * 0.4
            x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
            x[0] = 0
            pts[b, s * P: s * P + ne[b, s]] = x
    gmm = mb.gmm_handle(w, mu, sg)
    base = mb.stats_3dmfv(pts, ne, gmm, S).cpu().numpy()
    try:
        _lib.set_option("stats_variant", variant)
        got = mb.stats_3dmfv(pts, ne, gmm, S).cpu().numpy()
        if variant == 3:        # same arithmetic, different staging of the patch: bit-identical, also for an odd item count
            assert np.array_equal(got, base)
            odd_pts, odd_ne = np.ascontiguousarray(pts[:3, :P]), np.ascontiguousarray(ne[:3, :1])      # 3 items: the last CTA has one
            odd = mb.stats_3dmfv(odd_pts, odd_ne, gmm, 1).cpu().numpy()
            _lib.set_option("stats_variant", 0)
            assert np.array_equal(odd, mb.stats_3dmfv(odd_pts, odd_ne, gmm, 1).cpu().numpy())
    finally:
        _lib.set_option("stats_variant", 0)
    ref = c_oracle.mups(pts, ne, w, mu, sg, S)
    check_features(got, pts, ne, w, mu, sg, S, "variant %d res %d" % (variant, res), ref32=ref)
    check_features(base, pts, ne, w, mu, sg, S, "default res %d" % res, ref32=ref)


def test_half2_general_gmm_and_padding_rows():
    rng = np.random.RandomState(9)
    G, P, S, B = 200, 96, 2, 16
    w = rng.uniform(0.5, 1.5, G); w = (w / w.sum()).astype(np.float32)
    mu = rng.uniform(-0.9, 0.9, (G, 3)).astype(np.float32)
    sg = rng.uniform(0.15, 0.5, (G, 3)).astype(np.float32)
    ne = rng.randint(1, P + 1, (B, S)).astype(np.int32)
    pts = rng.uniform(-0.8, 0.8, (B, S * P, 3)).astype(np.float32)      # garbage beyond n_eff on purpose:
    got = mb.stats_3dmfv(pts, ne, mb.gmm_handle(w, mu, sg), S).cpu().numpy()   # slot n_eff takes part, the rest not
    ref = c_oracle.mups(pts, ne, w, mu, sg, S).reshape(got.shape)
    check_features(got, pts, ne, w, mu, sg, S, "general gmm", ref32=ref)
    p2 = pts.copy()
    for b in range(B):
        for s in range(S):
            p2[b, s * P + ne[b, s] + 1:(s + 1) * P] = 7.0
    assert np.array_equal(mb.stats_3dmfv(p2, ne, mb.gmm_handle(w, mu, sg), S).cpu().numpy(), got)
    # last-batch zero padding rows (n_eff = 0) are NaN/Inf in the reference; only the real rows must agree
    ne0 = ne.copy(); ne0[-3:] = 0
    p0 = pts.copy(); p0[-3:] = 0
    got0 = mb.stats_3dmfv(p0, ne0, mb.gmm_handle(w, mu, sg), S).cpu().numpy()
    assert np.array_equal(got0[:-3], got[:-3])


# =====================================================================================================
# both halves, drop-in dataset, full-size properties
# =====================================================================================================

def test_end_to_end_against_oracle():
    pts = orc.synthetic_cloud(50000, cloud_id=8, noise=0.001)
    radius = [0.01, 0.03, 0.05, 0.07]
    P = 512
    w, mu, sg = grid_gmm(8, 0.0156)
    q = np.random.RandomState(5).choice(50000, 256, replace=False)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    radii_abs = index.absolute_radii(radius)
    feats, patches, n_eff, total = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q, radii_abs, P, seed=SEED, return_patches=True)
    o_patches, o_neff, o_total = orc.gather_patches(pts, q, radius, P, seed=SEED)
    assert np.array_equal(total.cpu().numpy(), o_total) and np.array_equal(n_eff.cpu().numpy(), o_neff)
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), o_patches.view(np.uint32))
    n_out, worst = check_features(feats.cpu().numpy(), o_patches, o_neff, w, mu, sg, 4, "end to end",
                                  ref32=c_oracle.mups(o_patches, o_neff, w, mu, sg, 4))
    print("end to end: %d of %d elements outside the 1e-5/1e-6 band of float64 (worst err/bound %.3g)" % (n_out, feats.numel(), worst))


@pytest.mark.parametrize("res,var,P,kernel", [(8, 0.0156, 512, 0), (8, 0.0156, 96, 2), (16, 0.00390625, 64, 0), (4, 0.0625, 130, 0)])
def test_k6_selection_handoff_is_bit_identical(res, var, P, kernel):
    """K6 (mups_features with patches_dev = NULL; mups_ball_query_select + mups_3dmfv_selected): the ball query hands
    over positions instead of a patch tensor and the statistics kernel gathers / centres / normalises while staging.
    Same features bit for bit as the two-launch path through the patch tensor, for both statistics kernels, both
    ball-query kernels, the cluster kernel (16^3) and a row with an invalid centre index."""
    n = 40000
    radius = [0.01, 0.03, 0.05, 0.07]
    pts = orc.synthetic_cloud(n, cloud_id=17, noise=0.001)
    w, mu, sg = grid_gmm(res, var)
    gmm = mb.gmm_handle(w, mu, sg)
    q = np.random.RandomState(3).choice(n, 200, replace=False).astype(np.int64)
    try:
        _lib.set_option("query_kernel", kernel)
        index = mb.PointIndex(pts, cell_frac=max(radius), cell_scale=0.125 if kernel == 2 else None)
        radii = index.absolute_radii(radius)
        ref, patches, n_eff, total = mb.mups_features(index, gmm, q, radii, P, seed=SEED, return_patches=True)
        fused = mb.mups_features(index, gmm, q, radii, P, seed=SEED)
        assert torch.equal(fused, ref)
        pos, ne2, tot2 = index.select(q, radii, P, seed=SEED)
        assert torch.equal(ne2, n_eff) and torch.equal(tot2, total)
        assert bool(((pos >= 0).sum(-1) == n_eff).all()) and bool((pos[..., 0] >= 0).all())
        for fastpath in (True, False):
            two = mb.stats_3dmfv(patches, n_eff, gmm, 4, fastpath=fastpath)
            assert torch.equal(mb.stats_3dmfv_selected(index, gmm, q, radii, pos, n_eff, fastpath=fastpath), two), fastpath
        assert torch.equal(mb.mups_features(index, gmm, q, radii, P, seed=SEED, fastpath=False),
                           mb.stats_3dmfv(patches, n_eff, gmm, 4, fastpath=False))
        # an invalid centre leaves an empty row (n_eff = 0: NaN features like the reference's padding rows); the others are untouched
        qbad = q.copy()
        qbad[5] = -1
        fb = mb.mups_features(index, gmm, qbad, radii, P, seed=SEED)
        keep = np.arange(len(q)) != 5
        assert torch.equal(fb[torch.from_numpy(keep).cuda()], ref[torch.from_numpy(keep).cuda()])
    finally:
        _lib.set_option("query_kernel", 0)


def test_dataset_drop_in(tmp_path, half1):
    """PointcloudPatchDataset / get_data_loader keep the reference's interface and values."""
    g = {k[2:]: half1[k] for k in half1.files if k.startswith("A_")}
    pts, q, P = g["pts"], g["query_idx"], int(g["P"])
    radius = list(g["patch_radius"])
    S = len(radius)
    np.savetxt(tmp_path / "cloud.xyz", pts, fmt="%.9g")
    np.savetxt(tmp_path / "cloud.normals", np.tile([0.0, 0.0, 1.0], (len(pts), 1)))
    (tmp_path / "list.txt").write_text("cloud\n")
    ds = mb.pcpnet_dataset.PointcloudPatchDataset(
        root=str(tmp_path), shape_list_filename="list.txt", patch_radius=radius, points_per_patch=P,
        patch_features=["normal"], seed=SEED, identical_epochs=False, use_pca=False, center="point",
        point_tuple=1, cache_capacity=100, point_count_std=0, sparse_patches=False)
    assert ds.shape_names == ["cloud"] and ds.shape_patch_count == [len(pts)] and len(ds) == len(pts)
    assert ds.patch_radius_absolute[0] == list(g["radii_abs"])
    o_patches, o_neff, _ = orc.gather_patches(pts, q, radius, P, seed=SEED)
    for b in (0, 5, len(q) - 1):
        item = ds[int(q[b])]
        patch_pts, normal, trans, ne = item
        assert isinstance(patch_pts, torch.Tensor) and patch_pts.dtype == torch.float32 and tuple(patch_pts.shape) == (S * P, 3)
        assert tuple(trans.shape) == (3, 3) and torch.equal(trans, torch.eye(3))
        assert np.array_equal(ne, g["n_eff"][b].astype(np.float64)) and tuple(normal.shape) == (3,)
        assert np.array_equal(patch_pts.numpy(), o_patches[b])
    loader, dataset = mb.provider.get_data_loader(
        dataset_name="list.txt", batchSize=64, indir=str(tmp_path), patch_radius=radius, points_per_patch=P,
        outputs=[], patch_point_count_std=0, seed=SEED, identical_epochs=False, use_pca=False, patch_center="point",
        point_tuple=1, cache_capacity=100, patch_sample_order="full", workers=0, dataset_type="test", sparse_patches=False)
    assert len(loader) == (len(pts) + 63) // 64
    first = next(iter(loader))
    points, trans, ne = first
    assert tuple(points.shape) == (64, S * P, 3) and tuple(trans.shape) == (64, 3, 3) and tuple(ne.shape) == (64, S)
    ref_p, ref_ne, _ = orc.gather_patches(pts, np.arange(64), radius, P, seed=SEED)
    assert np.array_equal(points.cpu().numpy(), ref_p) and np.array_equal(ne.cpu().numpy(), ref_ne.astype(np.float64))
    # epochs: a second pass over the loader draws fresh subsamples (the reference's stateful stream does), the
    # oracle reproduces them from the epoch's seed; identical_epochs keeps the first pass's subsamples
    second = next(iter(loader))
    assert dataset.epoch == 1 and dataset.selection_seed() != SEED
    ref_p2, ref_ne2, _ = orc.gather_patches(pts, np.arange(64), radius, P, seed=dataset.selection_seed())
    assert np.array_equal(second[0].cpu().numpy(), ref_p2) and np.array_equal(second[-1].cpu().numpy(), ref_ne2.astype(np.float64))
    assert (ref_ne > P - 1).any() and not np.array_equal(ref_p2, ref_p)            # some patch was subsampled differently
    assert np.array_equal(second[-1].cpu().numpy(), ne.cpu().numpy())               # counts do not depend on the draw
    ds_same = mb.pcpnet_dataset.PointcloudPatchDataset(
        root=str(tmp_path), shape_list_filename="list.txt", patch_radius=radius, points_per_patch=P,
        patch_features=[], seed=SEED, identical_epochs=True, use_pca=False, center="point",
        point_tuple=1, cache_capacity=100, point_count_std=0, sparse_patches=False)
    ds_same.begin_epoch()
    assert ds_same.epoch == 0 and ds_same.selection_seed() == SEED


def test_dataset_interface_against_reference_run(tmp_path, golden_dir):
    """Two shapes, sparse patch centres (.pidx), normal / curvature targets, global index -> (shape, patch): the mirror
    against outputs of the unmodified reference dataset (tests/golden/make_golden.py::make_dataset_interface)."""
    ref = np.load(os.path.join(golden_dir, "dataset_reference.npz"))
    radius, P = list(ref["patch_radius"]), int(ref["P"])
    names = ["shape_a", "shape_b"]
    for name in names:
        np.savetxt(tmp_path / (name + ".xyz"), ref[name + "_pts"], fmt="%.9g")
        np.savetxt(tmp_path / (name + ".normals"), ref[name + "_normals"], fmt="%.9g")
        np.savetxt(tmp_path / (name + ".curv"), ref[name + "_curv"], fmt="%.9g")
        np.savetxt(tmp_path / (name + ".pidx"), ref[name + "_pidx"], fmt="%d")
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    ds = mb.pcpnet_dataset.PointcloudPatchDataset(
        root=str(tmp_path), shape_list_filename="list.txt", patch_radius=radius, points_per_patch=P,
        patch_features=["normal", "max_curvature", "min_curvature"], seed=SEED, identical_epochs=False,
        use_pca=False, center="point", point_tuple=1, cache_capacity=100, point_count_std=0, sparse_patches=True)
    assert ds.shape_patch_count == list(ref["shape_patch_count"]) and len(ds) == int(ref["length"])
    assert np.array_equal(np.asarray(ds.patch_radius_absolute, np.float64), ref["patch_radius_absolute"])
    for k, gi in enumerate(ref["item_index"]):
        assert tuple(ds.shape_index(int(gi))) == tuple(ref["item_shape_index"][k])
        patch_pts, normal, cmax, cmin, trans, ne = ds[int(gi)]
        assert normal.dtype == torch.float32 and np.array_equal(normal.numpy(), ref["item_normal"][k])
        assert cmax.dtype == torch.float32 and np.array_equal(cmax.numpy(), ref["item_max_curv"][k])
        assert np.array_equal(cmin.numpy(), ref["item_min_curv"][k])
        assert np.array_equal(trans.numpy(), ref["item_trans"][k])
        assert np.asarray(ne).dtype == np.float64 and np.array_equal(ne, ref["item_n_eff"][k])
        # the patch itself against the oracle at the sparse centre
        shape_ind, patch_ind = ref["item_shape_index"][k]
        centre = int(ref[names[shape_ind] + "_pidx"][patch_ind])
        o_p, o_ne, _ = orc.gather_patches(ref[names[shape_ind] + "_pts"], np.array([centre]), radius, P, seed=SEED)
        assert np.array_equal(patch_pts.numpy(), o_p[0]) and np.array_equal(ne, o_ne[0].astype(np.float64))
    with pytest.raises(IndexError):
        ds.shape_index(len(ds))
    batch = ds.get_batch([0, 41, 89])                     # one call across both shapes
    assert np.array_equal(batch[1].numpy(), ref["item_normal"][[0, 5, 7]])
    assert np.array_equal(batch[-1].cpu().numpy(), ref["item_n_eff"][[0, 5, 7]])
    # the training order ('random') interleaves the shapes patch by patch: still ONE ball-query launch per distinct shape,
    # rows delivered in the caller's order
    n = len(ds)
    mixed = [0, n - 1, 1, n - 2, 2, n - 3, 3]
    singles = [ds.get_batch([i]) for i in mixed]
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    inter = ds.get_batch(mixed)
    assert _lib.launch_count() - l0 == 2, "expected one ball-query launch per distinct shape"
    for k, one in enumerate(singles):
        assert torch.equal(inter[0][k], one[0][0]) and torch.equal(inter[-1][k], one[-1][0])
        for a, b in zip(inter[1:-2], one[1:-2]):
            assert torch.equal(a[k], b[0])


def test_full_size_properties():
    """BASELINE configs[1] shape (100k-point cloud, 4 scales, P=512, 8^3 grid): size-independent
    properties on 8192 of the queries + oracle spot check."""
    n, P = 100000, 512
    radius = [0.01, 0.03, 0.05, 0.07]
    pts = orc.synthetic_cloud(n, cloud_id=0)
    w, mu, sg = grid_gmm(8, 0.0156)
    gmm = mb.gmm_handle(w, mu, sg)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    radii_abs = index.absolute_radii(radius)
    q = np.random.RandomState(6).choice(n, 8192, replace=False)
    feats, patches, n_eff, total = mb.mups_features(index, gmm, q, radii_abs, P, seed=SEED, return_patches=True)
    f = feats.cpu().numpy().reshape(len(q), 512, 4, 20)
    ne, tot, pa = n_eff.cpu().numpy(), total.cpu().numpy(), patches.cpu().numpy().reshape(len(q), 4, P, 3)
    # counts: n_eff = min(P, total); balls are nested; the centre is its own neighbour
    assert np.array_equal(ne, np.minimum(tot, P)) and np.all(tot >= 1) and np.all(np.diff(tot, axis=1) >= 0)
    # patches live in the unit ball, contain the centre (an exact zero row inside n_eff), zero padded
    nrm = np.linalg.norm(pa.astype(np.float64), axis=-1)
    assert nrm.max() <= 1.0 + 1e-6
    slot = np.arange(P)[None, None, :]
    assert np.all(pa[slot.repeat(len(q), 0).repeat(4, 1) >= ne[..., None]] == 0)
    has_centre = ((nrm == 0) & (slot < ne[..., None])).sum(-1) >= 1
    assert np.all(has_centre[tot <= P])          # a subsampled patch may drop the centre, like rng.choice can
    # every one of the 20*S channels is L2-normalised over the Gaussians
    cn = np.sqrt((f.astype(np.float64) ** 2).sum(1))
    assert np.all(np.isfinite(f)) and np.all((np.abs(cn - 1) < 1e-5) | (cn == 0))
    # max channels >= 0 >= min channels whenever slots are masked (zeros enter the reductions)
    masked = ne < P - 1
    assert np.all(f[:, :, :, [0, 2, 3, 4, 11, 12, 13]].transpose(0, 2, 1, 3)[masked] >= 0)
    assert np.all(f[:, :, :, [5, 6, 7, 14, 15, 16]].transpose(0, 2, 1, 3)[masked] <= 0)
    # deterministic, independent of batch split and query order
    again = mb.mups_features(index, gmm, q, radii_abs, P, seed=SEED)
    assert torch.equal(again, feats)
    halves = torch.cat([mb.mups_features(index, gmm, q[:3000], radii_abs, P, seed=SEED),
                        mb.mups_features(index, gmm, q[3000:], radii_abs, P, seed=SEED)])
    assert torch.equal(halves, feats)
    perm = np.random.RandomState(7).permutation(len(q))
    assert torch.equal(mb.mups_features(index, gmm, q[perm], radii_abs, P, seed=SEED), feats[torch.from_numpy(perm).cuda()])
    # general (non-separable) kernel agrees with the fast path
    slow = mb.mups_features(index, gmm, q[:1024], radii_abs, P, seed=SEED, fastpath=False)
    check_features(slow[:256].cpu().numpy(), pa.reshape(len(q), 4 * P, 3)[:256], ne[:256], w, mu, sg, 4, "general kernel at full size",
                   ref32=feats[:256].cpu().numpy())
    # oracle spot check at full cloud size
    sub = np.arange(0, len(q), 64)
    o_patches, o_neff, o_total = orc.gather_patches(pts, q[sub], radius, P, seed=SEED)
    assert np.array_equal(tot[sub], o_total) and np.array_equal(pa[sub].reshape(len(sub), 4 * P, 3), o_patches)
    check_features(feats.cpu().numpy()[sub], o_patches, o_neff, w, mu, sg, 4, "full size",
                   ref32=c_oracle.mups(o_patches, o_neff, w, mu, sg, 4))


# =====================================================================================================
# BASELINE configs[3] / configs[4]: the dense clouds at their full sizes
# =====================================================================================================

def _dense_half1_check(pts, kd, q, radius, P, what, **index_kw):
    """Half 1 of `q` on the dense cloud against cKDTree + the shared seeded selection: everything bit-exact."""
    index = mb.PointIndex(pts, cell_frac=max(radius), **index_kw)
    out = index.ball_query(q, index.absolute_radii(radius), P, seed=SEED, return_indices=True)
    patches, n_eff, total, nbr = [t.cpu().numpy() for t in out]
    o_patches, o_neff, o_total, o_nbr = orc.gather_patches(pts, q, radius, P, seed=SEED, kdtree=kd, return_indices=True)
    assert np.array_equal(total, o_total), what + ": neighbour counts differ from cKDTree"
    assert np.array_equal(n_eff, o_neff) and np.array_equal(nbr, o_nbr), what + ": selected indices differ"
    assert np.array_equal(patches.view(np.uint32), o_patches.view(np.uint32)), what + ": patches not bit-exact"
    return patches, n_eff, total


def test_config_c4_dense_scan_cloud_full_size():
    """BASELINE configs[3]: 2 M-point scan-shape cloud (density spread > 20x), 4 scales, P = 512, 8^3 grid.  256 queries
    against the oracle through the hierarchical kernel the library picks for this size (fine grid, 1/8-radius cells),
    the flat kernel on round 1's grids (1/2 and 1/3-radius cells: hit list overflows, every pass re-scans) on 48 of
    them; features of all 256 through the gate of check_features."""
    n, P, radius = 2000000, 512, [0.01, 0.03, 0.05, 0.07]
    pts = orc.synthetic_cloud(n, cloud_id=1, kind="scan")
    kd = orc.build_kdtree(pts)
    q = np.random.RandomState(21).choice(n, 256, replace=False)
    assert mb.mups.grid_cell_scale(n) == 0.125
    patches, n_eff, total = _dense_half1_check(pts, kd, q, radius, P, "hierarchical, automatic grid")
    assert total.max() > 100000 and total[:, 0].min() < total[:, 0].max() / 20        # dense balls, non-uniform density
    try:
        _lib.set_option("query_kernel", 1)
        for scale in (0.5, 0.34):
            _dense_half1_check(pts, kd, q[:48], radius, P, "flat kernel, cell scale %g" % scale, cell_scale=scale)
    finally:
        _lib.set_option("query_kernel", 0)
    w, mu, sg = grid_gmm(8, 0.0156)
    feats = mb.stats_3dmfv(patches, n_eff, mb.gmm_handle(w, mu, sg), 4)
    n_out, worst = check_features(feats.cpu().numpy(), patches, n_eff, w, mu, sg, 4, "configs[3] features",
                                  ref32=c_oracle.mups(patches, n_eff, w, mu, sg, 4))
    print("configs[3]: 256 queries, max ball %d neighbours; %d of %d feature elements outside the 1e-5/1e-6 band of float64 "
          "(worst err/bound %.3g)" % (int(total.max()), n_out, feats.numel(), worst))
    # work-balanced contiguous sharding of the sweep-ordered query list
    order = np.argsort(pts[q, 2], kind="stable")
    work = total[order].sum(1).astype(np.float64)
    b = mb.dist.shard_bounds(len(q), 4, work)
    loads = np.array([work[b[r]:b[r + 1]].sum() for r in range(4)])
    even = mb.dist.shard_bounds(len(q), 4)
    loads_even = np.array([work[even[r]:even[r + 1]].sum() for r in range(4)])
    assert loads.max() / loads.mean() < loads_even.max() / loads_even.mean() and loads.max() / loads.mean() < 1.15


@pytest.fixture(scope="module")
def cloud_10m():
    n = 10000000
    pts = orc.synthetic_cloud(n, cloud_id=2)
    return pts, orc.build_kdtree(pts), np.random.RandomState(22).choice(n, 256, replace=False)


def test_config_c5_10m_points_half1(cloud_10m):
    """BASELINE configs[4]: 10 M points, 4 scales.  256 queries at P = 512 and 64 each at P = 256 / 1024 through the
    hierarchical kernel on the automatic grid (1/16-radius cells, 8 Morton bits); the flat kernel (round 1's 1/3-radius
    cells) on 32 queries including the largest ball, where the 9-bit radix threshold group exceeds boundary_cap = 512
    and the refinement levels run at their natural size."""
    pts, kd, q = cloud_10m
    radius = [0.01, 0.03, 0.05, 0.07]
    assert mb.mups.grid_cell_scale(len(pts)) == 0.0625
    _, _, total = _dense_half1_check(pts, kd, q, radius, 512, "P=512")
    assert total.max() > 250000 and (total[:, 1:] > 10 * 512).all() and (total[:, 3] > 100000).all()
    _dense_half1_check(pts, kd, q[:64], radius, 256, "P=256")
    _dense_half1_check(pts, kd, q[64:128], radius, 1024, "P=1024")
    big = np.concatenate([[int(np.argmax(total[:, 3]))], np.arange(31)])
    try:
        _lib.set_option("query_kernel", 1)
        _dense_half1_check(pts, kd, q[big], radius, 512, "flat kernel, cell scale 0.34", cell_scale=0.34)
    finally:
        _lib.set_option("query_kernel", 0)


@pytest.mark.parametrize("res,P,nq", [(8, 512, 256), (8, 256, 64), (8, 1024, 64), (16, 512, 64)])
def test_config_c5_10m_points_features(cloud_10m, res, P, nq):
    """configs[4] sweep: grid 8^3 / 16^3, P = 256 / 512 / 1024 on the 10 M-point cloud; every patch is full
    (n_eff = P at all four scales).  Both halves on the GPU, features through the gate of check_features."""
    pts, kd, q = cloud_10m
    radius = [0.01, 0.03, 0.05, 0.07]
    w, mu, sg = grid_gmm(res, 0.0156 if res == 8 else (1.0 / res) ** 2)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    feats, patches, n_eff, total = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q[:nq], index.absolute_radii(radius), P,
                                                    seed=SEED, return_patches=True)
    o_patches, o_neff, o_total = orc.gather_patches(pts, q[:nq], radius, P, seed=SEED, kdtree=kd)
    assert np.array_equal(total.cpu().numpy(), o_total) and np.array_equal(n_eff.cpu().numpy(), o_neff)
    assert np.array_equal(patches.cpu().numpy().view(np.uint32), o_patches.view(np.uint32))
    assert (o_neff == P).all()
    n_out, worst = check_features(feats.cpu().numpy(), o_patches, o_neff, w, mu, sg, 4, "configs[4] %d^3 P=%d" % (res, P),
                                  ref32=c_oracle.mups(o_patches, o_neff, w, mu, sg, 4))
    print("configs[4] %d^3 P=%d: %d queries, %d of %d elements outside the band (worst err/bound %.3g)"
          % (res, P, nq, n_out, feats.numel(), worst))


def test_smoke_entry_point():
    import __graft_entry__ as ge
    ge.smoke()


def test_sharded_slabs_and_host_pipeline():
    """Section 8e: per-rank slabs of a sharded query list concatenate to the single-GPU result bit for bit
    (here the ranks run one after the other on one device), with and without work-balanced cuts; and the
    host-buffer streaming API (MuPSPipeline.features_to_host) returns the same rows."""
    n, P = 6000, 128
    radius = [0.03, 0.08]
    pts = orc.synthetic_cloud(n, cloud_id=13, kind="scan")
    gmm = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    handle = mb.gmm_handle(gmm.weights_, gmm.means_, np.sqrt(gmm.covariances_))
    index = mb.PointIndex(pts, cell_frac=max(radius))
    radii = index.absolute_radii(radius)
    q = np.arange(n, dtype=np.int64)
    full, _, _, total = mb.mups_features(index, handle, q, radii, P, seed=SEED, return_patches=True)
    work = total.cpu().numpy()[:, -1].astype(np.float64)            # neighbour count at the largest radius
    for world, weights in ((4, None), (3, work)):
        slabs = []
        for rank in range(world):
            mine, lo, hi = mb.dist.shard_queries(q, rank, world, weights)
            slabs.append(mb.mups_features(index, handle, mine, radii, P, seed=SEED))
        assert torch.equal(torch.cat(slabs), full)
    bounds = mb.dist.shard_bounds(n, 3, work)
    loads = [work[bounds[r]:bounds[r + 1]].sum() for r in range(3)]
    assert max(loads) / (work.sum() / 3) < 1.05                    # balanced by estimated work, not by count
    pipe = mb.MuPSPipeline(gmm, radius, P, seed=SEED, chunk=1024)
    got = np.zeros((n, 20 * 2 * 512), np.float32)
    seen = []

    def consume(lo, hi, rows):
        got[lo:hi] = rows
        seen.append((lo, hi))
    assert pipe.features_to_host(torch.from_numpy(pts).pin_memory(), None, consume) == n
    assert sorted(seen) == [(lo, min(lo + 1024, n)) for lo in range(0, n, 1024)]
    assert np.array_equal(got, full.cpu().numpy().reshape(n, -1))
    assert pipe.d2h_bytes == got.nbytes and pipe.h2d_bytes == pts.nbytes
    dev = pipe.features_on_device(pts, q[:100])
    assert torch.equal(dev, full[:100])
    # consumer on the device (the reference's flow: MuPS feeds the CNN and never visits the host)
    kept = torch.zeros((n, 20 * 2 * 512), dtype=torch.float32, device="cuda")
    sub = np.arange(5, n, 3, dtype=np.int64)
    chunks = []

    def on_device(lo, hi, rows):
        assert rows.is_cuda and tuple(rows.shape) == (hi - lo, 20 * 2 * 512)
        kept[lo:hi].copy_(rows)
        chunks.append((lo, hi))
    assert pipe.features_to_consumer(torch.from_numpy(pts), sub, on_device) == len(sub)
    assert chunks[0] == (0, 1024) and chunks[-1][1] == len(sub)
    assert torch.equal(kept[:len(sub)], full.reshape(n, -1)[torch.from_numpy(sub).cuda()])


def test_wide_stores_and_peer_slab_gather():
    """MUPS_FLAG_WIDE_STORES (results written as 80-byte runs, for `out` in a peer GPU's memory) changes no bit, and
    dist.PeerSlabGather -- the statistics kernel storing its slab straight into the consumer rank's buffer -- returns
    the single-GPU tensor (world size 1 here; profiles/bench_peer_gather.py is the 2-GPU run)."""
    import torch.distributed as dist
    n, P = 5000, 96
    radius = [0.04, 0.09]
    pts = orc.synthetic_cloud(n, cloud_id=21)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    radii = index.absolute_radii(radius)
    q = np.arange(0, n, 2, dtype=np.int64)
    patches, n_eff, _ = index.ball_query(q, radii, P, seed=SEED)
    for res, var in ((8, 0.0156), (4, 0.0625), (16, 0.00390625)):
        w, mu, sg = grid_gmm(res, var)
        gmm = mb.gmm_handle(w, mu, sg)
        base = mb.stats_3dmfv(patches, n_eff, gmm, 2)
        assert torch.equal(mb.stats_3dmfv(patches, n_eff, gmm, 2, wide_stores=True), base), res
        ch = mb.stats_3dmfv(patches, n_eff, gmm, 2, layout="channel")
        assert torch.equal(mb.stats_3dmfv(patches, n_eff, gmm, 2, layout="channel", wide_stores=True), ch)
    w, mu, sg = grid_gmm(8, 0.0156)
    gmm = mb.gmm_handle(w, mu, sg)
    full = mb.stats_3dmfv(patches, n_eff, gmm, 2)
    port = socket.socket()
    port.bind(("127.0.0.1", 0))
    addr = "tcp://127.0.0.1:%d" % port.getsockname()[1]
    port.close()
    dist.init_process_group("nccl", init_method=addr, rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        try:
            gather = mb.dist.PeerSlabGather(len(q), (8, 8, 8, 40), dst=0)
        except RuntimeError as e:
            pytest.skip("peer memory is not available here: %s" % e)
        half = len(q) // 2
        for lo, hi in ((0, half), (half, len(q))):
            mb.stats_3dmfv(patches[lo:hi], n_eff[lo:hi], gmm, 2, out=gather.target(lo, hi), wide_stores=True)
        gather.finish()
        assert torch.equal(gather.result(), full)
    finally:
        dist.destroy_process_group()


def test_inference_driver(tmp_path):
    """.xyz list -> .normals / .experts / .experts_probs (test_n_est_w_experts.py:108-197) on a 3^3 grid."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator
    from nesti_net_b200.inference import estimate_normals
    radius, P = [0.05, 0.1], 64
    names = ["shape_a", "shape_b"]
    clouds = [orc.synthetic_cloud(700, cloud_id=11), orc.synthetic_cloud(500, cloud_id=12)]
    for name, pts in zip(names, clouds):
        np.savetxt(tmp_path / (name + ".xyz"), pts, fmt="%.9g")
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    gmm = mb.get_3d_grid_gmm([3, 3, 3], 0.11)
    torch.manual_seed(5)
    model = ExpertsNormalEstimator(n_rads=2, n_gaussians=27, n_experts=3).eval()        # checker-sized, on the host
    out = estimate_normals(str(tmp_path), "list.txt", str(tmp_path / "results"), model, gmm, radius, P, batch_size=256)
    assert list(out) == names
    w, mu, sg = orc.gmm_feed(gmm.weights_, gmm.means_, gmm.covariances_)
    for name, pts in zip(names, clouds):
        normals, experts, probs = out[name]
        assert normals.shape == (len(pts), 3) and experts.shape == (len(pts),) and probs.shape == (len(pts), 3)
        assert np.allclose(np.loadtxt(tmp_path / "results" / (name + ".normals")), normals, atol=1e-7)
        assert np.array_equal(np.loadtxt(tmp_path / "results" / (name + ".experts")).astype(int), experts)
        assert np.allclose(np.loadtxt(tmp_path / "results" / (name + ".experts_probs")).sum(1), 1.0, atol=1e-5)
        # the same network on oracle MuPS of the same shape
        o_patches, o_neff, _ = orc.gather_patches(pts, np.arange(len(pts)), radius, P, seed=SEED)
        ref_mups = torch.from_numpy(orc.mups_assemble(o_patches, w, mu, sg, o_neff, 2))
        ref_n, ref_e, _ = model.predict(ref_mups)
        same = ref_e.numpy() == experts
        assert same.mean() > 0.99
        assert mb.experts_net.angular_rms_deg(torch.from_numpy(normals[same]), ref_n[same]) < 1e-3
        assert mb.evaluate.evaluate_shape(normals, ref_n.numpy())["pgp5"] > 0.99


def test_cloud_normal_estimator_matches_the_dataset_driver(tmp_path):
    """inference.CloudNormalEstimator (no patch tensor, MuPS consumed on the device by the tensor-core engine) returns what
    inference.estimate_normals (reference-shaped dataset loop, patches materialised) returns for the same cloud."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator
    from nesti_net_b200.inference import CloudNormalEstimator, estimate_normals
    from nesti_net_b200.moe_engine import TensorCoreExperts
    radius, P = [0.01, 0.03, 0.05, 0.07], 512
    pts = orc.synthetic_cloud(3000, cloud_id=21)
    np.savetxt(tmp_path / "shape.xyz", pts, fmt="%.9g")
    (tmp_path / "list.txt").write_text("shape\n")
    pts = np.loadtxt(tmp_path / "shape.xyz").astype(np.float32)      # what the dataset will read back
    gmm = mb.get_3d_grid_gmm([8, 8, 8], 0.0156)
    torch.manual_seed(7)
    tc = TensorCoreExperts(ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval().cuda())
    ref = estimate_normals(str(tmp_path), "list.txt", str(tmp_path / "out"), tc, gmm, radius, P, batch_size=700, write=False)["shape"]
    got = CloudNormalEstimator(tc, gmm, radius, P, seed=SEED, chunk=1024)(pts)
    assert got[0].shape == (3000, 3) and got[1].shape == (3000,) and got[2].shape == (3000, 7)
    assert np.array_equal(got[1], ref[1])
    assert np.allclose(got[0], ref[0], atol=1e-6) and np.allclose(got[2], ref[2], atol=1e-6)
    sub = np.arange(5, 3000, 7)
    part = CloudNormalEstimator(tc, gmm, radius, P, seed=SEED)(pts, sub)
    assert np.array_equal(part[1], ref[1][sub]) and np.allclose(part[0], ref[0][sub], atol=1e-6)


def _conv_reference(x, cin_off, cin, w, scale, shift, relu, k):
    """fp32 reference of mups_conv3d_bn_relu on the bf16-rounded operands: TF 'SAME' cross-correlation, scale / shift, ReLU."""
    import torch.nn.functional as F
    xs = x[..., cin_off:cin_off + cin].float()
    if xs.ndim == 2:
        xs = xs[:, None, None, None, :]
    v = xs.permute(0, 4, 1, 2, 3)
    a, b = (k - 1) // 2, (k - 1) - (k - 1) // 2
    kk = w.float().reshape(k, k, k, w.shape[1], w.shape[2]).permute(3, 4, 0, 1, 2)[:, :cin]      # [Cout, Cin, kd, kh, kw]
    y = F.conv3d(F.pad(v, (a, b, a, b, a, b)), kk)
    y = y * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 4, 1)


@pytest.mark.parametrize("D,k,ct,cin_off,cin,cout,B", [
    (8, 1, 128, 0, 128, 128, 3), (8, 3, 128, 0, 128, 64, 2), (8, 5, 64, 0, 64, 64, 2), (8, 3, 128, 32, 32, 32, 2),
    (8, 5, 256, 0, 256, 128, 1), (4, 2, 768, 0, 768, 256, 5), (4, 4, 512, 0, 512, 256, 3), (2, 2, 512, 0, 512, 256, 37),
    (2, 1, 1536, 0, 1536, 512, 20), (1, 1, 1536, 0, 1536, 1024, 300), (1, 1, 128, 0, 128, 16, 130), (8, 3, 96, 0, 96, 16, 1),
    (8, 2, 64, 0, 64, 32, 3), (8, 4, 128, 64, 64, 48, 2), (8, 5, 384, 0, 384, 256, 2), (8, 1, 384, 0, 384, 256, 40),
    (8, 3, 256, 0, 256, 128, 3), (8, 5, 136, 8, 128, 128, 2),
    (8, 3, 128, 0, 128, 128, 600), (8, 4, 64, 0, 64, 128, 593),
    (8, 1, 128, 0, 128, 128, 160), (8, 1, 72, 8, 64, 512, 149), (4, 1, 256, 0, 256, 256, 1301),
    (8, 3, 64, 0, 64, 128, 150), (8, 3, 96, 0, 96, 64, 151), (8, 5, 136, 8, 128, 128, 149)])
def test_tcgen05_conv3d_against_torch(D, k, ct, cin_off, cin, cout, B):
    """mups_conv3d_bn_relu (tcgen05 / TMEM / TMA implicit GEMM, csrc/moe_conv.cu) against torch conv3d in fp32 on the same
    bf16-rounded operands: every volume edge and kernel edge of the reference's networks, 'SAME' padding for even kernels,
    channel-slice inputs and outputs, batch tails of the 128-row tile, the fully connected layers (D = 1), fp32 output."""
    from nesti_net_b200 import moe_engine as me
    torch.manual_seed(D * 100 + k * 10 + cout)
    dev = torch.device("cuda", 0)
    shape = (B, D, D, D, ct) if D > 1 else (B, ct)
    x = torch.randn(shape, device=dev).to(torch.bfloat16)
    layer = me.PackedConv.__new__(me.PackedConv)
    layer.k, layer.relu, layer.cout, layer.cout_pad, layer.cin_pad = k, True, cout, cout, cin
    layer.w = (torch.randn((k ** 3, cout, cin), device=dev) / np.sqrt(cin * k ** 3)).to(torch.bfloat16).contiguous()
    layer.scale = torch.rand(cout, device=dev) + 0.5
    layer.shift = torch.randn(cout, device=dev) * 0.1
    out = torch.full(shape[:-1] + (cout + 16,), 7.0, dtype=torch.bfloat16, device=dev)      # the kernel writes a channel slice
    f32 = torch.empty((x.numel() // ct, cout), dtype=torch.float32, device=dev)
    me.conv3d_bn_relu(x, cin_off, cin, layer, out, 8, f32)
    torch.cuda.synchronize()
    ref = _conv_reference(x, cin_off, cin, layer.w, layer.scale, layer.shift, True, k).reshape(-1, cout)
    err = (f32 - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), "fp32 output: max err %.3g" % err
    got = out.reshape(-1, cout + 16)
    assert torch.all(got[:, :8] == 7.0) and torch.all(got[:, 8 + cout:] == 7.0), "wrote outside its channel slice"
    assert torch.equal(got[:, 8:8 + cout], f32.to(torch.bfloat16)), "bf16 output is not the rounded fp32 output"
    # bf16x3 output straight from the epilogue (mups_conv3d_bn_relu_x3): one triplet, and two triplets split at 16 / at cout - 16 --
    # bit-identical to splitting the fp32 output (hi = bf16(v), lo = bf16(v - hi))
    hi = f32.to(torch.bfloat16)
    lo = (f32 - hi.float()).to(torch.bfloat16)
    for split in sorted({cout, 16, max(16, cout - 16)}):
        out3 = torch.full(shape[:-1] + (3 * cout + 16,), 7.0, dtype=torch.bfloat16, device=dev)
        me.conv3d_x3(x, cin_off, cin, layer, out3, 8, split)
        torch.cuda.synchronize()
        g3 = out3.reshape(-1, 3 * cout + 16)
        assert torch.all(g3[:, :8] == 7.0) and torch.all(g3[:, 8 + 3 * cout:] == 7.0), "wrote outside its triplets"
        off = 8
        for c0, c1 in ((0, split), (split, cout)):
            w3 = c1 - c0
            if w3 == 0:
                continue
            assert torch.equal(g3[:, off:off + w3], hi[:, c0:c1]) and torch.equal(g3[:, off + 2 * w3:off + 3 * w3], hi[:, c0:c1])
            assert torch.equal(g3[:, off + w3:off + 2 * w3], lo[:, c0:c1])
            off += 3 * w3
    if k == 1 and B >= 149:
        # >= 592 voxel tiles: conv_variant 8 = CTA pairs that share every weight tile through TMA multicast (odd tile counts get an
        # all-out-of-bounds partner): same MMAs in the same order
        _lib.set_option("conv_variant", 8)
        try:
            f32b = torch.empty_like(f32)
            outb = torch.full_like(out, 7.0)
            me.conv3d_bn_relu(x, cin_off, cin, layer, outb, 8, f32b)
            torch.cuda.synchronize()
        finally:
            _lib.set_option("conv_variant", 0)
        assert torch.equal(f32b, f32) and torch.equal(outb, out)
    if D == 8 and k > 1 and cout <= 128 and B >= 148:
        # conv_variant 9: the z-halo kernel as CTA pairs (clusters of two) that share every weight tile through TMA multicast --
        # whole-sample CTAs of two samples (an odd batch gets an all-out-of-bounds partner), or the two halves of one sample;
        # same MMAs in the same order
        _lib.set_option("conv_variant", 9)
        try:
            f32p = torch.empty_like(f32)
            outp = torch.full_like(out, 7.0)
            me.conv3d_bn_relu(x, cin_off, cin, layer, outp, 8, f32p)
            torch.cuda.synchronize()
        finally:
            _lib.set_option("conv_variant", 0)
        assert torch.equal(f32p, f32) and torch.equal(outp, out)
    if D == 8 and k > 1 and cout <= 128:
        # the default above was the z-halo kernel (one activation box per (dy, dx, channel block) serves all dz taps);
        # conv_variant 2 forces the per-tap kernel: same products, another summation order
        # (cout = 128: with the operand roles swapped, one N = 256 MMA over both voxel tiles; conv_variant 3 = unswapped)
        # (batches >= 592: one CTA per whole sample, two N = 256 accumulators; conv_variant 4 = half-sample CTAs)
        for variant in (2, 3, 4):
            _lib.set_option("conv_variant", variant)
            try:
                f32b = torch.empty_like(f32)
                me.conv3d_bn_relu(x, cin_off, cin, layer, None, 0, f32b)
                torch.cuda.synchronize()
            finally:
                _lib.set_option("conv_variant", 0)
            assert (f32b - ref).abs().max().item() < 2e-3 * max(1.0, ref.abs().max().item())
            assert (f32b - f32).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("D,k,is_max", [(8, 3, False), (4, 2, False), (8, 2, False), (8, 4, False), (8, 5, False), (4, 3, False),
                                        (8, 2, True), (4, 2, True), (2, 2, True)])
def test_pool3d_against_tf_semantics(D, k, is_max):
    """mups_pool3d against experts_net.avg_pool_same / max_pool_same (TF 'SAME': the average counts only valid cells) on a
    channel slice of a bf16 NDHWC tensor; the 8^3 average pools run the shared-memory tile kernel (separable box sum) and
    are also compared with the per-voxel kernel (pool_variant 1)."""
    from nesti_net_b200 import moe_engine as me
    from nesti_net_b200.experts_net import avg_pool_same, max_pool_same
    torch.manual_seed(D + k)
    x = torch.randn((5, D, D, D, 96), device="cuda").to(torch.bfloat16)
    got = me.pool3d(x, 16, 64, k, is_max)
    v = x[..., 16:80].float().permute(0, 4, 1, 2, 3)
    ref = (max_pool_same(v, 2, 2) if is_max else avg_pool_same(v, k)).permute(0, 2, 3, 4, 1)
    assert tuple(got.shape) == tuple(ref.shape)
    assert torch.equal(got, ref.to(torch.bfloat16)) if is_max else (got.float() - ref).abs().max().item() < 2e-2
    if not is_max and D == 8:
        _lib.set_option("pool_variant", 1)
        try:
            plain = me.pool3d(x, 16, 64, k, is_max)
        finally:
            _lib.set_option("pool_variant", 0)
        # same fp32 sums in another order: at most one bf16 rounding step apart
        assert ((got.float() - plain.float()).abs() <= 2.0 ** -7 * plain.float().abs() + 1e-6).all()
        odd = me.pool3d(x, 8, 24, k, is_max)               # 24 channels: not a multiple of 32 -> per-voxel kernel
        ref24 = avg_pool_same(x[..., 8:32].float().permute(0, 4, 1, 2, 3), k).permute(0, 2, 3, 4, 1)
        assert (odd.float() - ref24).abs().max().item() < 2e-2


@pytest.mark.parametrize("D,k,nf,cin", [(8, 3, 64, 96), (8, 3, 48, 40), (4, 2, 128, 256), (2, 2, 32, 64), (4, 1, 64, 128)])
def test_fused_pool_branch_against_the_unfused_layers(D, k, nf, cin):
    """mups_conv1_split_bn_relu + mups_avgpool3d_bn_relu (the pool branch's 1^3 convolution computed with `one` from one read
    of the input, the average pool afterwards) against the reference order avg_pool -> conv -> bias + BN -> ReLU in fp32 on
    the same bf16-rounded operands (models/experts_n_est.py:296-307)."""
    from nesti_net_b200 import moe_engine as me
    from nesti_net_b200.experts_net import avg_pool_same
    torch.manual_seed(D * 7 + k + nf)
    dev = torch.device("cuda", 0)
    B = 5
    x = torch.randn((B, D, D, D, cin + 8), device=dev).to(torch.bfloat16)

    def layer():
        l = me.PackedConv.__new__(me.PackedConv)
        l.k, l.relu, l.cout, l.cout_pad, l.cin_pad = 1, True, nf, nf, cin
        l.w = (torch.randn((1, nf, cin), device=dev) / np.sqrt(cin)).to(torch.bfloat16).contiguous()
        l.scale, l.shift = torch.rand(nf, device=dev) + 0.5, torch.randn(nf, device=dev) * 0.1
        return l
    one, pool = layer(), layer()
    out = torch.full((B, D, D, D, 2 * nf + 16), 7.0, dtype=torch.bfloat16, device=dev)
    both = me._stack_one_and_pool(one, pool, k > 1)
    if k == 1:
        me.conv1_split(x, 8, cin, both, nf, 2 * nf, out, 0, out, nf + 16)
    else:
        pre = torch.empty((B, D, D, D, nf), dtype=torch.bfloat16, device=dev)
        me.conv1_split(x, 8, cin, both, nf, nf, out, 0, pre, 0)
        me.avgpool_bn_relu(pre, 0, nf, k, pool.scale, pool.shift, True, out, nf + 16)
    torch.cuda.synchronize()
    xs = x[..., 8:].float()
    ref_one = torch.relu(xs @ one.w[0].float().t() * one.scale + one.shift)
    pooled = avg_pool_same(xs.permute(0, 4, 1, 2, 3), k).permute(0, 2, 3, 4, 1)
    ref_pool = torch.relu(pooled @ pool.w[0].float().t() * pool.scale + pool.shift)
    assert torch.all(out[..., nf:nf + 16] == 7.0), "wrote outside its channel slices"
    assert (out[..., :nf].float() - ref_one).abs().max().item() < 2e-2 * max(1.0, ref_one.abs().max().item())
    # the pool half rounds the convolution's output to bf16 before the average: same error level as rounding the pooled input
    assert (out[..., nf + 16:].float() - ref_pool).abs().max().item() < 2e-2 * max(1.0, ref_pool.abs().max().item())


def test_tensor_core_consumer_against_fp32_network():
    """The Mixture-of-Experts forward on the tcgen05 kernels (moe_engine.TensorCoreExperts) against the fp32 PyTorch
    network (experts_net.ExpertsNormalEstimator, TF32 off) on GPU MuPS of a real cloud, batch norm statistics randomised
    so that the folding is exercised: same expert for (almost) every query, gate probabilities close, normals within a
    fraction of a degree -- bf16 products with fp32 accumulation; the deviation is printed."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg
    from nesti_net_b200.moe_engine import TensorCoreExperts
    pts = orc.synthetic_cloud(30000, cloud_id=9, noise=0.001)
    radius, P = [0.01, 0.03, 0.05, 0.07], 512
    w, mu, sg = grid_gmm(8, 0.0156)
    q = np.random.RandomState(8).choice(30000, 200, replace=False)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    mups = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q, index.absolute_radii(radius), P, seed=SEED)
    torch.manual_seed(1234)
    net = ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm1d)):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.6, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.05)
            if isinstance(m, (torch.nn.Conv3d, torch.nn.Linear)):
                m.bias.normal_(0, 0.02)
    net = net.cuda()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            prob_ref, n_ref = net(mups)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    tc = TensorCoreExperts(net)
    prob, n_est = tc.forward(mups)
    torch.cuda.synchronize()
    assert tuple(prob.shape) == tuple(prob_ref.shape) and tuple(n_est.shape) == tuple(n_ref.shape)
    assert torch.isfinite(prob).all() and torch.isfinite(n_est).all()
    same = prob.argmax(0) == prob_ref.argmax(0)
    rms = angular_rms_deg(n_est.reshape(-1, 3), n_ref.reshape(-1, 3))
    dprob = float((prob - prob_ref).abs().max())
    print("tensor-core consumer vs fp32 network: angular RMS %.3g deg over all experts, same expert %d/%d, max |dprob| %.3g"
          % (rms, int(same.sum()), len(q), dprob))
    assert rms < 1.0 and int(same.sum()) >= int(0.9 * len(q)) and dprob < 0.05
    sel, expert, probs = tc.predict(mups)
    assert tuple(sel.shape) == (len(q), 3) and tuple(probs.shape) == (len(q), 7) and torch.equal(expert, prob.argmax(0))


def test_tensor_core_consumer_against_reference_text(golden_dir):
    """The tcgen05 engine against the outputs of the REFERENCE'S OWN network text run on the emulated TF ops
    (tests/golden/moe_tf_emulated.npz, case g8: 4 scales, 7 experts, 8^3; variables restored by their TensorFlow names
    with ExpertsNormalEstimator.load_tf_variables): every expert's normal within 1 degree, gate probabilities within 0.02,
    the same expert chosen -- and the fp32 network on the GPU within 1e-3 degrees of the same fixture."""
    import sys
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg
    from nesti_net_b200.moe_engine import TensorCoreExperts
    if golden_dir not in sys.path:
        sys.path.insert(0, golden_dir)
    import moe_weights
    g = np.load(os.path.join(golden_dir, "moe_tf_emulated.npz"))
    mups = torch.from_numpy(g["g8_mups"]).cuda()
    net = ExpertsNormalEstimator(4, 512, 7).eval()
    table = {n: (d, l) for n, d, l in net.tf_variables()}

    def get(name):
        d, layout = table[name]
        shp = tuple(d.shape)
        shp = shp[2:] + (shp[1], shp[0]) if layout == "conv" else shp[::-1] if layout == "fc" else shp
        return moe_weights.variable_value(name, shp)
    net.load_tf_variables(get)
    net = net.cuda()
    ref_prob, ref_n = torch.from_numpy(g["g8_experts_prob"]).cuda(), torch.from_numpy(g["g8_n_est"]).cuda()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            prob32, n32 = net(mups)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    rms32 = angular_rms_deg(n32.reshape(-1, 3), ref_n.reshape(-1, 3))
    prob, n_est = TensorCoreExperts(net).forward(mups)
    torch.cuda.synchronize()
    rms = angular_rms_deg(n_est.reshape(-1, 3), ref_n.reshape(-1, 3))
    dprob = float((prob - ref_prob).abs().max())
    print("vs the reference's network text: fp32 network %.3g deg, tensor-core engine %.3g deg, max |dprob| %.3g"
          % (rms32, rms, dprob))
    assert rms32 < 1e-3 and float((prob32 - ref_prob).abs().max()) < 1e-4
    assert rms < 1.0 and dprob < 0.02 and torch.equal(prob.argmax(0), ref_prob.argmax(0))


def test_tensor_core_consumer_three_scales():
    """BASELINE configs[0] has three scales: seven experts = two per scale + one three-scale expert whose first inception
    module is 128 / 3 = 42 filters wide (Python-2 division, models/experts_n_est.py:254) -- channel counts that are not
    multiples of 16 go through the padded layouts of the engine (42 -> 48, 21 -> 32)."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg
    from nesti_net_b200.moe_engine import TensorCoreExperts
    pts = orc.synthetic_cloud(20000, cloud_id=4, noise=0.001)
    radius, P = [0.01, 0.03, 0.07], 512
    w, mu, sg = grid_gmm(8, 0.0156)
    q = np.random.RandomState(3).choice(20000, 96, replace=False)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    mups = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q, index.absolute_radii(radius), P, seed=SEED)
    torch.manual_seed(99)
    net = ExpertsNormalEstimator(n_rads=3, n_gaussians=512, n_experts=7).eval()
    assert net.expert_dict[6] == [0, 1, 2] and net.expert_conv[6].mods[0].one.conv.out_channels == 42
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm1d)):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.6, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.05)
    net = net.cuda()
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            prob_ref, n_ref = net(mups)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    prob, n_est = TensorCoreExperts(net).forward(mups)
    torch.cuda.synchronize()
    same = prob.argmax(0) == prob_ref.argmax(0)
    rms = angular_rms_deg(n_est.reshape(-1, 3), n_ref.reshape(-1, 3))
    assert torch.isfinite(n_est).all() and rms < 1.0 and int(same.sum()) >= int(0.9 * len(q)) and float((prob - prob_ref).abs().max()) < 0.05


def _triplet(t, off, w):
    """(hi + lo as fp32, hi, lo, second hi) of the triplet at channels [off, off + 3 w) of a bf16 tensor."""
    hi, lo, hi2 = t[..., off:off + w], t[..., off + w:off + 2 * w], t[..., off + 2 * w:off + 3 * w]
    return hi.float() + lo.float(), hi, lo, hi2


def test_bf16x3_split_and_pool_kernels():
    """csrc/moe_split.cu: mups_split_bf16x3 (fp32 -> triplet [hi | lo | hi], hi = bf16(v), lo = bf16(v - hi): bit-exact, zero
    padding of the tail, untouched neighbours) and mups_pool3d_bf16x3 (TF 'SAME' average / 2-2 max pool on hi + lo)."""
    from nesti_net_b200 import moe_engine as me
    from nesti_net_b200.experts_net import avg_pool_same, max_pool_same
    torch.manual_seed(5)
    rows = 3 * 8 * 8 * 8
    for (n_src, src_off, w_src, w_dst) in [(80, 20, 20, 32), (64, 0, 64, 64), (48, 8, 40, 40), (35, 3, 17, 24)]:
        src = torch.randn((rows, n_src), device="cuda") * torch.logspace(-6, 3, n_src, device="cuda")
        dst = torch.full((3, 8, 8, 8, 3 * w_dst + 16), 7.0, device="cuda", dtype=torch.bfloat16)
        me.split_x3(src, src_off, w_src, dst, 8, w_dst)
        torch.cuda.synchronize()
        v = torch.zeros((rows, w_dst), device="cuda")
        v[:, :w_src] = src[:, src_off:src_off + w_src]
        hi = v.to(torch.bfloat16)
        lo = (v - hi.float()).to(torch.bfloat16)
        d = dst.view(rows, -1)
        _, g_hi, g_lo, g_hi2 = _triplet(d, 8, w_dst)
        assert torch.equal(g_hi, hi) and torch.equal(g_lo, lo) and torch.equal(g_hi2, hi)
        assert torch.all(d[:, :8] == 7.0) and torch.all(d[:, 8 + 3 * w_dst:] == 7.0), "wrote outside its triplet"
        # 16 significant bits: the pair reproduces v to 2^-17 relative
        assert ((hi.float() + lo.float() - v).abs() <= 2.0 ** -16 * v.abs()).all()
    for (D, k, is_max) in [(8, 3, False), (8, 5, False), (8, 2, False), (4, 3, False), (4, 4, False), (2, 2, False), (8, 2, True), (4, 2, True), (2, 2, True)]:
        B, w, ct = 3, 24, 3 * 24 + 3 * 40 + 8
        x = torch.full((B, D, D, D, ct), 3.0, device="cuda", dtype=torch.bfloat16)
        val = torch.randn((B * D ** 3, w), device="cuda")
        me.split_x3(val, 0, w, x, 3 * 40, w)                      # the triplet under test sits behind another segment
        Do = D // 2 if is_max else D
        y = torch.full((B, Do, Do, Do, 3 * w + 16), 7.0, device="cuda", dtype=torch.bfloat16)
        me.pool3d_x3(x, 3 * 40, w, k, is_max, y, 16)
        torch.cuda.synchronize()
        v5 = _triplet(x, 3 * 40, w)[0].permute(0, 4, 1, 2, 3)
        ref = (max_pool_same(v5, 2, 2) if is_max else avg_pool_same(v5, k)).permute(0, 2, 3, 4, 1)
        got, g_hi, g_lo, g_hi2 = _triplet(y, 16, w)
        assert torch.equal(g_hi, g_hi2) and torch.all(y[..., :16] == 7.0)
        if is_max:
            assert torch.equal(got, ref)
        else:
            assert ((got - ref).abs() <= 3e-5 * ref.abs() + 1e-6).all()
    # the pool branch of an inception module: raw fp32 convolution output -> average pool -> scale / shift / ReLU -> triplet
    for (D, k, c, relu) in [(8, 3, 64, True), (8, 5, 32, True), (8, 2, 96, False), (8, 4, 128, True), (8, 3, 40, True), (4, 2, 48, True),
                            (4, 3, 64, False), (2, 2, 64, True)]:
        B = 3
        src = torch.randn((B * D ** 3, c), device="cuda")
        scale, shift = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
        y = torch.full((B, D, D, D, 3 * c + 24), 7.0, device="cuda", dtype=torch.bfloat16)
        me.avgpool_f32_x3(src, B, D, c, k, scale, shift, relu, y, 8)
        torch.cuda.synchronize()
        ref = avg_pool_same(src.view(B, D, D, D, c).permute(0, 4, 1, 2, 3), k).permute(0, 2, 3, 4, 1) * scale + shift
        ref = torch.relu(ref) if relu else ref
        got, g_hi, g_lo, g_hi2 = _triplet(y, 8, c)
        assert torch.equal(g_hi, g_hi2) and torch.all(y[..., :8] == 7.0) and torch.all(y[..., 8 + 3 * c:] == 7.0)
        assert ((got - ref).abs() <= 3e-5 * ref.abs() + 2e-6).all(), (D, k, c)


def test_tensor_core_consumer_bf16x3_against_fp32_network():
    """moe_engine.TensorCoreExperts(precision="bf16x3"): every operand carried as two bf16 numbers, a_hi w_hi + a_lo w_hi +
    a_hi w_lo accumulated in fp32 by the SAME tcgen05 kernels over a three times longer channel axis.  Against the fp32
    PyTorch network (TF32 off) on GPU MuPS of a real cloud: same expert for every query, normals within 0.01 degrees (the
    plain bf16 mode: 0.2), gate probabilities within 1e-4.  The four-scale and the three-scale configuration (padded
    42-filter expert) are both run."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg
    from nesti_net_b200.moe_engine import TensorCoreExperts
    for radius, nq, seed in (([0.01, 0.03, 0.05, 0.07], 96, 1234), ([0.01, 0.03, 0.07], 40, 99)):
        pts = orc.synthetic_cloud(30000, cloud_id=9, noise=0.001)
        w, mu, sg = grid_gmm(8, 0.0156)
        q = np.random.RandomState(8).choice(30000, nq, replace=False)
        index = mb.PointIndex(pts, cell_frac=max(radius))
        mups = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q, index.absolute_radii(radius), 512, seed=SEED)
        torch.manual_seed(seed)
        net = ExpertsNormalEstimator(n_rads=len(radius), n_gaussians=512, n_experts=7).eval()
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm1d)):
                    m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.6, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.05)
                if isinstance(m, (torch.nn.Conv3d, torch.nn.Linear)):
                    m.bias.normal_(0, 0.02)
        net = net.cuda()
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        try:
            with torch.no_grad():
                prob_ref, n_ref = net(mups)
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        tc = TensorCoreExperts(net, precision="bf16x3")
        prob, n_est = tc.forward(mups)
        torch.cuda.synchronize()
        assert tuple(prob.shape) == tuple(prob_ref.shape) and tuple(n_est.shape) == tuple(n_ref.shape)
        assert torch.isfinite(prob).all() and torch.isfinite(n_est).all()
        same = prob.argmax(0) == prob_ref.argmax(0)
        rms = angular_rms_deg(n_est.reshape(-1, 3), n_ref.reshape(-1, 3))
        dprob = float((prob - prob_ref).abs().max())
        rel = float((n_est - n_ref).abs().max() / n_ref.abs().max())
        print("bf16x3 consumer vs fp32 network (%d scales): angular RMS %.3g deg over all experts, max rel %.3g, same expert %d/%d, "
              "max |dprob| %.3g" % (len(radius), rms, rel, int(same.sum()), len(q), dprob))
        assert rms < 0.01 and bool(same.all()) and dprob < 1e-4


def test_downstream_moe_normals():
    """Fourth gate of BASELINE.json: the same randomly initialised Mixture-of-Experts (PyTorch restatement of
    models/experts_n_est.py, fp32) evaluated on oracle MuPS and on GPU MuPS gives normals within 1e-4 angular
    RMS (degrees, the unit of utils/evaluate.py).  The network runs on the host in strict fp32: it is the checker
    here (profiles/bench_moe.py measures it on the GPU)."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg
    pts = orc.synthetic_cloud(30000, cloud_id=9, noise=0.001)
    radius = [0.01, 0.03, 0.05, 0.07]
    P = 512
    w, mu, sg = grid_gmm(8, 0.0156)
    q = np.random.RandomState(8).choice(30000, 24, replace=False)
    index = mb.PointIndex(pts, cell_frac=max(radius))
    gpu_mups = mb.mups_features(index, mb.gmm_handle(w, mu, sg), q, index.absolute_radii(radius), P, seed=SEED).cpu()
    o_patches, o_neff, _ = orc.gather_patches(pts, q, radius, P, seed=SEED)
    ora_mups = torch.from_numpy(c_oracle.mups(o_patches, o_neff, w, mu, sg, 4))
    torch.manual_seed(1234)
    net = ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7).eval()
    with torch.no_grad():
        prob_g, n_g = net(gpu_mups)
        prob_o, n_o = net(ora_mups)
    rms_all = angular_rms_deg(n_g.reshape(-1, 3), n_o.reshape(-1, 3))                # every expert's normal
    sel_g, exp_g, _ = net.predict(gpu_mups)
    sel_o, exp_o, _ = net.predict(ora_mups)
    same = exp_g == exp_o
    rms_sel = angular_rms_deg(sel_g[same], sel_o[same])
    print("MoE normals: angular RMS %.3g deg over all experts, %.3g deg for the selected expert, "
          "expert agreement %d/%d, max |dprob| %.3g" % (rms_all, rms_sel, int(same.sum()), len(q), float((prob_g - prob_o).abs().max())))
    assert rms_all <= 1e-4 and rms_sel <= 1e-4
    assert int(same.sum()) >= len(q) - 1          # a near-tie of two gate outputs may flip at most one query
    assert float((prob_g - prob_o).abs().max()) < 1e-5
