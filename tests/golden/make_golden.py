#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/.

Run in the BUILD container only (it imports the unmodified reference from
/root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

half1_reference.npz  -- outputs of the reference's own
    utils/pcpnet_dataset.py::PointcloudPatchDataset.__getitem__ (unmodified,
    imported from /root/reference/utils) on a small synthetic cloud:
    raw cKDTree neighbour lists, effective point counts, absolute radii and
    the patch tensors.  Rows of patches whose neighbourhood was not subsampled
    are stored re-ordered to ascending neighbour index (the kd-tree returns
    traversal order); subsampled patches are stored as the reference produced
    them (only the subset property can be checked for those, the reference's
    draw comes from a stateful stream shared by all patches).
half2_reference_numpy.npz -- outputs of the reference's own numpy code run
    here: ``utils/utils.py::get_3d_grid_gmm`` (:70-95, the grid GMM fed to the
    graph) and ``utils/utils.py::get_3DmFV`` (:260-330).  The latter differs
    from the TF path in ONE stage (it takes Q = p, no posterior normalisation,
    and has no mask), so it pins every other stage of the half-2 restatement
    (``get_3dmfv_n_est(..., _posterior="pdf")``) against the reference itself; and
    ``utils/utils.py::fisher_vector_per_point`` (:214-245), whose per-point terms use the posterior
    (sklearn ``predict_proba``) and pin that stage too.  Only the n_eff mask exists in the TF code alone.
half2_tf_emulated.npz -- outputs of the reference's own TensorFlow source text
    (utils/tf_util.py::get_3dmfv_n_est / get_3dmfv, the MuPS loop of
    models/experts_n_est.py::get_model), extracted with `ast` and executed with
    `tf` bound to tests/golden/tf1_emulation.py (numpy emulation of the
    primitive TF ops): the mask stage and the anisotropic prefactor run as the
    reference wrote them.
moe_tf_emulated.npz -- outputs of the reference's own NETWORK text (get_model's
    statements after the MuPS loop, scale_manager_net, conv_net_8g / _3g,
    normal_est_net, inception_module and the tf_util layers) run on
    tests/golden/tf1_emulation_nn.py with variables from tests/golden/moe_weights.py:
    pins the architecture of the consumer (experts_net.py, moe_engine.py).
half2_oracle.npz -- inputs/outputs of the recorded fp32 transliteration of
    utils/tf_util.py:655-753 / :578-652 (TensorFlow 1.12 is not installable:
    PARITY UNPINNED), written only after the independent float64 restatement
    agrees within 1e-6.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import mups_oracle as orc  # noqa: E402

REF_UTILS = "/root/reference/utils"


def make_half1():
    sys.path.insert(0, REF_UTILS)
    import pcpnet_dataset as ref  # the unmodified reference module

    cases = {}
    # case A: small patches so both branches (len<=P and len>P) occur
    # case B: the reference's P=512 with 4 scales
    specs = {
        "A": dict(n=4000, cloud_id=0, P=48, radii=[0.03, 0.08, 0.15], nq=96, noise=0.0),
        "B": dict(n=30000, cloud_id=1, P=512, radii=[0.01, 0.03, 0.05, 0.07], nq=32, noise=0.002),
    }
    for name, sp in specs.items():
        pts = orc.synthetic_cloud(sp["n"], cloud_id=sp["cloud_id"], noise=sp["noise"])
        # duplicate a few points and put a query on the bbox corner (edge cases)
        pts[10] = pts[11]
        pts[12] = pts[11]
        with tempfile.TemporaryDirectory() as d:
            # np.loadtxt(...).astype('float32') must give back the same float32 bits: %.9g round-trips
            np.savetxt(os.path.join(d, "cloud.xyz"), pts, fmt="%.9g")
            with open(os.path.join(d, "list.txt"), "w") as f:
                f.write("cloud\n")
            ds = ref.PointcloudPatchDataset(
                root=d, shape_list_filename="list.txt", patch_radius=sp["radii"],
                points_per_patch=sp["P"], patch_features=[], seed=3627473, identical_epochs=False,
                use_pca=False, center="point", point_tuple=1, cache_capacity=100,
                point_count_std=0, sparse_patches=False)
            shape = ds.shape_cache.get(0)
            assert np.array_equal(shape.pts, pts), "xyz text round trip changed the float32 bits"
            rng = np.random.RandomState(7)
            q = np.unique(np.concatenate([
                rng.choice(sp["n"], sp["nq"] - 6, replace=False),
                [10, 11, 12, int(pts[:, 0].argmax()), int(pts[:, 1].argmin()), int(pts[:, 2].argmax())]]))
            S, P = len(sp["radii"]), sp["P"]
            rads = ds.patch_radius_absolute[0]
            patches = np.zeros((len(q), S * P, 3), np.float32)
            n_eff = np.zeros((len(q), S), np.int32)
            nbr_flat, nbr_off, subsampled = [], [0], np.zeros((len(q), S), bool)
            for b, c in enumerate(q):
                item = ds[int(c)]
                pp = item[0].numpy().copy()
                ne = np.asarray(item[-1]).astype(np.int32)
                n_eff[b] = ne
                for s, rad in enumerate(rads):
                    raw = np.array(shape.kdtree.query_ball_point(shape.pts[int(c), :], rad))
                    nbr_flat.append(np.sort(raw))
                    nbr_off.append(nbr_off[-1] + len(raw))
                    if len(raw) <= P:
                        order = np.argsort(raw, kind="stable")
                        pp[s * P: s * P + len(raw)] = pp[s * P: s * P + len(raw)][order]
                    else:
                        subsampled[b, s] = True
                patches[b] = pp
            cases[name] = dict(
                pts=pts, query_idx=q.astype(np.int64), patch_radius=np.asarray(sp["radii"], np.float64),
                P=np.int64(P), radii_abs=np.asarray(rads, np.float64),
                bbdiag=np.float64(rads[0] / sp["radii"][0]),
                nbr_flat=np.concatenate(nbr_flat).astype(np.int32), nbr_off=np.asarray(nbr_off, np.int64),
                n_eff=n_eff, patches=patches, subsampled=subsampled)
            print("half1 case", name, "queries", len(q), "subsampled", int(subsampled.sum()),
                  "of", subsampled.size, "max len", int(np.diff(nbr_off).max()))
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat["%s_%s" % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, "half1_reference.npz"), **flat)


def make_dataset_interface():
    """The dataset interface around half 1, run on the unmodified reference: two shapes, sparse patch centres (.pidx),
    targets 'normal' / 'max_curvature' / 'min_curvature' (pcpnet_dataset.py:292-295, 345-352, 405-417), global index ->
    (shape, patch) mapping (:427-436)."""
    sys.path.insert(0, REF_UTILS)
    import pcpnet_dataset as ref
    rng = np.random.RandomState(31)
    radii, P = [0.04, 0.1], 32
    out = {"patch_radius": np.asarray(radii, np.float64), "P": np.int64(P)}
    with tempfile.TemporaryDirectory() as d:
        names = ["shape_a", "shape_b"]
        for k, name in enumerate(names):
            n = 2500 + 700 * k
            pts = orc.synthetic_cloud(n, cloud_id=40 + k, noise=0.001)
            nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
            curv = rng.uniform(-2, 2, size=(n, 2))
            pidx = np.sort(rng.choice(n, 40 + 10 * k, replace=False))
            np.savetxt(os.path.join(d, name + ".xyz"), pts, fmt="%.9g")
            np.savetxt(os.path.join(d, name + ".normals"), nrm.astype(np.float32), fmt="%.9g")
            np.savetxt(os.path.join(d, name + ".curv"), curv.astype(np.float32), fmt="%.9g")
            np.savetxt(os.path.join(d, name + ".pidx"), pidx, fmt="%d")
            out.update({name + "_pts": pts, name + "_normals": nrm.astype(np.float32), name + "_curv": curv.astype(np.float32),
                        name + "_pidx": pidx.astype(np.int64)})
        with open(os.path.join(d, "list.txt"), "w") as f:
            f.write("\n".join(names) + "\n")
        ds = ref.PointcloudPatchDataset(
            root=d, shape_list_filename="list.txt", patch_radius=radii, points_per_patch=P,
            patch_features=["normal", "max_curvature", "min_curvature"], seed=3627473, identical_epochs=False,
            use_pca=False, center="point", point_tuple=1, cache_capacity=100, point_count_std=0, sparse_patches=True)
        out["shape_patch_count"] = np.asarray(ds.shape_patch_count, np.int64)
        out["patch_radius_absolute"] = np.asarray(ds.patch_radius_absolute, np.float64)
        out["length"] = np.int64(len(ds))
        idx = np.asarray([0, 1, 17, 39, 40, 41, 63, 89], np.int64)        # both shapes, both ends
        items = [ds[int(i)] for i in idx]
        out["item_index"] = idx
        out["item_shape_index"] = np.asarray([ds.shape_index(int(i)) for i in idx], np.int64)
        out["item_normal"] = np.stack([it[1].numpy() for it in items])
        out["item_max_curv"] = np.stack([it[2].numpy() for it in items])
        out["item_min_curv"] = np.stack([it[3].numpy() for it in items])
        out["item_trans"] = np.stack([it[4].numpy() for it in items])
        out["item_n_eff"] = np.stack([np.asarray(it[5], np.float64) for it in items])
        print("dataset interface: len", len(ds), "counts", ds.shape_patch_count, "dtypes",
              items[0][1].dtype, items[0][2].dtype, np.asarray(items[0][5]).dtype)
    np.savez_compressed(os.path.join(HERE, "dataset_reference.npz"), **out)


def make_sampler_reference():
    """Index sequences of the reference's three samplers (pcpnet_dataset.py:41-138) over two epochs."""
    sys.path.insert(0, REF_UTILS)
    import pcpnet_dataset as ref

    class DS(object):
        shape_names = ["a", "b", "c", "d"]
        shape_patch_count = [10, 5, 20, 7]
    ds = DS()
    out = {"shape_patch_count": np.asarray(ds.shape_patch_count, np.int64)}
    out["sequential"] = np.asarray(list(ref.SequentialPointcloudPatchSampler(ds)), np.int64)
    for ident in (False, True):
        r = ref.RandomPointcloudPatchSampler(ds, patches_per_shape=8, seed=3627473, identical_epochs=ident)
        out["random_ident%d" % ident] = np.asarray([list(r), list(r)], np.int64)
        for seq_shapes in (False, True):
            q = ref.SequentialShapeRandomPointcloudPatchSampler(ds, 8, seed=3627473, sequential_shapes=seq_shapes,
                                                                identical_epochs=ident)
            out["shape_random_ident%d_seq%d" % (ident, seq_shapes)] = np.asarray([list(q), list(q)], np.int64)
            out["shape_random_ident%d_seq%d_local" % (ident, seq_shapes)] = np.concatenate(
                [np.asarray(v, np.int64) for v in q.shape_patch_inds])
    np.savez_compressed(os.path.join(HERE, "sampler_reference.npz"), **out)
    print("sampler reference:", {k: v.shape for k, v in out.items()})


def make_evaluation_reference():
    """utils/evaluate.py run as the script it is (runpy) on synthetic predictions: per-shape unoriented RMS angle,
    PGP5, PGP10 and the shape averages, parsed from the summary it writes.  `visualization` and `utils` (matplotlib /
    h5py importers, used only under EXPORT) are satisfied with empty shims; the evaluation code itself is untouched."""
    import ast
    import runpy
    import types
    rng = np.random.RandomState(17)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        data, res = os.path.join(d, "data") + os.sep, os.path.join(d, "run", "results") + os.sep
        os.makedirs(data)
        os.makedirs(res)
        names = ["s0", "s1", "s2"]
        for k, name in enumerate(names):
            n = 400 + 50 * k
            pts = rng.normal(size=(n, 3)).astype(np.float32)
            gt = rng.normal(size=(n, 3)); gt /= np.linalg.norm(gt, axis=1, keepdims=True)
            pidx = np.sort(rng.choice(n, 120, replace=False))
            # predictions: the ground truth perturbed by 1..25 degrees, random sign (unoriented), not normalised
            pert = gt + rng.normal(size=(n, 3)) * rng.uniform(0.02, 0.45, size=(n, 1))
            pred_full = pert * rng.choice([-1.0, 1.0], size=(n, 1)) * rng.uniform(0.5, 2.0, size=(n, 1))
            pred = pred_full[pidx] if k != 1 else pred_full          # shape s1: dense predictions, subset by .pidx
            np.savetxt(data + name + ".xyz", pts, fmt="%.9g")
            np.savetxt(data + name + ".normals", gt.astype(np.float32), fmt="%.9g")
            np.savetxt(data + name + ".pidx", pidx, fmt="%d")
            np.savetxt(res + name + ".normals", pred.astype(np.float32), fmt="%.9g")
            out.update({name + "_gt": gt.astype(np.float32), name + "_pred": pred.astype(np.float32), name + "_pidx": pidx.astype(np.int64)})
        with open(data + "evalset.txt", "w") as f:
            f.write("\n".join(names) + "\n")
        for mod in ("visualization", "utils"):
            sys.modules[mod] = types.ModuleType(mod)
        argv = sys.argv
        sys.argv = ["evaluate.py", "--normal_results_path", res, "--data_path", data, "--dataset_list", "evalset"]
        try:
            runpy.run_path(os.path.join(REF_UTILS, "evaluate.py"), run_name="__main__")
        finally:
            sys.argv = argv
            for mod in ("visualization", "utils"):
                sys.modules.pop(mod, None)
        with open(os.path.join(res, "summary", "evalset_evaluation_results.txt")) as f:
            lines = dict(line.strip().split(": ", 1) for line in f if ": " in line)
    def parse(v):
        return np.asarray(eval(v, {"__builtins__": {}}, {"np": np, "float32": np.float32, "float64": np.float64}), np.float64)
    out["rms"] = parse(lines["RMS per shape"])
    out["pgp10"] = parse(lines["PGP10 per shape"])
    out["pgp5"] = parse(lines["PGP5 per shape"])
    out["avg_rms"] = parse(lines["RMS not oriented (shape average)"])
    out["avg_rms_o"] = parse(lines["RMS oriented (shape average)"])
    out["avg_pgp10"] = parse(lines["PGP10 average"])
    out["avg_pgp5"] = parse(lines["PGP5 average"])
    out["names"] = np.asarray(names)
    print("evaluation reference:", {k: out[k] for k in ("rms", "pgp5", "pgp10", "avg_rms_o")})
    np.savez_compressed(os.path.join(HERE, "evaluation_reference.npz"), **out)


def make_half2():
    rng = np.random.RandomState(11)
    out = {}
    # (res, P, variance, n_eff list)
    specs = {
        "g3": (3, 16, 0.11, [1, 2, 5, 14, 15, 16]),
        "g8": (8, 64, 0.0156, [1, 2, 30, 62, 63, 64]),
        "g8p512": (8, 512, 0.0156, [17, 510, 511, 512]),
    }
    for name, (res, P, var, neffs) in specs.items():
        w, mu, sigma = orc.gmm_feed(*orc.get_3d_grid_gmm([res] * 3, var))
        B = len(neffs)
        pts = np.zeros((B, P, 3), np.float32)
        for b, ne in enumerate(neffs):
            x = rng.normal(size=(ne, 3)) * 0.4
            x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))   # inside the unit ball
            x[0] = 0.0                                                     # the centre is its own neighbour
            pts[b, :ne] = x
        ne = np.asarray(neffs, np.int32)
        fv = orc.get_3dmfv_n_est(pts, w, mu, sigma, flatten=False, n_original_points=ne)
        f64 = orc.get_3dmfv_n_est_f64(pts, w, mu, sigma, ne, masked=True)
        err = np.abs(fv - f64).max()
        assert err < 1e-6, (name, err)
        fv_plain = orc.get_3dmfv(pts, w, mu, sigma, flatten=False)
        f64_plain = orc.get_3dmfv_n_est_f64(pts, w, mu, sigma, None, masked=False)
        err2 = np.abs(fv_plain - f64_plain).max()
        assert err2 < 1e-6, (name, err2)
        print("half2 case", name, "fp32 vs f64 max abs err", err, err2)
        out[name + "_points"] = pts
        out[name + "_n_eff"] = ne
        out[name + "_w"] = w
        out[name + "_mu"] = mu
        out[name + "_sigma"] = sigma
        out[name + "_fv_n_est"] = fv
        out[name + "_fv_plain"] = fv_plain
    # general (non-grid) GMM: non-uniform w, anisotropic per-Gaussian sigma
    G, P = 27, 32
    w = rng.uniform(0.5, 1.5, G); w = (w / w.sum()).astype(np.float32)
    mu = rng.uniform(-0.8, 0.8, (G, 3)).astype(np.float32)
    sigma = rng.uniform(0.2, 0.5, (G, 3)).astype(np.float32)
    neffs = np.asarray([1, 7, 30, 31, 32], np.int32)
    pts = np.zeros((len(neffs), P, 3), np.float32)
    for b, ne in enumerate(neffs):
        x = rng.uniform(-0.7, 0.7, (ne, 3)); x[0] = 0
        pts[b, :ne] = x
    fv = orc.get_3dmfv_n_est(pts, w, mu, sigma, flatten=False, n_original_points=neffs)
    f64 = orc.get_3dmfv_n_est_f64(pts, w, mu, sigma, neffs, masked=True)
    assert np.abs(fv - f64).max() < 1e-6
    fv_plain = orc.get_3dmfv(pts, w, mu, sigma, flatten=False)
    f64_plain = orc.get_3dmfv_n_est_f64(pts, w, mu, sigma, None, masked=False)
    assert np.abs(fv_plain - f64_plain).max() < 1e-6
    print("half2 case general, fp32 vs f64", np.abs(fv - f64).max(), np.abs(fv_plain - f64_plain).max())
    out.update(gen_points=pts, gen_n_eff=neffs, gen_w=w, gen_mu=mu, gen_sigma=sigma,
               gen_fv_n_est=fv, gen_fv_plain=fv_plain)
    np.savez_compressed(os.path.join(HERE, "half2_oracle.npz"), **out)


def make_half2_reference_numpy():
    """Run the reference's numpy get_3d_grid_gmm / get_3DmFV unmodified.  utils/utils.py imports the
    reference's provider.py (h5py, absent here) and a private sklearn symbol that moved
    (utils.py:93): both are satisfied with import shims, the reference code itself is untouched."""
    import types
    import sklearn.mixture._gaussian_mixture as skgm
    sys.modules.setdefault("provider", types.ModuleType("provider"))
    shim = types.ModuleType("sklearn.mixture.gaussian_mixture")
    shim._compute_precision_cholesky = skgm._compute_precision_cholesky
    sys.modules.setdefault("sklearn.mixture.gaussian_mixture", shim)
    if REF_UTILS not in sys.path:
        sys.path.insert(0, REF_UTILS)
    import utils as ref_utils  # the unmodified reference module

    rng = np.random.RandomState(23)
    out = {}
    for n, var in ((3, 0.11), (8, 0.0156), (16, 0.00390625)):
        gmm = ref_utils.get_3d_grid_gmm(subdivisions=[n, n, n], variance=var)
        out["grid%d_variance" % n] = np.float64(var)
        out["grid%d_weights" % n] = np.asarray(gmm.weights_)
        out["grid%d_means" % n] = np.asarray(gmm.means_)
        out["grid%d_covariances" % n] = np.asarray(gmm.covariances_)
    # (name, res, variance, B, P): points inside the unit ball like real patches; fp32 inputs
    for name, res, var, B, P in (("g3", 3, 0.11, 5, 16), ("g8", 8, 0.0156, 4, 64), ("g8p512", 8, 0.0156, 2, 512)):
        gmm = ref_utils.get_3d_grid_gmm(subdivisions=[res] * 3, variance=var)
        # what the scripts feed the graph (train_n_est_w_experts.py:284-286)
        w, mu, sigma = (np.asarray(gmm.weights_, np.float32), np.asarray(gmm.means_, np.float32),
                        np.sqrt(gmm.covariances_).astype(np.float32))
        x = rng.normal(size=(B, P, 3)) * 0.4
        x /= np.maximum(1.0, np.linalg.norm(x, axis=2, keepdims=True))
        x[:, 0] = 0.0
        pts = x.astype(np.float32)
        fv = ref_utils.get_3DmFV(pts, w, mu, sigma, normalize=True)           # [B, 20, G]
        fv_raw = ref_utils.get_3DmFV(pts, w, mu, sigma, normalize=False)
        mine = orc.get_3dmfv_n_est(pts, w, mu, sigma, flatten=False,
                                   n_original_points=np.full(B, P, np.int32), _posterior="pdf")
        print("half2 reference numpy case", name, fv.shape, "oracle(pdf hook) vs reference max abs err",
              float(np.abs(mine - fv).max()))
        out.update({name + "_points": pts, name + "_w": w, name + "_mu": mu, name + "_sigma": sigma,
                    name + "_fv": np.asarray(fv, np.float64), name + "_fv_raw": np.asarray(fv_raw, np.float64)})
    # The posterior stage: the reference's fisher_vector_per_point (utils/utils.py:214-245) evaluates the per-point
    # derivative terms with Q = gmm.predict_proba(xx) (sklearn, float64) -- the posterior w p / sum_g w p of
    # tf_util.py:700-701 -- including the 1/sqrt(w), 1/sqrt(2w) factors.  Reduced over the points with the statements of
    # the reference's own get_3DmFV tail (:302-326: max/min/sum, /n_points, signed sqrt, l2_normalize, channel order)
    # this is get_3dmfv_n_est with nothing masked.
    for name, res, var, B, P in (("post_g3", 3, 0.11, 5, 16), ("post_g8", 8, 0.0156, 3, 64)):
        gmm = ref_utils.get_3d_grid_gmm(subdivisions=[res] * 3, variance=var)
        w, mu, sigma = (np.asarray(gmm.weights_, np.float32), np.asarray(gmm.means_, np.float32),
                        np.sqrt(gmm.covariances_).astype(np.float32))
        x = rng.normal(size=(B, P, 3)) * 0.4
        x /= np.maximum(1.0, np.linalg.norm(x, axis=2, keepdims=True))
        x[:, 0] = 0.0
        pts = x.astype(np.float32)
        fv = []
        for b in range(B):
            d_pi_all, d_mu_all, d_sig_all = ref_utils.fisher_vector_per_point(pts[b].astype(np.float64), gmm)
            d_pi = np.concatenate([np.max(d_pi_all[..., None], axis=0), np.sum(d_pi_all[..., None], axis=0)], axis=1)
            d_mu = np.concatenate([np.max(d_mu_all, axis=0), np.min(d_mu_all, axis=0), np.sum(d_mu_all, axis=0)], axis=1)
            d_sigma = np.concatenate([np.max(d_sig_all, axis=0), np.min(d_sig_all, axis=0), np.sum(d_sig_all, axis=0)], axis=1)
            d_pi, d_mu, d_sigma = d_pi / P, d_mu / P, d_sigma / P
            alpha = 0.5
            d_pi = np.sign(d_pi) * np.power(np.abs(d_pi), alpha)
            d_mu = np.sign(d_mu) * np.power(np.abs(d_mu), alpha)
            d_sigma = np.sign(d_sigma) * np.power(np.abs(d_sigma), alpha)
            d_pi = ref_utils.l2_normalize(d_pi, dim=0)
            d_mu = ref_utils.l2_normalize(d_mu, dim=0)
            d_sigma = ref_utils.l2_normalize(d_sigma, dim=0)
            fv.append(np.concatenate([d_pi, d_mu, d_sigma], axis=1).T)          # [20, G]
        fv = np.asarray(fv)
        mine = orc.get_3dmfv_n_est(pts, w, mu, sigma, flatten=False, n_original_points=np.full(B, P, np.int32))
        print("half2 posterior case", name, fv.shape, "oracle vs reference (sklearn posterior) max abs err",
              float(np.abs(mine - fv).max()))
        out.update({name + "_points": pts, name + "_w": w, name + "_mu": mu, name + "_sigma": sigma, name + "_fv": fv})
    np.savez_compressed(os.path.join(HERE, "half2_reference_numpy.npz"), **out)


def _reference_function_source(path, name):
    """Source text of top-level function `name` in the reference file `path` (unmodified)."""
    import ast
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


def _reference_mups_loop_source(path):
    """The MuPS statements of models/experts_n_est.py::get_model (:59-76): the lines from `n_rads = len(radius)` to the
    end of the `for s in range(n_rads)` loop, as written in the reference.  (Cut out textually: the file as shipped
    does not parse -- an unbalanced parenthesis at :103 -- so `ast` cannot be used on it.)"""
    lines = open(path).read().split("\n")
    top = next(i for i, ln in enumerate(lines) if ln.startswith("def get_model("))
    first = next(i for i, ln in enumerate(lines) if i > top and ln.strip().startswith("n_rads = len(radius)"))
    last = next(i for i, ln in enumerate(lines) if i > first and ln.strip().startswith("experts_prob = scale_manager_net("))
    pad = len(lines[first]) - len(lines[first].lstrip())
    return "\n".join(ln.rstrip("\r")[pad:] if ln.strip() else "" for ln in lines[first:last])


def make_half2_tf_emulated():
    """half2_tf_emulated.npz: the REFERENCE'S OWN TensorFlow code for half 2 -- utils/tf_util.py::get_3dmfv_n_est
    (:655-753), ::get_3dmfv (:578-652) and the MuPS loop of models/experts_n_est.py::get_model (:59-76) -- executed
    statement for statement in this container with `tf` bound to tests/golden/tf1_emulation.py (a numpy emulation of
    the ~25 primitive TF ops those functions call; TensorFlow 1.12 itself cannot be installed).  This is the only
    fixture that pins the n_eff mask stage (:691-703) and the sigma_0^3 prefactor for anisotropic sigma against the
    reference's text rather than against a restatement of it."""
    sys.path.insert(0, HERE)
    import tf1_emulation as tf
    ns = {"tf": tf, "np": np}
    for fn in ("get_3dmfv_n_est", "get_3dmfv"):
        exec(compile(_reference_function_source("/root/reference/utils/tf_util.py", fn), "tf_util.py::" + fn, "exec"), ns)
    loop_src = _reference_mups_loop_source("/root/reference/models/experts_n_est.py")

    class _TfUtil(object):
        get_3dmfv_n_est = staticmethod(ns["get_3dmfv_n_est"])

    rng = np.random.RandomState(77)
    out = {}
    cases = {
        "g3": dict(gmm=orc.gmm_feed(*orc.get_3d_grid_gmm([3, 3, 3], 0.11)), P=24, B=10),
        "g8": dict(gmm=orc.gmm_feed(*orc.get_3d_grid_gmm([8, 8, 8], 0.0156)), P=40, B=9),
    }
    G = 11
    w = rng.uniform(0.5, 1.5, G)
    cases["gen"] = dict(gmm=((w / w.sum()).astype(np.float32), rng.uniform(-0.8, 0.8, (G, 3)).astype(np.float32),
                             rng.uniform(0.2, 0.6, (G, 3)).astype(np.float32)), P=17, B=8)
    for name, c in cases.items():
        w, mu, sg = c["gmm"]
        P, B = c["P"], c["B"]
        edge = [1, 2, 3, P // 2, P - 2, P - 1, P]
        ne = np.array([edge[b % len(edge)] for b in range(B)], np.int32)
        pts = np.zeros((B, P, 3), np.float32)
        for b in range(B):
            x = rng.normal(size=(ne[b], 3)) * rng.uniform(0.15, 0.6)
            x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
            x[0] = 0
            pts[b, :ne[b]] = x
        tp, tw, tmu, tsg = tf.Tensor(pts), tf.Tensor(w), tf.Tensor(mu), tf.Tensor(sg)
        flat = ns["get_3dmfv_n_est"](tp, tw, tmu, tsg, flatten=True, n_original_points=tf.Tensor(ne)).a
        cube = ns["get_3dmfv_n_est"](tp, tw, tmu, tsg, flatten=False, n_original_points=tf.Tensor(ne)).a
        plain = ns["get_3dmfv"](tp, tw, tmu, tsg, flatten=False).a
        assert flat.dtype == np.float32 and flat.shape == (B, 20 * len(w)) and cube.shape == (B, 20, len(w))
        assert np.array_equal(flat.reshape(B, 20, len(w)), cube)
        out.update({name + "_points": pts, name + "_n_eff": ne, name + "_w": w, name + "_mu": mu, name + "_sigma": sg,
                    name + "_fv_n_est": cube, name + "_fv_plain": plain})
        # the transliteration of oracle/ must agree with the reference's text run on the emulated ops
        mine = orc.get_3dmfv_n_est(pts, w, mu, sg, flatten=False, n_original_points=ne)
        err = np.abs(mine - cube)
        assert np.all(err <= 1e-6 + 1e-5 * np.abs(cube)), (name, float(err.max()))
        mine = orc.get_3dmfv(pts, w, mu, sg, flatten=False)
        err = np.abs(mine - plain)
        assert np.all(err <= 1e-6 + 1e-5 * np.abs(plain)), (name + " plain", float(err.max()))
        print("tf-emulated %s: get_3dmfv_n_est %s, get_3dmfv %s" % (name, cube.shape, plain.shape))
    # MuPS assembly (experts_n_est.py:59-76): two scales of the g3 case
    w, mu, sg = cases["g3"]["gmm"]
    P, B = cases["g3"]["P"], cases["g3"]["B"]
    pts2 = np.concatenate([out["g3_points"], out["g3_points"][::-1]], axis=1)
    ne2 = np.stack([out["g3_n_eff"], out["g3_n_eff"][::-1]], axis=1)
    env = {"tf": tf, "np": np, "tf_util": _TfUtil, "points": tf.Tensor(pts2), "w": tf.Tensor(w), "mu": tf.Tensor(mu),
           "sigma": tf.Tensor(sg), "radius": [0.05, 0.1], "original_n_points": tf.Tensor(ne2), "range": range, "len": len,
           "int": int}
    exec(compile(loop_src, "experts_n_est.py::get_model[MuPS]", "exec"), env)
    mups = env["MuPS"].a
    assert mups.shape == (B, 3, 3, 3, 40)
    out.update({"mups_points": pts2, "mups_n_eff": ne2, "mups_out": mups})
    assert np.allclose(mups, orc.mups_assemble(pts2, w, mu, sg, ne2, 2), rtol=1e-5, atol=1e-6)
    np.savez_compressed(os.path.join(HERE, "half2_tf_emulated.npz"), **out)
    print("tf-emulated MuPS %s" % (mups.shape,))


def _reference_function_text(path, name):
    """Source text of top-level function `name`, cut out TEXTUALLY (from its `def` line to the next top-level statement):
    models/experts_n_est.py as shipped does not parse (unbalanced parenthesis at :103), so `ast` cannot be used on it."""
    lines = open(path).read().split("\n")
    first = next(i for i, ln in enumerate(lines) if ln.startswith("def %s(" % name))
    last = next((i for i, ln in enumerate(lines) if i > first and ln[:1] not in ("", " ", "\t", "#", "\r")), len(lines))
    return "\n".join(ln.rstrip("\r") for ln in lines[first:last])


def _reference_network_statements(path):
    """The network statements of models/experts_n_est.py::get_model (:78-106): from `experts_prob = scale_manager_net(`
    to `n_est = tf.stack(experts)`, as written in the reference -- except that ONE closing parenthesis is removed from the
    `experts.append(normal_est_net(...` statement (:102-103), which has one too many as shipped (a syntax error)."""
    lines = open(path).read().split("\n")
    top = next(i for i, ln in enumerate(lines) if ln.startswith("def get_model("))
    first = next(i for i, ln in enumerate(lines) if i > top and ln.strip().startswith("experts_prob = scale_manager_net("))
    last = next(i for i, ln in enumerate(lines) if i > first and ln.strip().startswith("n_est = tf.stack(experts)"))
    pad = len(lines[first]) - len(lines[first].lstrip())
    body = [ln.rstrip("\r")[pad:] if ln.strip() else "" for ln in lines[first:last + 1]]
    bad = [i for i, ln in enumerate(body) if ln.rstrip().endswith("divider=len(expert_dict[i]))))")]
    assert len(bad) == 1, "the reference's syntax error is expected exactly once"
    body[bad[0]] = body[bad[0]].rstrip()[:-1]
    return "\n".join(body)


def make_moe_tf_emulated():
    """moe_tf_emulated.npz: the REFERENCE'S OWN network code -- the statements of get_model after the MuPS loop
    (models/experts_n_est.py:78-106: gate, default expert assignment, channel slices, experts), scale_manager_net,
    conv_net_8g, conv_net_3g, normal_est_net, inception_module (:155-310) and the layers they call in utils/tf_util.py
    (conv3d, fully_connected, max_pool3d, avg_pool3d, batch_norm_template, batch_norm_for_conv3d / _fc and the variable
    helpers) -- executed statement for statement with `tf` bound to tests/golden/tf1_emulation_nn.py, is_training = False,
    variables restored by name from tests/golden/moe_weights.py.  Python-2 semantics are kept where the text relies on them
    (`128 / divider` is an integer division).  Pins the architecture of nesti-net_b200/experts_net.py (and through it the
    tensor-core engine) against the reference's text instead of a reading of it."""
    import types
    sys.path.insert(0, HERE)
    import tf1_emulation_nn as tf
    import moe_weights  # noqa: F401
    util_ns = {"tf": tf, "np": np}
    for fn in ("_variable_on_cpu", "_variable_with_weight_decay", "conv3d", "fully_connected", "max_pool3d", "avg_pool3d",
               "batch_norm_template", "batch_norm_for_fc", "batch_norm_for_conv3d"):
        exec(compile(_reference_function_source("/root/reference/utils/tf_util.py", fn), "tf_util.py::" + fn, "exec"), util_ns)
    tf_util = types.SimpleNamespace(**{k: v for k, v in util_ns.items() if callable(v) and k not in ("tf", "np")})
    model = "/root/reference/models/experts_n_est.py"
    py2_len = lambda x: tf.Py2Int(len(x))                       # noqa: E731  (so that `128 / divider` floors, as under 2.7)
    net_ns = {"tf": tf, "np": np, "tf_util": tf_util, "len": py2_len}
    for fn in ("scale_manager_net", "conv_net_8g", "conv_net_3g", "normal_est_net", "inception_module"):
        exec(compile(_reference_function_text(model, fn), "experts_n_est.py::" + fn, "exec"), net_ns)
    statements = compile(_reference_network_statements(model), "experts_n_est.py::get_model[network]", "exec")

    rng = np.random.RandomState(2018)
    out = {}
    cases = {"g3": dict(res=3, var=0.11, n_rads=2, n_experts=3, B=4),
             "g3s3": dict(res=3, var=0.11, n_rads=3, n_experts=4, B=3),      # a three-scale expert: 128 / 3 under Python 2
             "g8": dict(res=8, var=0.0156, n_rads=4, n_experts=7, B=2)}     # the Nesti-Net default
    for name, c in cases.items():
        w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([c["res"]] * 3, c["var"]))
        P, B, S = 48, c["B"], c["n_rads"]
        pts = np.zeros((B, S * P, 3), np.float32)
        ne = rng.randint(P // 3, P + 1, size=(B, S)).astype(np.int32)
        for b in range(B):
            for s_ in range(S):
                x = rng.normal(size=(ne[b, s_], 3)) * rng.uniform(0.2, 0.6)
                x /= np.maximum(1.0, np.linalg.norm(x, axis=1, keepdims=True))
                pts[b, s_ * P:s_ * P + ne[b, s_]] = x
        mups = orc.mups_assemble(pts, w, mu, sg, ne, S).astype(np.float32)
        tf.reset_graph()
        env = dict(net_ns)
        env.update({"MuPS": tf.Tensor(mups), "bn_decay": None, "is_training": tf.Tensor(np.bool_(False)),
                    "weight_decay": 0.005, "n_experts": c["n_experts"], "n_gaussians": tf.Py2Int(len(w)),
                    "n_rads": tf.Py2Int(S), "expert_dict": None})
        exec(statements, env)
        prob, n_est = env["experts_prob"].a, env["n_est"].a
        assert prob.shape == (c["n_experts"], B) and n_est.shape == (c["n_experts"], B, 3)
        assert np.allclose(prob.sum(axis=0), 1.0, atol=1e-6) and np.all(np.isfinite(n_est))
        names = sorted(tf.created_variables)
        out.update({name + "_mups": mups, name + "_experts_prob": prob, name + "_n_est": n_est,
                    name + "_n_rads": np.int32(S), name + "_n_experts": np.int32(c["n_experts"]),
                    name + "_variables": np.array(["%s %s" % (n, "x".join(map(str, tf.created_variables[n]))) for n in names])})
        print("tf-emulated MoE %s: %d variables, %.1f M parameters, prob %s, n_est %s"
              % (name, len(names), sum(int(np.prod(tf.created_variables[n])) for n in names) / 1e6, prob.shape, n_est.shape))
    np.savez_compressed(os.path.join(HERE, "moe_tf_emulated.npz"), **out)


def make_rotation_reference():
    """utils/eulerangles.py::euler2mat and the augmentation of train_n_est_w_experts.py:262-272 run on the
    unmodified reference module (a py2 builtin, ``reduce``, is supplied from functools)."""
    import functools
    if REF_UTILS not in sys.path:
        sys.path.insert(0, REF_UTILS)
    import eulerangles as ref
    ref.reduce = functools.reduce
    rng = np.random.RandomState(0)
    angles = [2 * np.pi * rng.randn(3) for _ in range(5)] + [np.array([0.2, 0.0, 0.3]), np.array([0.0, 0.0, 0.3]),
                                                             np.zeros(3)]
    mats = [ref.euler2mat(z=a[0], y=a[1], x=a[2]) for a in angles]
    pts = rng.randn(4, 64, 3).astype(np.float32)
    nrm = rng.randn(4, 3).astype(np.float32)
    a = 2 * np.pi * rng.randn(3)
    R = np.transpose(ref.euler2mat(z=a[0], y=a[1], x=a[2]))
    rot = np.zeros(pts.shape, dtype=np.float32)
    rotn = np.zeros(nrm.shape, dtype=np.float32)
    for k in range(pts.shape[0]):                      # the loop of train_n_est_w_experts.py:267-271
        rot[k, ...] = np.dot(pts[k, ...].reshape((-1, 3)), R)
        rotn[k, ...] = np.dot(nrm[k, ...], R)
    np.savez(os.path.join(HERE, "rotation_reference.npz"), angles=np.array(angles), mats=np.array(mats), aug_angles=a,
             points=pts, normals=nrm, rotated_points=rot, rotated_normals=rotn)
    print("rotation reference: %d matrices, batch %s" % (len(mats), pts.shape))


if __name__ == "__main__":
    make_half1()
    make_dataset_interface()
    make_sampler_reference()
    make_evaluation_reference()
    make_half2_reference_numpy()
    make_rotation_reference()
    make_half2()
    make_half2_tf_emulated()
    make_moe_tf_emulated()
