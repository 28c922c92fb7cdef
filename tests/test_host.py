"""CPU tests of the host logic: the C-ABI library loads and exports every symbol include/mups.h
declares, argument validation and error strings (no compute without a GPU), the grid GMM
mirror, query sharding, the world_size-2 gloo path of the slab gather, and that the product
never reaches into oracle/."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import nesti_net_b200 as mb
from nesti_net_b200 import _lib
from oracle import mups_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mups.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mups_[a-z0-9_]+)\s*\(", body)))
    assert declared, "no declarations parsed"
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), "libmups_b200.so does not export %s" % name
    assert sorted(_lib.EXPORTS) == declared
    assert L.mups_abi_version() == 1


def test_argument_validation_without_compute():
    L = _lib.load()
    h = ctypes.c_void_p()
    fp = ctypes.POINTER(ctypes.c_float)
    w = np.full(8, 1 / 8, np.float32)
    mu = np.zeros((8, 3), np.float32)
    sg = np.ones((8, 3), np.float32)
    args = lambda a: a.ctypes.data_as(fp)
    assert L.mups_gmm_create(ctypes.byref(h), args(w), args(mu), args(sg), 0) == _lib.MUPS_ERR_INVALID
    assert b"out of range" in L.mups_last_error()
    bad = sg.copy(); bad[3, 1] = 0
    assert L.mups_gmm_create(ctypes.byref(h), args(w), args(mu), args(bad), 8) == _lib.MUPS_ERR_INVALID
    assert b"sigma[3][1]" in L.mups_last_error()
    assert L.mups_3dmfv(None, None, None, 1, 1, 16, 1, None, None) == _lib.MUPS_ERR_INVALID
    assert L.mups_ball_query(None, None, 1, None, 1, 16, 0, None, None, None, None, None) == _lib.MUPS_ERR_INVALID
    assert L.mups_index_create(ctypes.byref(h), None, 10, 0.07, None) == _lib.MUPS_ERR_INVALID
    assert L.mups_set_option(b"no_such_option", 1) == _lib.MUPS_ERR_INVALID
    assert L.mups_set_option(b"boundary_cap", 0) == _lib.MUPS_ERR_INVALID
    assert L.mups_set_option(b"boundary_cap", 512) == _lib.MUPS_OK
    assert L.mups_set_option(b"fuse_candidates", -1) == _lib.MUPS_ERR_INVALID
    assert L.mups_set_option(b"fuse_candidates", 12288) == _lib.MUPS_OK
    assert L.mups_set_option(b"stats_variant", 65) == _lib.MUPS_ERR_INVALID
    assert L.mups_set_option(b"stats_variant", 0) == _lib.MUPS_OK
    with pytest.raises(ValueError):
        _lib.check(L.mups_set_option(b"boundary_cap", 100000))
    # the consumer's entry points validate their channel geometry before touching the device
    one = ctypes.c_void_p(1)            # non-NULL placeholder: validation fails before any dereference
    assert L.mups_conv3d_bn_relu(one, 4, 8, 64, 0, 60, one, 64, 64, 3, one, one, 1, one, 64, 0, None, None) == _lib.MUPS_ERR_INVALID
    assert b"multiples of 8" in L.mups_last_error()
    assert L.mups_conv3d_bn_relu(one, 4, 3, 64, 0, 64, one, 64, 64, 3, one, one, 1, one, 64, 0, None, None) == _lib.MUPS_ERR_INVALID
    assert b"volume edge" in L.mups_last_error()
    assert L.mups_conv1_split_bn_relu(one, 4, 8, 64, 0, 64, one, 64, 128, one, one, 64, one, 64, 0, 40, one, 64, 0, None) == _lib.MUPS_ERR_INVALID
    assert b"split" in L.mups_last_error()
    assert L.mups_avgpool3d_bn_relu(one, 4, 8, 64, 0, 64, 1, one, one, 1, one, 64, 0, None) == _lib.MUPS_ERR_INVALID
    assert b"identity" in L.mups_last_error()
    assert L.mups_pool3d(one, 4, 8, 64, 0, 64, 3, 1, one, None) == _lib.MUPS_ERR_INVALID        # max pool: window 2 only
    assert L.mups_split_bf16x3(one, 10, 80, 20, 20, one, 96, 0, 30, None) == _lib.MUPS_ERR_INVALID  # part width: a multiple of 8
    assert b"multiple of 8" in L.mups_last_error()
    assert L.mups_split_bf16x3(one, 10, 80, 70, 20, one, 96, 0, 32, None) == _lib.MUPS_ERR_INVALID  # columns [70, 90) of 80
    assert L.mups_split_bf16x3(one, 10, 80, 20, 20, one, 96, 8, 32, None) == _lib.MUPS_ERR_INVALID  # triplet [8, 104) of 96 channels
    assert b"triplet" in L.mups_last_error()
    assert L.mups_pool3d_bf16x3(one, 4, 8, 96, 0, 32, 3, 1, one, 96, 0, None) == _lib.MUPS_ERR_INVALID   # max pool: window 2 only
    assert L.mups_pool3d_bf16x3(one, 4, 8, 96, 8, 32, 3, 0, one, 96, 0, None) == _lib.MUPS_ERR_INVALID   # triplet [8, 104) of 96
    assert L.mups_conv3d_bn_relu_x3(one, 4, 8, 64, 0, 64, one, 64, 64, 3, one, one, 1, one, 192, 0, 24, None) == _lib.MUPS_ERR_INVALID
    assert b"split" in L.mups_last_error()
    assert L.mups_conv3d_bn_relu_x3(one, 4, 8, 64, 0, 64, one, 64, 64, 3, one, one, 1, one, 128, 0, 64, None) == _lib.MUPS_ERR_INVALID  # 192 channels needed
    assert L.mups_avgpool3d_f32_bn_relu_x3(one, 4, 8, 64, 1, one, one, 1, one, 192, 0, None) == _lib.MUPS_ERR_INVALID  # window >= 2
    assert L.mups_avgpool3d_f32_bn_relu_x3(one, 4, 8, 64, 3, one, one, 1, one, 192, 8, None) == _lib.MUPS_ERR_INVALID  # triplet [8, 200) of 192
    for name, top in ((b"pool_variant", 1), (b"conv_variant", 9)):
        assert L.mups_set_option(name, top + 1) == _lib.MUPS_ERR_INVALID and L.mups_set_option(name, 0) == _lib.MUPS_OK
    if not torch.cuda.is_available():
        # no CPU fallback: a valid request fails loudly with MUPS_ERR_CUDA
        assert L.mups_gmm_create(ctypes.byref(h), args(w), args(mu), args(sg), 8) == _lib.MUPS_ERR_CUDA
        assert b"CUDA" in L.mups_last_error() or b"device" in L.mups_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_python_api_fails_loudly_without_gpu():
    w, mu, sg = orc.gmm_feed(*orc.get_3d_grid_gmm([3] * 3, 0.1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.tf_util.get_3dmfv_n_est(np.zeros((1, 8, 3), np.float32), w, mu, sg, n_original_points=[4])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.PointIndex(np.zeros((10, 3), np.float32))


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "nesti-net_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert "libmups_oracle" not in text and "c_oracle" not in text, f


def test_grid_gmm_mirror_is_bit_identical():
    for n, var in [(3, 0.11), (5, 0.04), (8, 0.0156), (16, 0.00390625)]:
        g = mb.get_3d_grid_gmm([n] * 3, var)
        w, mu, cov = orc.get_3d_grid_gmm([n] * 3, var)
        assert np.array_equal(g.weights_, w) and np.array_equal(g.means_, mu) and np.array_equal(g.covariances_, cov)
    g = mb.get_3d_grid_gmm([2, 3, 4], 0.1)
    w, mu, cov = orc.get_3d_grid_gmm([2, 3, 4], 0.1)
    assert np.array_equal(g.means_, mu)


def test_shard_bounds():
    sb = mb.dist.shard_bounds
    assert list(sb(10, 3)) == [0, 4, 7, 10]
    assert list(sb(0, 4)) == [0, 0, 0, 0, 0]
    assert list(sb(3, 8)) == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    rng = np.random.RandomState(0)
    for world in (1, 2, 4, 8):
        w = rng.gamma(0.5, 1.0, 100000)            # heavy-tailed work estimate
        b = sb(len(w), world, w)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
        loads = np.add.reduceat(w, b[:-1])[: world]
        assert loads.max() <= w.sum() / world + w.max() + 1e-9
    q = np.arange(101)
    parts = [mb.dist.shard_queries(q, r, 4)[0] for r in range(4)]
    assert np.array_equal(np.concatenate(parts), q)
    with pytest.raises(ValueError):
        sb(5, 2, [1, 2, 3])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_queries, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q = np.arange(n_queries, dtype=np.int64) * 3
        weights = (np.arange(n_queries) % 7 + 1).astype(np.float64)       # unequal shards
        mine, lo, hi = mb.dist.shard_queries(q, rank, world, weights)
        # a stand-in slab (the product computes slabs on the GPU; here only the plumbing is under test)
        slab = torch.from_numpy(np.stack([mine.astype(np.float32), mine.astype(np.float32) ** 2], 1)).reshape(len(mine), 2)
        full = mb.dist.gather_slabs(slab)
        expect = torch.from_numpy(np.stack([q.astype(np.float32), q.astype(np.float32) ** 2], 1))
        ret[rank] = bool(torch.equal(full, expect)) and (hi - lo) == len(mine)
    finally:
        dist.destroy_process_group()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` needs no GPU: one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mups_query_points_per_s" and d["unit"] == "query points/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]")


def test_numa_binding_helpers_degrade_gracefully():
    assert mb.dist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    info = mb.dist.gpu_numa_info(0)                       # no GPU here: no PCI address, node -1
    assert info["numa_node"] == -1 or info["cpus"]
    before = os.sched_getaffinity(0)
    applied = mb.dist.bind_to_gpu_numa_node(0)
    assert applied or os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_gloo_world_size_2_slab_gather():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(world, _free_port(), 1001, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_samplers():
    class DS(object):
        shape_names = ["a", "b", "c"]
        shape_patch_count = [10, 5, 20]
    ds = DS()
    seq = mb.pcpnet_dataset.SequentialPointcloudPatchSampler(ds)
    assert list(seq) == list(range(35)) and len(seq) == 35
    rnd = mb.pcpnet_dataset.RandomPointcloudPatchSampler(ds, patches_per_shape=8, seed=1, identical_epochs=True)
    a, b = list(rnd), list(rnd)
    assert len(a) == len(rnd) == 8 + 5 + 8 and len(set(a)) == len(a) and a == b and max(a) < 35
    ssr = mb.pcpnet_dataset.SequentialShapeRandomPointcloudPatchSampler(ds, 8, seed=1, sequential_shapes=True)
    idx = list(ssr)
    assert len(idx) == 21 and all(i < 10 for i in idx[:8]) and all(10 <= i < 15 for i in idx[8:13]) and all(i >= 15 for i in idx[13:])


def test_samplers_against_reference_run(golden_dir):
    """Same index sequences as the reference's samplers (two epochs each, with and without identical_epochs)."""
    ref = np.load(os.path.join(golden_dir, "sampler_reference.npz"))

    class DS(object):
        shape_names = ["a", "b", "c", "d"]
        shape_patch_count = [int(c) for c in ref["shape_patch_count"]]
    ds = DS()
    P = mb.pcpnet_dataset
    assert np.array_equal(list(P.SequentialPointcloudPatchSampler(ds)), ref["sequential"])
    for ident in (False, True):
        r = P.RandomPointcloudPatchSampler(ds, patches_per_shape=8, seed=3627473, identical_epochs=ident)
        assert np.array_equal([list(r), list(r)], ref["random_ident%d" % ident])
        for seq_shapes in (False, True):
            q = P.SequentialShapeRandomPointcloudPatchSampler(ds, 8, seed=3627473, sequential_shapes=seq_shapes,
                                                              identical_epochs=ident)
            assert np.array_equal([list(q), list(q)], ref["shape_random_ident%d_seq%d" % (ident, seq_shapes)])
            assert np.array_equal(np.concatenate([np.asarray(v, np.int64) for v in q.shape_patch_inds]),
                                  ref["shape_random_ident%d_seq%d_local" % (ident, seq_shapes)])


def test_dataset_rejects_options_off_the_hot_path(tmp_path):
    (tmp_path / "list.txt").write_text("cloud\n")
    kw = dict(root=str(tmp_path), shape_list_filename="list.txt", patch_radius=[0.05], points_per_patch=16,
              patch_features=[], seed=1)
    with pytest.raises(NotImplementedError):
        mb.pcpnet_dataset.PointcloudPatchDataset(use_pca=True, **kw)
    with pytest.raises(ValueError):
        mb.pcpnet_dataset.PointcloudPatchDataset(use_pca=False, center="bogus", **kw)
    with pytest.raises(ValueError):
        mb.pcpnet_dataset.PointcloudPatchDataset(use_pca=False, **dict(kw, patch_features=["bogus"]))
    with pytest.raises(ValueError):
        mb.provider.get_data_loader(outputs=["bogus"], indir=str(tmp_path), dataset_name="list.txt")


def test_rotation_augmentation_matches_reference_run(golden_dir):
    """euler2mat and the per-batch rotation of train_n_est_w_experts.py:262-272, against outputs of the
    reference's own utils/eulerangles.py (tests/golden/make_golden.py): bit-identical."""
    ref = np.load(os.path.join(golden_dir, "rotation_reference.npz"))
    for a, m in zip(ref["angles"], ref["mats"]):
        assert np.array_equal(mb.provider.euler2mat(z=a[0], y=a[1], x=a[2]), m)
    rp, rn, R = mb.provider.rotation_augmentation(ref["points"], ref["normals"], angles=ref["aug_angles"])
    assert rp.dtype == torch.float32 and np.array_equal(rp.numpy(), ref["rotated_points"])
    assert np.array_equal(rn.numpy(), ref["rotated_normals"])
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
    # random angles: 2 pi randn(3) from the supplied generator
    _, _, R2 = mb.provider.rotation_augmentation(ref["points"], ref["normals"], rng=np.random.RandomState(5))
    a = 2 * np.pi * np.random.RandomState(5).randn(3)
    assert np.array_equal(R2, mb.provider.euler2mat(z=a[0], y=a[1], x=a[2]).T)


def test_evaluation_metrics_against_reference_run(golden_dir):
    """RMS / PGP5 / PGP10 of utils/evaluate.py, run as a script on synthetic predictions by make_golden.py
    (sparse and dense prediction files, unnormalised and sign-flipped normals): the mirror gives the same numbers
    (the reference evaluates in float32, the mirror in float64)."""
    ref = np.load(os.path.join(golden_dir, "evaluation_reference.npz"))
    recs = []
    for k, name in enumerate(ref["names"]):
        recs.append(mb.evaluate.evaluate_shape(ref[name + "_pred"], ref[name + "_gt"], ref[name + "_pidx"]))
        assert abs(recs[-1]["rms"] - ref["rms"][k]) < 2e-4 * ref["rms"][k]
        assert recs[-1]["pgp5"] == pytest.approx(ref["pgp5"][k], abs=1e-12)
        assert recs[-1]["pgp10"] == pytest.approx(ref["pgp10"][k], abs=1e-12)
        assert recs[-1]["n"] == 120
    assert abs(np.mean([r["rms"] for r in recs]) - ref["avg_rms"]) < 2e-4 * ref["avg_rms"]
    assert abs(np.mean([r["rms_o"] for r in recs]) - ref["avg_rms_o"]) < 2e-4 * ref["avg_rms_o"]
    assert np.mean([r["pgp10"] for r in recs]) == pytest.approx(float(ref["avg_pgp10"]), abs=1e-12)


def test_evaluation_metrics():
    """RMS angle / PGP5 / PGP10 as defined in the reference's utils/evaluate.py:133-154."""
    ev = mb.evaluate
    gt = np.array([[0, 0, 1.0], [0, 0, 2.0], [1.0, 0, 0], [0, 1.0, 0]])
    pred = np.array([[0, 0, -3.0], [0, np.sin(np.deg2rad(7)), np.cos(np.deg2rad(7))], [1.0, 0, 0], [0, 0, 1.0]])
    ang = ev.angle_errors_deg(pred, gt)
    assert np.allclose(ang, [0, 7, 0, 90], atol=1e-9)
    assert np.allclose(ev.angle_errors_deg(pred, gt, oriented=True), [180, 7, 0, 90], atol=1e-6)
    assert abs(ev.rms_angle(pred, gt) - np.sqrt((49 + 8100) / 4.0)) < 1e-9
    rec = ev.evaluate_shape(pred, gt)
    assert rec["pgp5"] == 0.5 and rec["pgp10"] == 0.75 and rec["n"] == 4 and rec["rms_o"] > rec["rms"]
    sparse = ev.evaluate_shape(pred, gt, pidx=[1, 2])
    assert sparse["n"] == 2 and abs(sparse["rms"] - np.sqrt(49 / 2.0)) < 1e-9


def test_experts_net_tf_semantics_cpu():
    """The PyTorch restatement of the MoE consumer: shapes, TF 'SAME' pooling semantics,
    expert-to-scale assignment (models/experts_n_est.py:82-103)."""
    from nesti_net_b200.experts_net import (ExpertsNormalEstimator, angular_rms_deg, avg_pool_same, max_pool_same)
    torch.manual_seed(0)
    y = torch.arange(27.).reshape(1, 1, 3, 3, 3)
    ref = torch.zeros_like(y)
    for i in range(3):
        for j in range(3):
            for k in range(3):
                ref[0, 0, i, j, k] = y[0, 0, i:min(i + 2, 3), j:min(j + 2, 3), k:min(k + 2, 3)].mean()
    assert torch.allclose(avg_pool_same(y, 2), ref)                  # even kernel: window extends after, valid cells only
    assert torch.equal(avg_pool_same(y, 1), y)
    mp = max_pool_same(y, 3, 2)                                       # 3 -> ceil(3/2) = 2, padded by one on both sides
    assert tuple(mp.shape) == (1, 1, 2, 2, 2) and mp[0, 0, 0, 0, 0] == y[0, 0, :2, :2, :2].max() and mp.max() == 26
    net = ExpertsNormalEstimator(n_rads=2, n_gaussians=27, n_experts=5).eval()
    assert net.expert_dict == {0: [0], 1: [0], 2: [1], 3: [1], 4: [0, 1]}
    # the 3^3 expert is conv_net_3g unchanged (no width divider, :276-277); the 8^3 expert divides (:254)
    assert net.expert_conv[4].mods[0].one.conv.in_channels == 40 and net.expert_conv[4].mods[0].one.conv.out_channels == 128
    net8 = ExpertsNormalEstimator(n_rads=4, n_gaussians=512, n_experts=7)
    assert net8.expert_dict[6] == [0, 1, 2, 3] and net8.expert_conv[6].mods[0].one.conv.out_channels == 32
    assert net8.expert_conv[0].mods[0].one.conv.out_channels == 128 and net8.gate_conv.out_features == 1536
    assert net.expert_conv[2].mods[0].one.conv.in_channels == 20
    x = torch.randn(3, 3, 3, 3, 40) * 0.05
    with torch.no_grad():
        prob, n_est = net(x)
        normal, expert, probs = net.predict(x)
    assert tuple(prob.shape) == (5, 3) and tuple(n_est.shape) == (5, 3, 3) and torch.allclose(prob.sum(0), torch.ones(3))
    assert tuple(normal.shape) == (3, 3) and torch.equal(expert, prob.argmax(0)) and tuple(probs.shape) == (3, 5)
    # an expert only sees its own scales' channels
    x2 = x.clone(); x2[..., 20:] += 1.0
    with torch.no_grad():
        _, n2 = net(x2)
    assert torch.equal(n2[0], n_est[0]) and not torch.equal(n2[2], n_est[2])
    assert angular_rms_deg(normal, -normal) < 1e-9 and abs(angular_rms_deg(torch.tensor([[1., 0, 0]]), torch.tensor([[0., 1, 0]])) - 90) < 1e-9
    with pytest.raises(ValueError):
        ExpertsNormalEstimator(n_rads=2, n_gaussians=125)


def _restore_reference_variables(net, golden_dir):
    import sys
    if golden_dir not in sys.path:
        sys.path.insert(0, golden_dir)
    import moe_weights
    shapes = {}

    def get(name):
        dst = next(d for n, d, _ in net.tf_variables() if n == name)
        layout = next(l for n, _, l in net.tf_variables() if n == name)
        shp = tuple(dst.shape)
        if layout == "conv":
            shp = shp[2:] + (shp[1], shp[0])
        elif layout == "fc":
            shp = shp[::-1]
        shapes[name] = shp
        return moe_weights.variable_value(name, shp)
    net.load_tf_variables(get)
    return shapes


@pytest.mark.parametrize("case", ["g3", "g3s3", "g8"])
def test_experts_net_against_reference_text_on_emulated_tf(golden_dir, case):
    """experts_net.ExpertsNormalEstimator against the REFERENCE'S OWN network text (models/experts_n_est.py:78-106,155-310
    with the layers of utils/tf_util.py:254-351,406-495) executed on the numpy emulation of the primitive TF ops
    (tests/golden/moe_tf_emulated.npz, make_golden.py::make_moe_tf_emulated): the variables the reference's graph creates
    (names and shapes), the gate probabilities and every expert's normal.  g3s3 has a three-scale expert, whose width is
    128 / 3 = 42 under the reference's Python 2; g8 is the Nesti-Net default (4 scales, 7 experts, 8^3)."""
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg, canonical_tf_names
    g = np.load(os.path.join(golden_dir, "moe_tf_emulated.npz"))
    mups = torch.from_numpy(g[case + "_mups"])
    net = ExpertsNormalEstimator(int(g[case + "_n_rads"]), int(np.prod(mups.shape[1:4])), int(g[case + "_n_experts"])).eval()
    shapes = _restore_reference_variables(net, golden_dir)
    ours = sorted("%s %s" % (n, "x".join(map(str, s))) for n, s in shapes.items())
    assert ours == sorted(g[case + "_variables"].tolist()), "variable names / shapes differ from the reference graph's"
    with torch.no_grad():
        prob, n_est = net(mups)
    ref_prob, ref_n = g[case + "_experts_prob"], g[case + "_n_est"]
    assert np.allclose(prob.numpy(), ref_prob, rtol=1e-4, atol=1e-6), float(np.abs(prob.numpy() - ref_prob).max())
    scale = float(np.abs(ref_n).max())
    assert np.allclose(n_est.numpy(), ref_n, rtol=1e-4, atol=1e-4 * scale), float(np.abs(n_est.numpy() - ref_n).max())
    assert angular_rms_deg(n_est.reshape(-1, 3), torch.from_numpy(ref_n).reshape(-1, 3)) < 1e-2
    assert np.array_equal(prob.numpy().argmax(0), ref_prob.argmax(0))
    # a real TF 1.x checkpoint names the moving statistics after the ExponentialMovingAverage shadow variables
    ck = {"fc1noise/bn/fc1noise/bn/moments/Squeeze/ExponentialMovingAverage:0": 1,
          "fc1noise/bn/fc1noise/bn/moments/Squeeze_1/ExponentialMovingAverage": 2, "fc1noise/weights": 3}
    assert canonical_tf_names(ck) == {"fc1noise/bn/moving_mean": 1, "fc1noise/bn/moving_variance": 2, "fc1noise/weights": 3}


def test_bf16x3_engine_logic_on_emulated_ops(monkeypatch):
    """Host logic of the consumer's bf16x3 mode (moe_engine.TensorCoreExperts(precision="bf16x3")): segment / triplet
    bookkeeping, the [w_hi | w_hi | w_lo] weight expansion, channel offsets of every split -- with the three device entry
    points it is built from (mups_conv3d_bn_relu with fp32 output, mups_split_bf16x3, mups_pool3d_bf16x3) replaced by torch
    emulations of their documented semantics.  The emulated forward must agree with the fp32 network to fp32-grade accuracy
    (plain bf16 is three orders of magnitude away).  The real kernels are checked on the GPU (tests/test_gpu.py)."""
    import torch.nn.functional as F
    from nesti_net_b200 import moe_engine as me
    from nesti_net_b200.experts_net import ExpertsNormalEstimator, angular_rms_deg, avg_pool_same

    def emu_conv(x, cin_off, cin, layer, out=None, cout_off=0, out_f32=None):
        assert out is None and out_f32 is not None, "in bf16x3 mode the plain entry point is only used for its fp32 output"
        assert x.dtype == torch.bfloat16 and cin % 8 == 0 and cin_off % 8 == 0 and cin <= layer.cin_pad
        xs = x.float()[..., cin_off:cin_off + cin]
        if xs.ndim == 2:
            xs = xs[:, None, None, None, :]
        k = layer.k
        w = layer.w.float()[:, :, :cin].reshape(k, k, k, layer.cout_pad, cin).permute(3, 4, 0, 1, 2)
        pl, pr = (k - 1) // 2, k - 1 - (k - 1) // 2                       # TF 'SAME': the smaller half first
        y = F.conv3d(F.pad(xs.permute(0, 4, 1, 2, 3), (pl, pr, pl, pr, pl, pr)), w)
        y = y.permute(0, 2, 3, 4, 1).reshape(-1, layer.cout_pad) * layer.scale + layer.shift
        out_f32.copy_(torch.relu(y) if layer.relu else y)
        return out_f32

    def put_triplet(dst2d, off, w, v):
        hi = v.to(torch.bfloat16)
        lo = (v - hi.float()).to(torch.bfloat16)
        dst2d[:, off:off + w], dst2d[:, off + w:off + 2 * w], dst2d[:, off + 2 * w:off + 3 * w] = hi, lo, hi

    def emu_split(src, src_off, w_src, dst, dst_off, w_dst):
        assert src.dtype == torch.float32 and dst.dtype == torch.bfloat16 and w_dst % 8 == 0 and dst_off % 8 == 0
        assert dst_off + 3 * w_dst <= dst.shape[-1] and src_off + w_src <= src.shape[-1]
        v = torch.zeros((src.shape[0], w_dst))
        v[:, :w_src] = src[:, src_off:src_off + w_src]
        put_triplet(dst.view(-1, dst.shape[-1]), dst_off, w_dst, v)

    def emu_pool(x, c_off, w, k, is_max, y, y_off):
        assert c_off % 8 == 0 and w % 8 == 0 and c_off + 3 * w <= x.shape[-1] and y_off + 3 * w <= y.shape[-1]
        v = (x[..., c_off:c_off + w].float() + x[..., c_off + w:c_off + 2 * w].float()).permute(0, 4, 1, 2, 3)
        assert torch.equal(x[..., c_off:c_off + w], x[..., c_off + 2 * w:c_off + 3 * w]), "third part of a triplet is hi again"
        p = F.max_pool3d(v, 2, 2) if is_max else avg_pool_same(v, k)
        put_triplet(y.view(-1, y.shape[-1]), y_off, w, p.permute(0, 2, 3, 4, 1).reshape(-1, w))

    def emu_avgpool_f32(src, B, D, c, k, scale, shift, relu, out, y_off):
        assert src.dtype == torch.float32 and tuple(src.shape) == (B * D ** 3, c) and k >= 2 and y_off + 3 * c <= out.shape[-1]
        p = avg_pool_same(src.view(B, D, D, D, c).permute(0, 4, 1, 2, 3), k).permute(0, 2, 3, 4, 1).reshape(-1, c) * scale + shift
        put_triplet(out.view(-1, out.shape[-1]), y_off, c, torch.relu(p) if relu else p)

    def emu_conv_x3(x, cin_off, cin, layer, out, cout_off, split=None):
        split = layer.cout_pad if split is None else split
        assert split % 16 == 0 and 16 <= split <= layer.cout_pad and cout_off % 8 == 0 and cout_off + 3 * layer.cout_pad <= out.shape[-1]
        f = emu_conv(x, cin_off, cin, layer, None, 0, torch.empty((x.numel() // x.shape[-1], layer.cout_pad)))
        o2 = out.view(-1, out.shape[-1])
        put_triplet(o2, cout_off, split, f[:, :split])
        if split < layer.cout_pad:
            put_triplet(o2, cout_off + 3 * split, layer.cout_pad - split, f[:, split:])

    monkeypatch.setattr(me, "conv3d_x3", emu_conv_x3)
    monkeypatch.setattr(me, "avgpool_f32_x3", emu_avgpool_f32)
    monkeypatch.setattr(me, "conv3d_bn_relu", emu_conv)
    monkeypatch.setattr(me, "split_x3", emu_split)
    monkeypatch.setattr(me, "pool3d_x3", emu_pool)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)      # the constructor refuses to run without a device

    torch.manual_seed(7)
    net = ExpertsNormalEstimator(n_rads=2, n_gaussians=512, n_experts=3).eval()     # experts on scale 0, scale 1, both scales
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm1d)):
                m.running_mean.normal_(0, 0.05); m.running_var.uniform_(0.6, 1.5); m.weight.uniform_(0.7, 1.3); m.bias.normal_(0, 0.05)
    mups = (torch.rand(2, 8, 8, 8, 40) - 0.5) * 0.2
    with torch.no_grad():
        prob_ref, n_ref = net(mups)
    with pytest.raises(ValueError):
        me.TensorCoreExperts(net, device="cpu", precision="fp8")
    tc = me.TensorCoreExperts(net, device="cpu", precision="bf16x3")
    assert tc.gate.out_segs is not None and tc.gate.steps[0].one.cin_pad == 3 * 64
    prob, n_est = tc.forward(mups)
    assert tuple(prob.shape) == tuple(prob_ref.shape) and tuple(n_est.shape) == tuple(n_ref.shape)
    rms = angular_rms_deg(n_est.reshape(-1, 3), n_ref.reshape(-1, 3))
    rel = float((n_est - n_ref).abs().max() / n_ref.abs().max())
    dprob = float((prob - prob_ref).abs().max())
    print("bf16x3 (emulated ops) vs fp32 network: angular RMS %.3g deg, max rel %.3g, max |dprob| %.3g" % (rms, rel, dprob))
    assert rms < 5e-3 and rel < 1e-4 and dprob < 1e-5
