"""GPU parity suite (-m gpu): every CUDA entry point, called through the reference-shaped host API
(which goes through the C-ABI), against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): integer index tensors bit-exact; fp32 `cent` bit-exact as well
(same IEEE operations in the same order); aggregated GridConv features within 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

from gridgcn_b200 import synth

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("nebidx", "nebidxmsk", "cent", "centmsk", "actual_centnum")


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _check5(got, want, what):
    for g, w, n in zip(got, want, NAMES):
        g = g.cpu().numpy()
        assert g.dtype == w.dtype and g.shape == w.shape, (what, n, g.dtype, g.shape, w.shape)
        if not np.array_equal(g, w):
            bad = np.argwhere(g != w)
            raise AssertionError("%s: %s differs at %d places, first %s: got %s want %s" % (
                what, n, len(bad), bad[0], g[tuple(bad[0])], w[tuple(bad[0])]))


GRID_CASES = [
    ("cfg1_P64_k3", 2, 1024, "surface",
     dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("cfg1_P32_k5", 2, 1024, "surface",
     dict(max_p_grid=32, max_o_grid=1024, kernel_size=5, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("seg8192_L0", 3, 8192, "surface",
     dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("seg8192_L1", 3, 1024, "surface",
     dict(max_p_grid=32, max_o_grid=256, kernel_size=3, loc=1, voxel_size=(0.133333,) * 3, grid_size=(15,) * 3)),
    ("overflow", 2, 4096, "ball",
     dict(max_p_grid=8, max_o_grid=100, kernel_size=3, loc=1, voxel_size=(0.25,) * 3, grid_size=(8,) * 3)),
    ("P128_k7_cls", 2, 1024, "surface",
     dict(max_p_grid=128, max_o_grid=128, kernel_size=7, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("one_voxel", 2, 128, "ball",
     dict(max_p_grid=128, max_o_grid=1, kernel_size=1, loc=1, voxel_size=(2.0,) * 3, grid_size=(1,) * 3)),
    ("loc0_aniso", 2, 700, "surface",
     dict(max_p_grid=16, max_o_grid=64, kernel_size=3, loc=0, voxel_size=(0.2, 0.25, 0.5), grid_size=(10, 8, 4))),
    ("big_cloud_global_path", 1, 20000, "ball",
     dict(max_p_grid=64, max_o_grid=512, kernel_size=3, loc=1, voxel_size=(0.1,) * 3, grid_size=(20,) * 3)),
]


@pytest.mark.parametrize("name,B,N,kind,kw", GRID_CASES, ids=[c[0] for c in GRID_CASES])
def test_gridify_and_knn_match_oracle(gg, cuda_dev, oracle_mod, name, B, N, kind, kw):
    data, npts = synth.make_batch(B, N, seed0=100, kind=kind, voxels=(kw["voxel_size"][0],))
    npts[-1, 0] = N - N // 7  # ragged
    kw = dict(kw, coord_shift=(1.0, 1.0, 1.0))
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    _check5(gg.Gridify(d, n, stride=1, **kw), oracle_mod.gridify(data, npts, **kw), name + "/Gridify")
    _check5(gg.GridifyKNN(d, n, stride=1, **kw), oracle_mod.gridify_knn(data, npts, **kw),
            name + "/GridifyKNN")
    _check5(gg.GridifyKNN(d, n, stride=1, dist_fma=True, **kw),
            oracle_mod.gridify_knn(data, npts, dist_fma=True, **kw), name + "/GridifyKNN fma")


STRICT_CASES = [c for c in GRID_CASES if c[0] in ("cfg1_P64_k3", "seg8192_L0", "seg8192_L1", "overflow",
                                                 "P128_k7_cls", "loc0_aniso", "big_cloud_global_path")] + [
    ("tiny_P4_heavy_overflow", 3, 900, "ball",
     dict(max_p_grid=4, max_o_grid=40, kernel_size=3, loc=1, voxel_size=(0.5,) * 3, grid_size=(4,) * 3)),
]


@pytest.mark.parametrize("name,B,N,kind,kw", STRICT_CASES, ids=[c[0] for c in STRICT_CASES])
def test_gridify_strict_reservoir_matches_oracle(gg, cuda_dev, oracle_mod, name, B, N, kind, kw):
    """The reference's K2 reservoir (gridify.cu:259-270) has a schedule-independent seed: reproduced exactly
    (XORWOW, insertion index, last writer wins), against the oracle's sequential replay."""
    data, npts = synth.make_batch(B, N, seed0=400, kind=kind, voxels=(kw["voxel_size"][0],))
    npts[-1, 0] = N - N // 5
    kw = dict(kw, coord_shift=(1.0, 1.0, 1.0))
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    want = oracle_mod.gridify(data, npts, strict_reservoir=True, **kw)
    _check5(gg.Gridify(d, n, stride=1, strict_reservoir=True, **kw), want, name + "/Gridify strict")
    keep_first = oracle_mod.gridify(data, npts, strict_reservoir=False, **kw)
    if name not in ("cfg1_P64_k3", "P128_k7_cls"):  # the others overflow: the two rules really differ
        assert not np.array_equal(want[0], keep_first[0])
    # integer-valued weights other than 1 (what later layers see: neighbour counts)
    data[..., 3] = np.random.default_rng(1).integers(1, 40, size=data.shape[:2]).astype(np.float32)
    d = _t(data, cuda_dev)
    _check5(gg.Gridify(d, n, stride=1, strict_reservoir=True, **kw),
            oracle_mod.gridify(data, npts, strict_reservoir=True, **kw), name + "/Gridify strict, weights")


@pytest.mark.parametrize("N,frac", [(8192, 0.6), (8192, 1.0), (40000, 0.7)], ids=["smem_build", "all_in_one_voxel", "multi_kernel_build"])
def test_dense_voxels_ranked_by_scans(gg, cuda_dev, oracle_mod, N, frac):
    """Voxel segments longer than 1024 points (clustered / duplicated points) are ranked by ordered block scans instead
    of the per-point count (grid_build.cuh rank_heavy_segments, both build paths): same table, bit-exact operators."""
    data, npts = synth.make_batch(2, N, seed0=900, kind="ball", voxels=(0.1,))
    rng = np.random.default_rng(5)
    for b in range(2):
        k = int(N * frac)
        sel = rng.permutation(N)[:k]
        data[b, sel, :3] = 0.033 * rng.random((k, 3)).astype(np.float32) + (0.31 if b == 0 else -0.47)  # inside one voxel
        if frac == 1.0:
            data[b, :, :3] = data[b, 0, :3]  # every point identical: exact ties everywhere
    kw = dict(max_p_grid=64, max_o_grid=512, kernel_size=3, loc=1, coord_shift=(1.0, 1.0, 1.0), voxel_size=(0.1,) * 3,
              grid_size=(20,) * 3)
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    _check5(gg.Gridify(d, n, stride=1, strict_reservoir=False, **kw), oracle_mod.gridify(data, npts, strict_reservoir=False, **kw),
            "dense/Gridify")
    _check5(gg.Gridify(d, n, stride=1, strict_reservoir=True, **kw), oracle_mod.gridify(data, npts, strict_reservoir=True, **kw),
            "dense/Gridify strict")
    _check5(gg.GridifyKNN(d, n, stride=1, **kw), oracle_mod.gridify_knn(data, npts, **kw), "dense/GridifyKNN")


def test_edge_cases(gg, cuda_dev, oracle_mod):
    data, npts = synth.make_batch(3, 64, seed0=1)
    npts[0, 0] = 0
    data[1, :, :3] += 10.0
    data[2, 32:] = data[2, :32]  # duplicates -> exact ties
    kw = dict(max_p_grid=4, max_o_grid=8, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.5,) * 3, grid_size=(4,) * 3)
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    _check5(gg.Gridify(d, n, **kw), oracle_mod.gridify(data, npts, **kw), "edge/Gridify")
    _check5(gg.GridifyKNN(d, n, **kw), oracle_mod.gridify_knn(data, npts, **kw), "edge/GridifyKNN")
    _check5(gg.Gridify(d, n, strict_reservoir=True, **kw), oracle_mod.gridify(data, npts, strict_reservoir=True, **kw),
            "edge/Gridify strict")
    for ks in (1, 3):  # empty cloud, cloud outside the grid, duplicates; kernel 1: a voxel only covers itself
        kc = dict(kw, kernel_size=ks, max_o_grid=3)
        _check5(gg.Gridify_occaware(d, n, seed=3, **kc), oracle_mod.gridify_occaware(data, npts, seed=3, **kc),
                "edge/occaware k%d" % ks)
    # B = 0 is a no-op
    out = gg.Gridify(d[:0], n[:0], **kw)
    assert out[0].shape == (0, 8, 4)


def test_chained_layers_feed_centres_back(gg, cuda_dev, oracle_mod):
    """Layer i+1 voxelises layer i's centres (ggcn_models_g.py:160): a 1-ulp difference in `cent`
    could flip a voxel, so the chain is checked end to end, weights w = neighbour counts."""
    data, npts = synth.make_batch(2, 8192, seed0=7)
    ladder = [(0.05, 40, 1024, 64), (0.133333, 15, 256, 32), (0.4, 5, 24, 32)]
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    hd, hn = data, npts
    for vox, grid, O, P in ladder:
        kw = dict(max_p_grid=P, max_o_grid=O, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
                  voxel_size=(vox,) * 3, grid_size=(grid,) * 3)
        got = gg.Gridify(d, n, **kw)
        want = oracle_mod.gridify(hd, hn, **kw)
        _check5(got, want, "chain vox %g" % vox)
        d, n = got[2], got[4]
        hd, hn = want[2], want[4]


CAS_CASES = [
    ("seg8192_L0", 3, 8192, "surface",
     dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("few_slots", 2, 2048, "surface",
     dict(max_p_grid=16, max_o_grid=256, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("clipped_aniso", 2, 600, "ball",
     dict(max_p_grid=8, max_o_grid=20, kernel_size=3, loc=1, voxel_size=(0.25, 0.25, 0.5), grid_size=(8, 8, 4))),
    ("kernel5", 2, 4096, "surface",
     dict(max_p_grid=32, max_o_grid=128, kernel_size=5, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("no_challengers", 2, 200, "surface",
     dict(max_p_grid=16, max_o_grid=512, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("cover_in_global", 1, 20000, "ball",
     dict(max_p_grid=16, max_o_grid=512, kernel_size=3, loc=0, voxel_size=(0.04,) * 3, grid_size=(50,) * 3)),
]


@pytest.mark.parametrize("name,B,N,kind,kw", CAS_CASES, ids=[c[0] for c in CAS_CASES])
def test_gridify_occaware_matches_oracle(gg, cuda_dev, oracle_mod, name, B, N, kind, kw):
    """Coverage-aware sampling: the CUDA path against the canonical-schedule restatement (both follow
    the SASS of the reference's cubins; parity with the reference itself is unpinned)."""
    data, npts = synth.make_batch(B, N, seed0=300, kind=kind, voxels=(kw["voxel_size"][0],))
    npts[-1, 0] = N - N // 9
    kw = dict(kw, coord_shift=(1.0, 1.0, 1.0))
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    for seed in (0, 123456789012345):
        _check5(gg.Gridify_occaware(d, n, stride=1, seed=seed, **kw),
                oracle_mod.gridify_occaware(data, npts, seed=seed, **kw), "%s/occaware seed %d" % (name, seed))
    _check5(gg.Gridify_occaware(d, n, stride=1, seed=9, knn_query=True, **kw),
            oracle_mod.gridify_occaware(data, npts, seed=9, knn_query=True, **kw), name + "/occaware knn")


def test_gridify_occaware_golden_and_coverage(gg, cuda_dev):
    z = np.load(os.path.join(GOLDEN, "gridify_occaware.npz"))
    kw = dict(max_p_grid=16, max_o_grid=256, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.05,) * 3, grid_size=(40,) * 3)
    d, n = _t(z["data"], cuda_dev), _t(z["npts"], cuda_dev)
    got = gg.Gridify_occaware(d, n, seed=2026, **kw)
    _check5(got, [z[k] for k in NAMES], "golden/occaware")
    # size-independent property at full size: CAS covers more occupied voxels than keep-first sampling
    data, npts = synth.make_batch(4, 8192, seed0=900)
    kw = dict(kw, max_p_grid=64, max_o_grid=1024)
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    cas = gg.Gridify_occaware(d, n, seed=1, **kw)[2].cpu().numpy()
    rvs = gg.Gridify(d, n, **kw)[2].cpu().numpy()
    for b in range(4):
        occ = set(map(tuple, np.floor((data[b, :, :3] + np.float32(1)) / np.float32(0.05)).astype(int)))

        def covered(cent):
            cv = np.floor((cent[b, :, :3] + np.float32(1)) / np.float32(0.05)).astype(int)
            cov = {(c[0] + i, c[1] + j, c[2] + k) for c in cv for i in (-1, 0, 1) for j in (-1, 0, 1)
                   for k in (-1, 0, 1)}
            return len(occ & cov) / len(occ)
        assert covered(cas) > 0.95 and covered(cas) > covered(rvs)


def test_gridify_up_matches_oracle(gg, cuda_dev, oracle_mod):
    for (nd, nu, P, ks, vox, grid) in ((24, 256, 5, 3, 0.4, 5), (256, 1024, 5, 3, 0.133333, 15),
                                       (1024, 8192, 5, 3, 0.05, 40), (300, 900, 40, 5, 0.25, 8),
                                       (300, 900, 128, 3, 0.5, 4), (64, 100, 3, 1, 0.25, 8)):
        down, dn = synth.make_batch(2, nd, seed0=21, voxels=(vox,))
        up, un = synth.make_batch(2, nu, seed0=31, voxels=(vox,))
        dn[1, 0], un[1, 0] = nd - nd // 5, nu - nu // 9
        up[0, 5, :3] = 5.0
        kw = dict(max_p_grid=P, max_o_grid=nu, kernel_size=ks, coord_shift=(1, 1, 1),
                  voxel_size=(vox,) * 3, grid_size=(grid,) * 3)
        want = oracle_mod.gridify_up(down, up, dn, un, **kw)
        got = gg.GridifyUp(_t(down, cuda_dev), _t(up, cuda_dev), _t(dn, cuda_dev), _t(un, cuda_dev), **kw)
        assert np.array_equal(got[0].cpu().numpy(), want[0]), (nd, nu, P, ks)
        assert np.array_equal(got[1].cpu().numpy(), want[1]), (nd, nu, P, ks)


def test_knn_and_ball_knn_match_oracle(gg, cuda_dev, oracle_mod):
    rng = np.random.default_rng(0)
    for (B, n, m) in ((2, 256, 24), (2, 1024, 256), (1, 8192, 1024), (2, 100, 3000)):
        unknown = rng.uniform(-1, 1, size=(B, n, 3)).astype(np.float32)
        known = rng.uniform(-1, 1, size=(B, m, 3)).astype(np.float32)
        known[0, m // 2:] = known[0, :m - m // 2]  # exact ties -> lowest index wins
        downnum = np.full((B, 1), m, np.int32)
        upnum = np.full((B, 1), n, np.int32)
        downnum[-1, 0], upnum[-1, 0] = m - m // 4, n - n // 3
        args = [_t(a, cuda_dev) for a in (unknown, known, downnum, upnum)]
        for k in (1, 3, 5, 8, 20):
            want = oracle_mod.knn(unknown, known, downnum, upnum, k=k)
            got = gg.contrib.KNN(*args, k=k).cpu().numpy()
            assert np.array_equal(got, want), ("knn", B, n, m, k)
        for k, radius in ((5, 0.3), (3, 0.05), (6, 1.02)):
            want = oracle_mod.ball_knn(unknown, known, downnum, upnum, k=k, radius=radius)
            got = gg.contrib.BallKNN(*args, k=k, radius=radius).cpu().numpy()
            assert np.array_equal(got, want), ("ball", B, n, m, k)
    with pytest.raises(gg._lib.GridGcnError):
        gg.contrib.BallKNN(*args, k=7, radius=0.5)


def test_knn_lattice_vector(gg, cuda_dev):
    pts = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], np.float32)
    d2 = ((pts - 1.0) ** 2).sum(1)
    faces = [i for i in range(27) if d2[i] == 1]
    edge0 = min(i for i in range(27) if d2[i] == 2)
    unknown = _t(np.array([[[1.0, 1.0, 1.0]]], np.float32), cuda_dev)
    known = _t(pts[None], cuda_dev)
    dn, un = _t(np.array([[27]], np.int32), cuda_dev), _t(np.array([[1]], np.int32), cuda_dev)
    assert gg.contrib.KNN(unknown, known, dn, un, k=8)[0, 0].tolist() == [13] + faces + [edge0]


def _rel_err(got, want):
    """max |got - want| / (|want| + rms(want)): elementwise relative error with the tensor's RMS as
    the floor for near-zero entries."""
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2))) + 1e-30
    return float(np.max(np.abs(got - want) / (np.abs(want) + rms)))


def _strict_err(got, want):
    """The elementwise reading of "1e-3 rel fp32" (VERDICT r01 item 7): relative error wherever
    |want| >= 1e-2 rms(want), absolute error in units of rms below that.  Returns (rel_big, abs_small_over_rms)."""
    got, want = got.astype(np.float64), want.astype(np.float64)
    rms = float(np.sqrt(np.mean(want ** 2))) + 1e-30
    big = np.abs(want) >= 1e-2 * rms
    d = np.abs(got - want)
    rel_big = float(np.max(d[big] / np.abs(want[big]))) if big.any() else 0.0
    abs_small = float(np.max(d[~big])) / rms if (~big).any() else 0.0
    return rel_big, abs_small


def _assert_feats(got, want, what):
    """Both tolerances, both numbers in the message: rms-floored 1e-3 and the strict elementwise one."""
    loose = _rel_err(got, want)
    rel_big, abs_small = _strict_err(got, want)
    assert loose <= 1e-3 and rel_big <= 1e-3 and abs_small <= 1e-3, \
        "%s: rms-floored %.3g, strict relative %.3g, small-entry absolute/rms %.3g" % (what, loose, rel_big, abs_small)


CONV_CASES = [
    # name, B, N, O, K, Cin, mlp
    ("l0_K8", 2, 256, 32, 8, 0, [16, 32]),
    ("l1_K8", 2, 256, 32, 8, 16, [16, 32]),
    ("l0_seg", 2, 2048, 256, 64, 0, [32, 32, 64]),
    ("l1_seg", 2, 512, 128, 32, 64, [64, 64, 128]),
    ("l3_seg", 2, 64, 16, 32, 256, [256, 256, 512]),
    ("K128", 1, 512, 16, 128, 32, [32, 64]),
    ("K5_decoder_like", 2, 128, 100, 5, 32, [128]),
    ("K16_single_stage", 2, 256, 40, 16, 8, [24]),
    # first-layer shapes of the compact (M=64, three CTAs per SM) kernel: K | 64, Cout <= 64, ragged centre counts
    ("l0_K32_C40", 2, 512, 50, 32, 0, [32, 32, 40]),
    ("l0_K16_C64", 2, 512, 77, 16, 0, [32, 64]),
    ("l0_K4_C24", 1, 128, 30, 4, 0, [8, 16, 24]),
    ("l0_K64_C64_narrow_f0", 1, 1024, 33, 64, 0, [16, 32, 64]),
    ("l0_K128_C64", 1, 1024, 21, 128, 0, [32, 32, 64]),  # one centre per tile: halves combined across the lane pair
    # first-layer shapes that must take the general kernel (Cout > 64, K not a divisor of 64)
    ("l0_C96", 1, 512, 40, 32, 0, [32, 96]),
    ("l0_K24", 1, 512, 40, 24, 0, [32, 32, 64]),
]


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name,B,N,O,K,Cin,mlp", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_gridconv_matches_oracle(gg, cuda_dev, oracle_mod, name, B, N, O, K, Cin, mlp, precision):
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    rng = np.random.default_rng(5)
    data, npts = synth.make_batch(B, N, seed0=60, voxels=(0.25,))
    kw = dict(max_p_grid=K, max_o_grid=O, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    nebidx, _, cent, centmsk, _ = oracle_mod.gridify_knn(data, npts, **kw)
    table = data if Cin == 0 else np.concatenate(
        [data, rng.uniform(0, 1, size=(B, N, Cin)).astype(np.float32)], axis=2)
    layer = gridconv.init_layer(np.random.default_rng(23), Cin, mlp, 10)
    want = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, layer)
    conv = gg.GridConv(layer, cuda_dev, precision=precision)
    got = conv(_t(table, cuda_dev), _t(nebidx, cuda_dev), _t(cent, cuda_dev), _t(centmsk, cuda_dev))
    got = got.cpu().numpy()
    assert np.array_equal(got[..., :4], want[..., :4])  # centre columns are copied
    _assert_feats(got[..., 4:], want[..., 4:], "%s/%s" % (name, precision))  # north_star: 1e-3 rel fp32
    assert np.abs(want[..., 4:]).max() > 0  # the comparison is not vacuous


CLS_CASES = [
    # name, B, N, O, K, Cin, pt_mlp, att_ele, att_full, localfdim, attfdim
    ("cls_l0", 2, 1024, 200, 64, 0, [64, 64, 128], [64, 128, 128], "next", 3, 4),
    ("cls_l1_wide_scratch", 2, 256, 64, 64, 128, [128, 128, 256], [128, 256, 256], "next", 3, 4),
    ("cls_l2_one_voxel_K128", 2, 128, 1, 128, 256, [256, 256, 512], [256, 512, 512], "next", 3, 4),
    ("att_full_last", 2, 256, 40, 16, 16, [16, 32], [8, 24, 32], "last", 3, 10),
    ("explicit_att_no_concat", 2, 256, 40, 8, 16, [32], [8, 16, 32], "", 0, 4),
    ("four_att_stages", 1, 256, 30, 8, 8, [16, 32], [8, 16, 24, 32], "next", 3, 4),
]


@pytest.mark.parametrize("case", CLS_CASES, ids=[c[0] for c in CLS_CASES])
def test_gridconv_classification_block(gg, cuda_dev, oracle_mod, case):
    """The classification flavour of the block (classification/models/gcn_module_g.py: geo vector in front
    of the gathered features, explicit attention widths, att_full concat) -- fp32 CUDA-core kernel."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    name, B, N, O, K, Cin, pt, att, att_full, localfdim, attfdim = case
    rng = np.random.default_rng(8)
    data, npts = synth.make_batch(B, N, seed0=70, voxels=(0.25,))
    if O == 1:   # group-all layer of the cls ladder: one voxel, kernel 1
        kw = dict(max_p_grid=K, max_o_grid=1, kernel_size=1, loc=1, coord_shift=(1, 1, 1),
                  voxel_size=(2.0,) * 3, grid_size=(1,) * 3)
    else:
        kw = dict(max_p_grid=K, max_o_grid=O, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
                  voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    nebidx, _, cent, centmsk, _ = oracle_mod.gridify(data, npts, **kw)
    table = data if Cin == 0 else np.concatenate(
        [data, rng.uniform(0, 1, size=(B, N, Cin)).astype(np.float32)], axis=2)
    layer = gridconv.init_layer(np.random.default_rng(29), Cin, pt, attfdim, att_ele_lst=att,
                                att_full=att_full, localfdim=localfdim)
    want = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, layer)
    for precision in ("fp32", "tf32x3"):  # CUDA-core kernel / chain of tensor-core row GEMMs (gridconv_cls_tc.cu)
        conv = gg.GridConv(layer, cuda_dev, precision=precision)
        got = conv(_t(table, cuda_dev), _t(nebidx, cuda_dev), _t(cent, cuda_dev), _t(centmsk, cuda_dev)).cpu().numpy()
        assert np.array_equal(got[..., :4], want[..., :4])
        _assert_feats(got[..., 4:], want[..., 4:], "%s/%s" % (name, precision))
    assert np.abs(want[..., 4:]).max() > 0
    if att_full or localfdim:  # the single-pass tf32 option exists for the segmentation block only: loud refusal
        conv = gg.GridConv(layer, cuda_dev, precision="tf32")
        with pytest.raises(gg._lib.GridGcnError):
            conv(_t(table, cuda_dev), _t(nebidx, cuda_dev), _t(cent, cuda_dev), _t(centmsk, cuda_dev))


@pytest.mark.parametrize("precision", ["fp32", "tf32x3"])
def test_classification_stack_matches_oracle(gg, cuda_dev, oracle_mod, precision):
    """The shipped ModelNet40 ladder (classification/configs/configs.yaml:44-68: kernel 7/3/1, O 1024/128/1,
    P 64/64/128, attfdim 4, localfdim 3, att_full next) end to end: Gridify indices bit-exact per layer,
    features within 1e-3."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import stack
    cfg = stack.cls1024_shipped()
    params = stack.init_params(cfg, seed=4)
    data, npts = synth.make_batch(2, cfg.num_points, seed0=210, voxels=cfg.voxels)
    enc = stack.GridGcnEncoder(cfg, params, cuda_dev, precision=precision)
    out = enc(_t(data, cuda_dev), _t(npts, cuda_dev), keep_trace=True)
    table, loc, num = data, data, npts
    for i, (l, p) in enumerate(zip(cfg.layers, params)):
        want = oracle_mod.gridify(loc, num, max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid,
                                  kernel_size=l.kernel_size, loc=cfg.loc, coord_shift=cfg.coord_shift,
                                  voxel_size=(l.voxel_size,) * 3, grid_size=(l.grid_size,) * 3)
        tr = enc.trace[i]
        _check5([tr[n] for n in NAMES], want, "cls1024_shipped layer %d" % i)
        table = gridconv_oracle.gridconv_layer(table, want[0], want[2], want[3], p, pre_relu=cfg.pre_relu)
        _assert_feats(tr["table"].cpu().numpy()[..., 4:], table[..., 4:], "cls1024_shipped/%s layer %d" % (precision, i))
        loc, num = want[2], want[4]
    assert out.shape == (2, 1, 4 + 512)


def test_gridconv_attention_variants_and_ball_misses(gg, cuda_dev, oracle_mod):
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    rng = np.random.default_rng(3)
    B, N, O, K, Cin = 2, 200, 50, 6, 12
    table = rng.uniform(-1, 1, size=(B, N, 4 + Cin)).astype(np.float32)
    nebidx = rng.integers(0, N, size=(B, O, K)).astype(np.int32)
    nebidx[:, :, -1] = -1  # BallKNN miss: take() clips after the batch offset (utils/ops.py:90-92)
    cent = rng.uniform(-1, 1, size=(B, O, 4)).astype(np.float32)
    centmsk = (rng.uniform(size=(B, O)) > 0.3).astype(np.float32)
    for attfdim in (10, 4, 3, 0):
        layer = gridconv.init_layer(np.random.default_rng(1), Cin, [20, 40], attfdim)
        want = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, layer, pre_relu=True)
        got = gg.GridConv(layer, cuda_dev)(_t(table, cuda_dev), _t(nebidx, cuda_dev),
                                           _t(cent, cuda_dev), _t(centmsk, cuda_dev)).cpu().numpy()
        assert _rel_err(got[..., 4:], want[..., 4:]) <= 1e-3, attfdim


def test_sub_g_update_reference_signature(gg, cuda_dev, oracle_mod):
    """The reference-shaped entry (gathered neighbours in (B,4+C,O,P)) gives the same features."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    z = np.load(os.path.join(GOLDEN, "gridconv_l1.npz"))
    layer = gridconv.init_layer(np.random.default_rng(23), 16, [16, 32], 10)
    neighbors = gridconv_oracle.batch_take_g(z["table"], z["nebidx"]).transpose(0, 3, 1, 2)
    cent = z["cent"]
    feats = gg.sub_g_update(_t(cent[:, :, :3].transpose(0, 2, 1), cuda_dev),
                            _t(cent[:, :, 3:4].transpose(0, 2, 1), cuda_dev),
                            _t(neighbors, cuda_dev), True, _t(z["centmsk"], cuda_dev), None, 10,
                            pt_mlp_lst=[16, 32], layer=layer)
    want = z["out"][..., 4:].transpose(0, 2, 1)
    assert feats.shape == want.shape
    assert _rel_err(feats.cpu().numpy(), want) <= 1e-3


def test_golden_fixtures_on_gpu(gg, cuda_dev):
    from gridgcn_b200 import gridconv
    kw1 = dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
               voxel_size=(0.05,) * 3, grid_size=(40,) * 3)
    z = np.load(os.path.join(GOLDEN, "gridify_cfg1.npz"))
    _check5(gg.Gridify(_t(z["data"], cuda_dev), _t(z["npts"], cuda_dev), strict_reservoir=False, **kw1),
            [z[n] for n in NAMES], "golden gridify_cfg1")
    z = np.load(os.path.join(GOLDEN, "gridifyknn_cfg1.npz"))
    _check5(gg.GridifyKNN(_t(z["data"], cuda_dev), _t(z["npts"], cuda_dev),
                          **dict(kw1, max_p_grid=32, kernel_size=5)),
            [z[n] for n in NAMES], "golden gridifyknn_cfg1")
    z = np.load(os.path.join(GOLDEN, "gridify_overflow.npz"))
    kwo = dict(max_p_grid=8, max_o_grid=100, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
               voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    _check5(gg.Gridify(_t(z["data"], cuda_dev), _t(z["npts"], cuda_dev), strict_reservoir=False, **kwo),
            [z[n] for n in NAMES], "golden overflow")
    _check5(gg.GridifyKNN(_t(z["data"], cuda_dev), _t(z["npts"], cuda_dev), **kwo),
            [z["knn_" + n] for n in NAMES], "golden overflow knn")
    z = np.load(os.path.join(GOLDEN, "gridify_up.npz"))
    got = gg.GridifyUp(_t(z["downdata"], cuda_dev), _t(z["updata"], cuda_dev), _t(z["downnum"], cuda_dev),
                       _t(z["upnum"], cuda_dev), max_p_grid=5, max_o_grid=1024, kernel_size=3,
                       coord_shift=(1, 1, 1), voxel_size=(0.133333,) * 3, grid_size=(15,) * 3)
    assert np.array_equal(got[0].cpu().numpy(), z["nebidx"])
    assert np.array_equal(got[1].cpu().numpy(), z["nebidxmsk"])
    for case, fn, extra in (("knn", gg.contrib.KNN, {}),
                            ("ballknn", gg.contrib.BallKNN, dict(radius=0.133333 * 3 * 1.7 / 2))):
        z = np.load(os.path.join(GOLDEN, case + ".npz"))
        got = fn(_t(z["unknown"], cuda_dev), _t(z["known"], cuda_dev), _t(z["downnum"], cuda_dev),
                 _t(z["upnum"], cuda_dev), k=5, **extra)
        assert np.array_equal(got.cpu().numpy(), z["idx"]), case
    for case, cin in (("gridconv_l0", 0), ("gridconv_l1", 16)):
        z = np.load(os.path.join(GOLDEN, case + ".npz"))
        layer = gridconv.init_layer(np.random.default_rng(23), cin, [16, 32], 10)
        got = gg.GridConv(layer, cuda_dev)(_t(z["table"], cuda_dev), _t(z["nebidx"], cuda_dev),
                                           _t(z["cent"], cuda_dev), _t(z["centmsk"], cuda_dev))
        assert _rel_err(got.cpu().numpy()[..., 4:], z["out"][..., 4:]) <= 1e-3, case
    z = np.load(os.path.join(GOLDEN, "gridify_strict.npz"))
    _check5(gg.Gridify(_t(z["data"], cuda_dev), _t(z["npts"], cuda_dev), strict_reservoir=True, **kwo),
            [z[n] for n in NAMES], "golden strict reservoir")
    z = np.load(os.path.join(GOLDEN, "gridconv_cls.npz"))
    layer = gridconv.init_layer(np.random.default_rng(31), 16, [16, 16, 32], 4, att_ele_lst=[16, 32, 32],
                                att_full="next", localfdim=3)
    got = gg.GridConv(layer, cuda_dev, precision="fp32")(_t(z["table"], cuda_dev), _t(z["nebidx"], cuda_dev),
                                                         _t(z["cent"], cuda_dev), _t(z["centmsk"], cuda_dev))
    assert _rel_err(got.cpu().numpy()[..., 4:], z["out"][..., 4:]) <= 1e-3, "golden cls block"


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_full_stack_matches_oracle(gg, cuda_dev, oracle_mod, precision):
    """The whole encoder (query -> GridConv per layer) against the oracle chain, index tensors
    bit-exact at every layer, features within 1e-3."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import stack
    cases = [(stack.tiny(8), 3), (stack.cls1024_4layer(32), 2), (stack.seg8192_4layer(16), 1),
             (stack.seg8192_shipped(), 1), (stack.seg8192_4layer(64), 2)]  # the headline shape in BOTH precisions
    if precision == "tf32x3":  # the product precision also covers BASELINE's batch of config 2 (B = 32), the K sweep
        cases += [(stack.cls1024_4layer(32), 32),             # of config 5 and the config-4 shape (B = 3 per GPU)
                  (stack.seg81920_shipped(), 3),
                  (stack.seg8192_4layer(32, "gridify"), 1), (stack.seg8192_4layer(128), 1)]
        cas = stack.seg8192_4layer(32, "occaware_knn")  # coverage-aware sampling feeding the GridConv ladder
        cas.cas_seed = 11
        cases += [(cas, 2), (stack.seg8192_shipped("occaware"), 1)]
    for cfg, B in cases:
        params = stack.init_params(cfg, seed=1)
        data, npts = synth.make_batch(B, cfg.num_points, seed0=200, voxels=cfg.voxels)
        enc = stack.GridGcnEncoder(cfg, params, cuda_dev, precision=precision)
        out = enc(_t(data, cuda_dev), _t(npts, cuda_dev), keep_trace=True)
        if cfg.query.startswith("occaware"):
            q = lambda *a, **kw: oracle_mod.gridify_occaware(*a, seed=cfg.cas_seed,  # noqa: E731
                                                             knn_query=cfg.query.endswith("knn"), **kw)
        else:
            q = oracle_mod.gridify_knn if cfg.query == "gridifyknn" else oracle_mod.gridify
        table, loc, num = data, data, npts
        for i, (l, p) in enumerate(zip(cfg.layers, params)):
            want = q(loc, num, max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid,
                     kernel_size=l.kernel_size, loc=cfg.loc, coord_shift=cfg.coord_shift,
                     voxel_size=(l.voxel_size,) * 3, grid_size=(l.grid_size,) * 3)
            tr = enc.trace[i]
            _check5([tr[n] for n in NAMES], want, "%s layer %d" % (cfg.name, i))
            table = gridconv_oracle.gridconv_layer(table, want[0], want[2], want[3], p,
                                                   pre_relu=cfg.pre_relu)
            _assert_feats(tr["table"].cpu().numpy()[..., 4:], table[..., 4:], "%s B=%d layer %d" % (cfg.name, B, i))
            loc, num = want[2], want[4]
        assert out.shape == (B, cfg.layers[-1].max_o_grid, 4 + cfg.layers[-1].pt_mlp_lst[-1])


def test_cuda_graph_replay_matches_eager(gg, cuda_dev):
    """The whole encoder forward captured into one CUDA graph replays bit-identically, also on new inputs."""
    from gridgcn_b200 import stack
    cfg = stack.cls1024_4layer(32)
    params = stack.init_params(cfg, seed=3)
    enc = stack.GridGcnEncoder(cfg, params, cuda_dev, precision="tf32x3")
    d0, n0 = synth.make_batch(4, cfg.num_points, seed0=10, voxels=cfg.voxels)
    d1, n1 = synth.make_batch(4, cfg.num_points, seed0=20, voxels=cfg.voxels)
    n1[2, 0] = 900
    replay = stack.capture_graph(enc, _t(d0, cuda_dev), _t(n0, cuda_dev))
    for d, n in ((d0, n0), (d1, n1), (d0, n0)):
        want = enc(_t(d, cuda_dev), _t(n, cuda_dev)).clone()
        got = replay(_t(d, cuda_dev), _t(n, cuda_dev))
        torch.cuda.synchronize()
        assert torch.equal(got, want)


def test_full_size_properties(gg, cuda_dev):
    """BASELINE sizes (B=16 x 8192 points, K=64) through size-independent properties: masks count
    the valid slots, every valid id is in range and lies inside the centre's 3x3x3 voxel
    neighbourhood, KNN rows are sorted by distance, centres are distinct voxels, idempotence."""
    B, N, O, P = 16, 8192, 1024, 64
    data, npts = synth.make_batch(B, N, seed0=300)
    d, n = _t(data, cuda_dev), _t(npts, cuda_dev)
    kw = dict(max_p_grid=P, max_o_grid=O, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.05,) * 3, grid_size=(40,) * 3)
    for fn in (gg.Gridify, gg.GridifyKNN):
        a = fn(d, n, **kw)
        b = fn(d, n, **kw)
        for x, y in zip(a, b):
            assert torch.equal(x, y)  # deterministic, unlike the reference
        nebidx, msk, cent, cmsk, num = [t.cpu().numpy() for t in a]
        assert nebidx.min() >= 0 and nebidx.max() < N
        assert np.array_equal(cmsk.sum(1).astype(np.int32), num[:, 0])
        vox = np.floor((data[..., :3] + np.float32(1.0)) / np.float32(0.05)).astype(np.int64)
        for bb in range(0, B, 5):
            nc = int(num[bb, 0])
            cv = np.floor((cent[bb, :nc, :3] + np.float32(1.0)) / np.float32(0.05)).astype(np.int64)
            assert len(np.unique(cv, axis=0)) == nc  # one centre per occupied voxel
            nv = vox[bb][nebidx[bb, :nc]]  # (nc, P, 3)
            assert np.abs(nv - cv[:, None, :]).max() <= 1
            if fn is gg.GridifyKNN:
                u = (cv + 0.5) * 0.05  # shifted frame (gridifyknn.cu:253-255), candidates raw
                dd = ((u[:, None, :] - data[bb][nebidx[bb, :nc]][..., :3]) ** 2).sum(-1)
                for o in range(0, nc, 97):
                    if len(set(nebidx[bb, o].tolist())) == P:  # no padding in this row
                        assert np.all(np.diff(dd[o]) >= -1e-6)


def test_gridconv_plain_tf32_is_close(gg, cuda_dev, oracle_mod):
    """GRIDGCN_PRECISION_TF32 (one tensor-core pass, operands rounded to tf32) is a speed option, not
    the parity mode: it must stay within 1e-2 of the oracle; TF32X3 is the mode held to 1e-3."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    rng = np.random.default_rng(5)
    B, N, O, K, Cin, mlp = 2, 512, 128, 32, 64, [64, 64, 128]
    data, npts = synth.make_batch(B, N, seed0=60, voxels=(0.25,))
    kw = dict(max_p_grid=K, max_o_grid=O, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    nebidx, _, cent, centmsk, _ = oracle_mod.gridify_knn(data, npts, **kw)
    table = np.concatenate([data, rng.uniform(0, 1, size=(B, N, Cin)).astype(np.float32)], axis=2)
    layer = gridconv.init_layer(np.random.default_rng(23), Cin, mlp, 10)
    want = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, layer)
    errs = {}
    for prec in ("tf32", "tf32x3"):
        got = gg.GridConv(layer, cuda_dev, precision=prec)(
            _t(table, cuda_dev), _t(nebidx, cuda_dev), _t(cent, cuda_dev), _t(centmsk, cuda_dev)).cpu().numpy()
        errs[prec] = _rel_err(got[..., 4:], want[..., 4:])
    print("rel err:", errs)
    assert errs["tf32x3"] <= 1e-3 and errs["tf32"] <= 1e-2
    assert errs["tf32x3"] < errs["tf32"]


def _rand_up_case(rng, B, Nd, O, K, cd, cu):
    f_last = rng.uniform(-1, 1, size=(B, Nd, 4 + cd)).astype(np.float32)
    f_this = rng.uniform(-1, 1, size=(B, O, 4 + cu)).astype(np.float32)
    nebidx = rng.integers(0, Nd, size=(B, O, K)).astype(np.int32)
    nebidx[:, ::3, -1] = -1  # BallKNN misses
    cent = f_this[:, :, :4].copy()
    msk = (rng.uniform(size=(B, O)) > 0.2).astype(np.float32)
    return f_last, nebidx, cent, f_this, msk


@pytest.mark.parametrize("precision", ["tf32x3", "fp32"])
def test_decoder_layer_matches_oracle(gg, cuda_dev, precision):
    """sub_g_update with center_ori_feats (decoder half, gcn_module_g_att.py:267-285)."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import gridconv
    rng = np.random.default_rng(11)
    for (B, Nd, O, K, cd, cu) in ((2, 24, 256, 5, 256, 128), (1, 256, 1024, 5, 128, 64), (2, 64, 300, 5, 32, 0)):
        f_last, nebidx, cent, f_this, msk = _rand_up_case(rng, B, Nd, O, K, cd, cu)
        layer = gridconv.init_up_layer(np.random.default_rng(3), cd, cu, [128], 10, [128], [128])
        for mask in (msk, None):
            want = gridconv_oracle.gridconv_up_layer(f_last, nebidx, cent, f_this, mask, layer)
            up = gg.GridConvUp(layer, cuda_dev, precision=precision)
            got = up(_t(f_last, cuda_dev), _t(nebidx, cuda_dev), _t(cent, cuda_dev), _t(f_this, cuda_dev),
                     _t(mask, cuda_dev) if mask is not None else None).cpu().numpy()
            assert np.array_equal(got[..., :4], want[..., :4])
            assert _rel_err(got[..., 4:], want[..., 4:]) <= 1e-3, (B, Nd, O, cd, cu, mask is None)


@pytest.mark.parametrize("B,fetch", [(2, "ballknn"), (12, "ballknn"), (2, "knn"), (2, "gridifyup")],
                         ids=["B2_ballknn", "B12_baseline_batch", "B2_real_knn", "B2_gridifyup"])
def test_seg_graph_encoder_decoder_head(gg, cuda_dev, oracle_mod, B, fetch):
    """BASELINE config 3: the shipped ScanNet-8192 graph, encoder + decoder + head, against the oracle
    chain (B = 12 is the shipped batch, configs.yaml:69): neighbour indices bit-exact at every decoder level
    (BallKNN; KNN for real_knn, ggcn_models_g.py:74-83; GridifyUp for up_neigh_fetch False), logits within 1e-3.
    The encoder query is Gridify in its default = the reference's reservoir."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import stack
    cfg = stack.seg8192_shipped()
    up = stack.UpCfg(neigh_fetch=fetch, voxel_size_lst=(0.4, 0.133333, 0.05), grid_size_lst=(5, 15, 40),
                     max_o_grid_lst=(256, 1024, 8192))
    params = stack.init_seg_params(cfg, up, seed=2)
    data, npts = synth.make_batch(B, cfg.num_points, seed0=400, voxels=cfg.voxels)
    if fetch == "ballknn" and B == 2:  # duplicates: exact distance ties through the whole graph (real loaders sample with replacement)
        data[1, 4096:] = data[1, :4096]
    net = stack.GridGcnSeg(cfg, up, params, cuda_dev)
    logits = net(_t(data, cuda_dev), _t(npts, cuda_dev), keep_trace=True).cpu().numpy()
    # oracle chain
    tables, cents, nums, masks = [data], [data], [npts], [None]
    table, loc, num = data, data, npts
    for l, p in zip(cfg.layers, params["enc"]):
        nebidx, _, cent, centmsk, num = oracle_mod.gridify(
            loc, num, max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid, kernel_size=l.kernel_size,
            loc=cfg.loc, coord_shift=cfg.coord_shift, voxel_size=(l.voxel_size,) * 3,
            grid_size=(l.grid_size,) * 3)
        table = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, p)
        tables.append(table); cents.append(cent); nums.append(num); masks.append(centmsk)
        loc = cent
    f_last, nl = tables[-1], len(cfg.layers)
    g_tables = [data] + [t["table"].cpu().numpy() for t in net.enc.trace]  # what the GPU decoder consumed
    misses = 0
    for i, p in enumerate(params["dec"]):
        dn, upl = nl - i, nl - i - 1
        if fetch == "ballknn":
            nebidx = oracle_mod.ball_knn(cents[upl][:, :, :3], cents[dn][:, :, :3], nums[dn], nums[upl],
                                         k=up.max_p_grid, radius=up.voxel_size_lst[i] * up.kernel_size * 1.7 / 2)
        elif fetch == "knn":
            nebidx = oracle_mod.knn(cents[upl][:, :, :3], cents[dn][:, :, :3], nums[dn], nums[upl], k=up.max_p_grid)
        else:
            nebidx, _ = oracle_mod.gridify_up(cents[dn], cents[upl], nums[dn], nums[upl], max_p_grid=up.max_p_grid,
                                              max_o_grid=cents[upl].shape[1], kernel_size=up.kernel_size,
                                              coord_shift=cfg.coord_shift, voxel_size=(up.voxel_size_lst[i],) * 3,
                                              grid_size=(up.grid_size_lst[i],) * 3)
        misses += int((nebidx < 0).sum())
        assert np.array_equal(net.trace[i]["nebidx"].cpu().numpy(), nebidx), "decoder level %d indices" % i
        mask = masks[upl] if i != nl - 1 else None
        got = net.trace[i]["table"].cpu().numpy()
        # (1) the operator on IDENTICAL inputs (north_star's statement): this level of the oracle fed with the
        #     tensors the GPU graph itself consumed -- strict elementwise tolerance;
        g_prev = g_tables[-1] if i == 0 else net.trace[i - 1]["table"].cpu().numpy()
        same_in = gridconv_oracle.gridconv_up_layer(g_prev, nebidx, cents[upl], g_tables[upl], mask, p)
        _assert_feats(got[..., 4:], same_in[..., 4:], "decoder level %d (identical inputs)" % i)
        # (2) the whole chain against the oracle chain: six layers of fp32-class rounding accumulate on both sides,
        #     so the rms-floored form of the tolerance
        f_last = gridconv_oracle.gridconv_up_layer(f_last, nebidx, cents[upl], tables[upl], mask, p)
        err = _rel_err(got[..., 4:], f_last[..., 4:])
        assert err <= 1e-3, "decoder level %d (chain): rel err %.3g" % (i, err)
    if fetch == "ballknn":
        assert misses > 0  # the take()-clip path of BallKNN misses (-1) is exercised through the full decoder
    assert logits.shape == (B, cfg.num_points, 21)
    _assert_feats(logits, gridconv_oracle.seg_head(net.trace[-1]["table"].cpu().numpy()[..., 4:], params["head"]),
                  "logits (identical inputs)")
    assert _rel_err(logits, gridconv_oracle.seg_head(f_last[..., 4:], params["head"])) <= 1e-3


@pytest.mark.parametrize("B,precision", [(2, "fp32"), (32, "fp32"), (32, "tf32x3")],
                         ids=["B2_fp32", "B32_baseline_batch_fp32", "B32_baseline_batch_tf32x3"])
def test_classification_graph_with_head(gg, cuda_dev, oracle_mod, B, precision):
    """BASELINE config 2's model family: the shipped ModelNet40 ladder + FC 512-256-40 head
    (classification/models/ggcn_models_g.py:25-35, :37-111) from the reference's YAML keys, at the batch
    BASELINE names (32) -- class scores within 1e-3 of the oracle chain."""
    from oracle import gridconv_oracle
    from gridgcn_b200 import stack, gridconv
    cfg = stack.cls1024_shipped()
    params = stack.init_cls_params(cfg, seed=6)
    data, npts = synth.make_batch(B, cfg.num_points, seed0=500, voxels=cfg.voxels)
    net = stack.GridGcnCls(cfg, params, cuda_dev, precision=precision)
    scores = net(_t(data, cuda_dev), _t(npts, cuda_dev)).cpu().numpy()
    probs = net(_t(data, cuda_dev), _t(npts, cuda_dev), probs=True).cpu().numpy()
    table, loc, num = data, data, npts
    for l, p in zip(cfg.layers, params["enc"]):
        want = oracle_mod.gridify(loc, num, max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid,
                                  kernel_size=l.kernel_size, loc=cfg.loc, coord_shift=cfg.coord_shift,
                                  voxel_size=(l.voxel_size,) * 3, grid_size=(l.grid_size,) * 3)
        table = gridconv_oracle.gridconv_layer(table, want[0], want[2], want[3], p, pre_relu=cfg.pre_relu)
        loc, num = want[2], want[4]
    x = table[:, 0, 4:].astype(np.float32)
    for j, st in enumerate(params["head"]):
        if j < 2:
            w, b = gridconv.fold_bn(st["weight"], st["bias"], st["gamma"], st["beta"], st["moving_mean"], st["moving_var"])
            x = np.maximum(x @ w.T + b, 0).astype(np.float32)
        else:
            x = (x @ np.asarray(st["weight"], np.float32).reshape(len(st["bias"]), -1).T + st["bias"]).astype(np.float32)
    assert scores.shape == (B, 40)
    _assert_feats(scores, x, "class scores")
    e = np.exp(x - x.max(-1, keepdims=True))
    assert np.allclose(probs, e / e.sum(-1, keepdims=True), rtol=2e-3, atol=1e-6)


ROWMLP_CASES = [  # rows, c1, c2, cout, relu_in, relu_out, scale, cent
    (1000, 132, 0, 128, False, True, False, False),   # decoder centre branch: [cent | feat] rows -> 128
    (300, 128, 128, 128, True, True, True, True),     # update MLP: concat of two views, pre-ReLU, mask, table layout
    (77, 128, 0, 21, False, False, False, False),     # segmentation head: 21 classes (tail columns)
    (5, 4, 0, 128, False, True, False, False),        # level 0 of the decoder: the input points themselves
    (260, 512, 0, 256, False, True, False, False),    # wide output (two weight rows per loader thread)
    (33, 6, 0, 10, False, True, False, False),        # widths not a multiple of 4: CUDA-core fallback inside the tc entry
]


@pytest.mark.parametrize("case", ROWMLP_CASES, ids=["r%d_%d+%d_to_%d" % c[:4] for c in ROWMLP_CASES])
def test_rowmlp_tensor_core_matches_numpy(gg, cuda_dev, case):
    """gridgcn_rowmlp_tc_fwd (persistent tcgen05 GEMM, tf32x3) against float64 numpy, and against the CUDA-core
    operator it replaces in the tensor-core precisions."""
    from gridgcn_b200 import gridconv
    rows, c1, c2, cout, relu_in, relu_out, use_scale, use_cent = case
    rng = np.random.default_rng(rows)
    a = rng.normal(size=(rows, 4 + c1)).astype(np.float32)       # in1 = a strided view [:, 4:]
    b = rng.normal(size=(rows, c2 + 8)).astype(np.float32) if c2 else None
    w = (rng.normal(size=(cout, c1 + c2)) / np.sqrt(c1 + c2)).astype(np.float32)
    bias = rng.normal(size=cout).astype(np.float32)
    scale = (rng.uniform(size=rows) > 0.3).astype(np.float32) if use_scale else None
    cent = rng.normal(size=(rows, 4)).astype(np.float32) if use_cent else None
    x = np.concatenate([a[:, 4:]] + ([b[:, :c2]] if c2 else []), axis=1).astype(np.float64)
    if relu_in:
        x = np.maximum(x, 0)
    want = x @ w.astype(np.float64).T + bias
    if relu_out:
        want = np.maximum(want, 0)
    if use_scale:
        want = want * scale[:, None]
    ta, tb = _t(a, cuda_dev), (_t(b, cuda_dev) if c2 else None)
    outs = []
    for tc in (True, False):
        out = gridconv.rowmlp(ta[:, 4:], tb[:, :c2] if c2 else None, _t(w, cuda_dev), _t(bias, cuda_dev),
                              relu_in=relu_in, relu_out=relu_out, row_scale=_t(scale, cuda_dev) if use_scale else None,
                              out_col=4 if use_cent else 0, cent=_t(cent, cuda_dev) if use_cent else None, tc=tc)
        out = out.cpu().numpy()
        if use_cent:
            assert np.array_equal(out[:, :4], cent)
            out = out[:, 4:]
        outs.append(out)
        assert out.shape == want.shape
        _assert_feats(out, want.astype(np.float32), "rowmlp tc=%s" % tc)
    assert np.allclose(outs[0], outs[1], rtol=1e-4, atol=1e-5)
