"""CPU suite, part 1: the oracle against its independent numpy twin, the reference's only
known-answer geometry (the 27-point lattice of utils/ops.py:282-299) and the committed golden
fixtures.  No GPU, no product code under test here."""
import os

import numpy as np
import pytest

from oracle import np_twin
from gridgcn_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
NAMES = ("nebidx", "nebidxmsk", "cent", "centmsk", "actual_centnum")


def _same(a, b, what):
    for x, y, n in zip(a, b, NAMES):
        assert x.dtype == y.dtype and x.shape == y.shape, (what, n)
        assert np.array_equal(x, y), "%s: %s differs" % (what, n)


CASES = [
    # name, N, B, kind, kwargs
    ("cfg1_gridify", 1024, 2, "surface",
     dict(max_p_grid=64, max_o_grid=1024, kernel_size=3, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("cfg1_knn", 1024, 2, "surface",
     dict(max_p_grid=32, max_o_grid=1024, kernel_size=5, loc=1, voxel_size=(0.05,) * 3, grid_size=(40,) * 3)),
    ("overflow_P_and_O", 2048, 1, "ball",
     dict(max_p_grid=8, max_o_grid=100, kernel_size=3, loc=1, voxel_size=(0.25,) * 3, grid_size=(8,) * 3)),
    ("loc0_aniso", 512, 2, "surface",
     dict(max_p_grid=16, max_o_grid=64, kernel_size=3, loc=0, voxel_size=(0.2, 0.25, 0.5), grid_size=(10, 8, 4))),
    ("kernel1", 300, 1, "ball",
     dict(max_p_grid=128, max_o_grid=1, kernel_size=1, loc=1, voxel_size=(2.0,) * 3, grid_size=(1,) * 3)),
]


@pytest.mark.parametrize("name,N,B,kind,kw", CASES, ids=[c[0] for c in CASES])
def test_gridify_oracle_vs_twin(oracle_mod, name, N, B, kind, kw):
    data, npts = synth.make_batch(B, N, seed0=11, kind=kind, voxels=(kw["voxel_size"][0],))
    npts[-1, 0] = N - 37  # ragged: the last cloud has fewer valid points
    kw = dict(kw, coord_shift=(1.0, 1.0, 1.0))
    _same(oracle_mod.gridify(data, npts, strict_reservoir=False, **kw), np_twin.gridify(data, npts, **kw), name + "/gridify")  # the twin implements the keep-first rule
    _same(oracle_mod.gridify_knn(data, npts, **kw), np_twin.gridify_knn(data, npts, **kw), name + "/knn")


def test_gridify_knn_fma_mode(oracle_mod):
    data, npts = synth.make_batch(1, 512, seed0=3)
    kw = dict(max_p_grid=16, max_o_grid=256, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.1,) * 3, grid_size=(20,) * 3)
    a = oracle_mod.gridify_knn(data, npts, dist_fma=True, **kw)
    b = np_twin.gridify_knn(data, npts, dist_fma=True, **kw)
    _same(a, b, "fma")


def test_duplicates_tie_break_by_index(oracle_mod):
    """Real loaders sample with replacement (ggcn_gpu_scannet_loader.py:239): duplicate points give
    exact distance ties, which the strict-< insertion resolves by arrival (= index) order."""
    data, npts = synth.make_batch(1, 256, seed0=5)
    data[0, 128:] = data[0, :128]
    kw = dict(max_p_grid=16, max_o_grid=256, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    a = oracle_mod.gridify_knn(data, npts, **kw)
    _same(a, np_twin.gridify_knn(data, npts, **kw), "dups")
    nebidx, nc = a[0][0], int(a[4][0, 0])
    for o in range(nc):  # a duplicate pair (i, i+128) must appear in index order when both present
        row = nebidx[o].tolist()
        for i in range(128):
            if i in row and i + 128 in row:
                assert row.index(i) < row.index(i + 128)


def test_empty_and_out_of_grid(oracle_mod):
    data, npts = synth.make_batch(2, 64, seed0=1)
    npts[0, 0] = 0                 # empty cloud
    data[1, :, :3] += 10.0         # every point outside the grid -> rejected (gridify.cu:136-138)
    kw = dict(max_p_grid=4, max_o_grid=8, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.5,) * 3, grid_size=(4,) * 3)
    for fn in (oracle_mod.gridify, oracle_mod.gridify_knn):
        nebidx, msk, cent, cmsk, num = fn(data, npts, **kw)
        assert not num.any() and not nebidx.any() and not msk.any() and not cmsk.any()
        assert np.all(cent == 1.0)  # init values, gridify-inl.h:117-121


def test_strict_reservoir_is_a_reservoir(oracle_mod):
    """Strict mode (host XORWOW, gridify.cu:259-270) keeps the multiset size and only ever selects
    ids that the raster walk visits; without overflow it equals keep-first."""
    data, npts = synth.make_batch(1, 2048, seed0=2, kind="ball", voxels=(0.25,))
    kw = dict(max_p_grid=16, max_o_grid=64, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    keep = oracle_mod.gridify(data, npts, strict_reservoir=False, **kw)
    strict = oracle_mod.gridify(data, npts, strict_reservoir=True, **kw)
    for k in (1, 3, 4):
        assert np.array_equal(keep[k], strict[k])
    assert not np.array_equal(keep[0], strict[0])  # overflow regime: the reservoir replaced ids
    big = dict(kw, max_p_grid=128)
    k2 = oracle_mod.gridify(data[:, :200], np.full((1, 1), 200, np.int32), strict_reservoir=False, **big)
    s2 = oracle_mod.gridify(data[:, :200], np.full((1, 1), 200, np.int32), strict_reservoir=True, **big)
    if (k2[1].sum(-1) < 128).all():
        _same(k2, s2, "no overflow")


def test_gridify_up_oracle_vs_twin(oracle_mod):
    down, dn = synth.make_batch(2, 256, seed0=21, voxels=(0.25,))
    up, un = synth.make_batch(2, 700, seed0=31, voxels=(0.25,))
    dn[1, 0], un[1, 0] = 200, 650
    up[0, 5, :3] = 5.0  # an up point outside the grid keeps the zero row
    for P, ks in ((5, 3), (40, 3), (3, 1), (16, 5)):
        kw = dict(max_p_grid=P, max_o_grid=700, kernel_size=ks, coord_shift=(1, 1, 1),
                  voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
        a = oracle_mod.gridify_up(down, up, dn, un, **kw)
        b = np_twin.gridify_up(down, up, dn, un, **kw)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (P, ks)
        assert not a[0][0, 5].any() and not a[1][0, 5].any()


def _lattice():
    pts = np.array([[i, j, k] for i in range(3) for j in range(3) for k in range(3)], np.float32)
    return pts


def test_knn_lattice_sanity_vector(oracle_mod):
    """utils/ops.py:282-299: query the centre of a 3x3x3 integer lattice with k=8.  Expected by hand:
    self (d2=0), the 6 face neighbours (d2=1) in index order, then the lowest-index edge neighbour
    (d2=2) because insertion uses strict < (k_nn-inl.h:75)."""
    pts = _lattice()
    known = pts[None]
    unknown = np.array([[[1.0, 1.0, 1.0]]], np.float32)
    one = np.array([[27]], np.int32), np.array([[1]], np.int32)
    idx = oracle_mod.knn(unknown, known, one[0], one[1], k=8)[0, 0].tolist()
    d2 = ((pts - 1.0) ** 2).sum(1)
    faces = [i for i in range(27) if d2[i] == 1]
    edge0 = min(i for i in range(27) if d2[i] == 2)
    assert idx == [13] + faces + [edge0]
    # 2x scaled copy (utils/ops.py:290-293) gives the same ids
    idx2 = oracle_mod.knn(unknown * 2, known * 2, one[0], one[1], k=8)[0, 0].tolist()
    assert idx2 == idx
    # BallKNN with radius 1 keeps self + faces only, misses are -1 (ball_k_nn-inl.h:69,77)
    ball = oracle_mod.ball_knn(unknown, known, one[0], one[1], k=6, radius=0.9)[0, 0].tolist()
    assert ball == [13, -1, -1, -1, -1, -1]
    ball = oracle_mod.ball_knn(unknown, known, one[0], one[1], k=6, radius=1.0)[0, 0].tolist()
    assert ball == [13] + faces[:5]


def test_knn_oracle_vs_twin(oracle_mod):
    rng = np.random.default_rng(0)
    unknown = rng.uniform(-1, 1, size=(2, 50, 3)).astype(np.float32)
    known = rng.uniform(-1, 1, size=(2, 40, 3)).astype(np.float32)
    known[0, 20:] = known[0, :20]  # exact ties
    downnum = np.array([[40], [33]], np.int32)
    upnum = np.array([[50], [41]], np.int32)
    for k in (1, 3, 5):
        a = oracle_mod.knn(unknown, known, downnum, upnum, k=k)
        b = np_twin.knn(unknown, known, downnum, upnum, k=k)
        assert np.array_equal(a, b), k
        a = oracle_mod.ball_knn(unknown, known, downnum, upnum, k=k, radius=0.6)
        b = np_twin.knn(unknown, known, downnum, upnum, k=k, radius=0.6)
        assert np.array_equal(a, b), k
    assert not oracle_mod.knn(unknown, known, downnum, upnum, k=3)[1, 41:].any()
    with pytest.raises(ValueError):
        oracle_mod.ball_knn(unknown, known, downnum, upnum, k=7, radius=0.5)  # best[6]


def test_oracle_threads_do_not_change_results(oracle_mod):
    data, npts = synth.make_batch(6, 512, seed0=40)
    kw = dict(max_p_grid=16, max_o_grid=128, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.1,) * 3, grid_size=(20,) * 3)
    oracle_mod.set_threads(1)
    a = oracle_mod.gridify_knn(data, npts, **kw)
    oracle_mod.set_threads(4)
    b = oracle_mod.gridify_knn(data, npts, **kw)
    oracle_mod.set_threads(1)
    _same(a, b, "threads")


def test_golden_fixtures(oracle_mod):
    """tests/golden/*.npz were written by tests/golden/make_golden.py (from this oracle: the
    reference cannot run here); they pin the oracle against silent drift."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz"))
    assert files, "no golden fixtures committed"
    from tests.golden import make_golden
    for f in files:
        z = np.load(os.path.join(GOLDEN, f), allow_pickle=False)
        got = make_golden.run_case(oracle_mod, str(z["case"]))
        for k, v in got.items():
            if f.startswith("gridconv") and k == "out":  # BLAS summation order may differ per host
                assert np.allclose(z[k], v, rtol=1e-5, atol=1e-6), (f, k)
            else:
                assert np.array_equal(z[k], v), (f, k)


# ------------------------------------------------------------------------------------------------
# Coverage-Aware Sampling (Gridify_occaware): restated from the SASS of additional.so, parity unpinned
# ------------------------------------------------------------------------------------------------
CAS_KW = dict(max_p_grid=16, max_o_grid=256, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.05,) * 3, grid_size=(40,) * 3)


def _covered_fraction(data, cent, centnum, b, voxel=0.05):
    vox = np.floor((data[b, :, :3] + np.float32(1)) / np.float32(voxel)).astype(int)
    occ = set(map(tuple, vox))
    cv = np.floor((cent[b, :centnum[b, 0], :3] + np.float32(1)) / np.float32(voxel)).astype(int)
    cov = {(c[0] + dx, c[1] + dy, c[2] + dz) for c in cv for dx in (-1, 0, 1) for dy in (-1, 0, 1)
           for dz in (-1, 0, 1)}
    return len(occ & cov) / len(occ)


def test_occaware_c_oracle_matches_numpy_twin(oracle_mod):
    from gridgcn_b200 import synth
    from oracle import np_twin
    data, npts = synth.make_batch(2, 1500, seed0=3)
    npts[1, 0] = 1333
    for seed in (0, 77, 2 ** 40 + 5):
        a = oracle_mod.gridify_occaware(data, npts, seed=seed, **CAS_KW)
        b = np_twin.gridify_occaware(data, npts, seed=seed, **{k: v for k, v in CAS_KW.items()})
        for x, y, n in zip(a, b, ("nebidx", "nebidxmsk", "cent", "centmsk", "actual_centnum")):
            assert np.array_equal(x, y), (seed, n)
    # boundary voxels (neighbourhoods clipped by the grid) and an anisotropic grid
    kw = dict(max_p_grid=8, max_o_grid=20, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
              voxel_size=(0.25, 0.25, 0.5), grid_size=(8, 8, 4))
    data, npts = synth.make_batch(2, 600, seed0=9, kind="ball", voxels=(0.25,))
    a = oracle_mod.gridify_occaware(data, npts, seed=1, **kw)
    b = np_twin.gridify_occaware(data, npts, seed=1, **kw)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_occaware_properties(oracle_mod):
    from gridgcn_b200 import synth
    data, npts = synth.make_batch(3, 4096, seed0=11)
    kw = dict(CAS_KW, max_o_grid=512)
    cas = oracle_mod.gridify_occaware(data, npts, seed=5, **kw)
    rvs = oracle_mod.gridify(data, npts, **kw)
    assert np.array_equal(cas[4], rvs[4]) and np.array_equal(cas[3], rvs[3])  # count / mask unchanged
    for b in range(3):
        # the selected centres are distinct occupied voxels ...
        cv = np.floor((cas[2][b, :, :3] + np.float32(1)) / np.float32(0.05)).astype(int)
        assert len(set(map(tuple, cv))) == 512
        # ... and cover (much) more of the occupied space than keep-first sampling: the point of CAS
        assert _covered_fraction(data, cas[2], cas[4], b) > _covered_fraction(data, rvs[2], rvs[4], b) + 0.1
        assert _covered_fraction(data, cas[2], cas[4], b) > 0.9
    # different seeds give different (but equally valid) selections
    other = oracle_mod.gridify_occaware(data, npts, seed=6, **kw)
    assert not np.array_equal(other[2], cas[2])
    # no challengers (occupied voxels <= max_o): identical to Gridify
    few, nf = synth.make_batch(2, 200, seed0=2)
    a = oracle_mod.gridify_occaware(few, nf, seed=5, **kw)
    b = oracle_mod.gridify(few, nf, **kw)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    # batch independence: a cloud's result does not depend on its position in the batch
    solo = oracle_mod.gridify_occaware(data[2:3], npts[2:3], seed=5, **kw)
    for x, y in zip(solo, cas):
        assert np.array_equal(x[0], y[2])


def test_occaware_golden(oracle_mod):
    z = np.load(os.path.join(GOLDEN, "gridify_occaware.npz"))
    got = oracle_mod.gridify_occaware(z["data"], z["npts"], seed=2026, **CAS_KW)
    for g, n in zip(got, ("nebidx", "nebidxmsk", "cent", "centmsk", "actual_centnum")):
        assert np.array_equal(g, z[n]), n


# ------------------------------------------------------------------------------------------------
# Third-party-algorithm pin of the GridConv oracle's building blocks.  MXNet 1.5 cannot be installed here,
# so Convolution(1x1) / BatchNorm(eval, eps 1e-3, fix_gamma False) / Pooling(max) / take(mode clip) -- the MXNet
# operators utils/ops.py:78-93,149-158 and gcn_module_g_att.py:57-59 call -- are checked against another
# vendor's LIBRARY implementation of the same published operators (torch.nn.functional), not against a second
# hand-written restatement.
# ------------------------------------------------------------------------------------------------
def test_gridconv_oracle_blocks_against_torch_library_ops():
    import torch
    import torch.nn.functional as TF
    from oracle import gridconv_oracle as go
    from gridgcn_b200 import gridconv
    rng = np.random.default_rng(12)
    B, Cin, Cout, O, P = 3, 10, 16, 7, 5
    st = gridconv.init_stage(rng, Cin, Cout)
    x = rng.normal(size=(B, Cin, O, P)).astype(np.float32)
    got = go._conv_bn_relu(x, st)
    w = torch.from_numpy(np.asarray(st["weight"], np.float32).reshape(Cout, Cin, 1, 1))
    y = TF.conv2d(torch.from_numpy(x), w, torch.from_numpy(st["bias"]))
    y = TF.batch_norm(y, torch.from_numpy(st["moving_mean"]), torch.from_numpy(st["moving_var"]),
                      torch.from_numpy(st["gamma"]), torch.from_numpy(st["beta"]), training=False, eps=1e-3)
    want = TF.relu(y).numpy()
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    # the folded form the kernels consume (gridconv.fold_bn) is the same function
    fw, fb = gridconv.fold_bn(st["weight"], st["bias"], st["gamma"], st["beta"], st["moving_mean"], st["moving_var"])
    folded = np.maximum(np.einsum("oc,bcnp->bonp", fw, x) + fb[None, :, None, None], 0)
    assert np.allclose(folded, want, rtol=1e-5, atol=1e-5)
    # 1-D stage of the decoder / head (mlp1d_c, utils/ops.py:141-147,236-242)
    st1 = gridconv.init_stage(rng, Cin, Cout)
    x1 = rng.normal(size=(B, O, Cin)).astype(np.float32)
    y1 = TF.conv1d(torch.from_numpy(x1).transpose(1, 2), torch.from_numpy(np.asarray(st1["weight"], np.float32).reshape(Cout, Cin, 1)),
                   torch.from_numpy(st1["bias"]))
    y1 = TF.relu(TF.batch_norm(y1, torch.from_numpy(st1["moving_mean"]), torch.from_numpy(st1["moving_var"]),
                               torch.from_numpy(st1["gamma"]), torch.from_numpy(st1["beta"]), training=False, eps=1e-3))
    assert np.allclose(go._conv1d_bn_relu(x1.transpose(0, 2, 1), st1), y1.numpy(), rtol=1e-5, atol=1e-6)  # (B, C, O) layout
    # max pooling over the P slots (Pooling kernel (1, P), :57-59) and take(mode="clip") after the batch offset
    pooled = TF.max_pool2d(torch.from_numpy(got), kernel_size=(1, P)).numpy()[..., 0]
    assert np.array_equal(pooled, got.max(axis=3))
    data = rng.normal(size=(B, 11, 6)).astype(np.float32)
    idx = rng.integers(-1, 11, size=(B, O, P)).astype(np.int32)  # -1 = BallKNN miss
    flat = torch.from_numpy(data.reshape(B * 11, 6))
    gi = torch.from_numpy(idx.astype(np.int64)) + (torch.arange(B) * 11)[:, None, None]
    want_take = flat[gi.clamp(0, B * 11 - 1)].numpy()
    assert np.array_equal(go.batch_take_g(data, idx), want_take)
    # one whole layer: the oracle against a torch-library composition of the same graph
    layer = gridconv.init_layer(np.random.default_rng(3), 6, [8, 12], 10)
    table = rng.uniform(-1, 1, size=(B, 11, 4 + 6)).astype(np.float32)
    nidx = rng.integers(0, 11, size=(B, O, P)).astype(np.int32)
    cent = rng.uniform(-1, 1, size=(B, O, 4)).astype(np.float32)
    msk = (rng.uniform(size=(B, O)) > 0.3).astype(np.float32)
    out = go.gridconv_layer(table, nidx, cent, msk, layer)

    def cbr(t, stg):
        wt = torch.from_numpy(np.asarray(stg["weight"], np.float32).reshape(len(stg["bias"]), -1, 1, 1))
        t = TF.conv2d(t, wt, torch.from_numpy(stg["bias"]))
        return TF.relu(TF.batch_norm(t, torch.from_numpy(stg["moving_mean"]), torch.from_numpy(stg["moving_var"]),
                                     torch.from_numpy(stg["gamma"]), torch.from_numpy(stg["beta"]), training=False, eps=1e-3))
    nb = torch.from_numpy(go.batch_take_g(table, nidx)).permute(0, 3, 1, 2)           # (B, 4+C, O, P)
    cxyz = torch.from_numpy(cent[:, :, :3]).permute(0, 2, 1)[:, :, :, None].expand(B, 3, O, P)
    geo = nb[:, :3] - cxyz
    dist = geo.pow(2).sum(1, keepdim=True).sqrt()
    att = torch.cat([dist, geo, cxyz, nb[:, :3]], dim=1)                              # attfdim 10 (:209-222)
    f = nb[:, 4:]
    for stg in layer["feat"]:
        f = cbr(f, stg)
    a = att
    for stg in layer["att"]:
        a = cbr(a, stg)
    pooled = TF.max_pool2d(f * a, kernel_size=(1, P))[..., 0]                          # (B, C, O)
    feats = (TF.relu(pooled) * torch.from_numpy(msk)[:, None, :]).permute(0, 2, 1).numpy()
    assert np.array_equal(out[..., :4], cent)
    assert np.allclose(out[..., 4:], feats, rtol=1e-4, atol=1e-5)
