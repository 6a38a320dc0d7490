"""CPU suite, part 3: host-side logic of the product (BN folding, parameter naming, ladders,
synthetic inputs) -- no GPU needed."""
import os

import numpy as np

from gridgcn_b200 import gridconv, stack, synth


def test_fold_bn_matches_conv_bn():
    from oracle import gridconv_oracle
    rng = np.random.default_rng(0)
    st = gridconv.init_stage(rng, 10, 16)
    x = rng.normal(size=(2, 10, 3, 5)).astype(np.float32)
    ref = gridconv_oracle._conv_bn_relu(x, st)
    w, b = gridconv.fold_bn(st["weight"], st["bias"], st["gamma"], st["beta"], st["moving_mean"],
                            st["moving_var"])
    got = np.maximum(np.einsum("oc,bcnp->bonp", w, x) + b[None, :, None, None], 0)
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_named_params_follow_reference_scopes():
    layer = gridconv.init_layer(np.random.default_rng(0), 0, [32, 32, 64], 10)
    names = gridconv.named_params(layer, "sub_g_0")
    assert "sub_g_0/conv1_weight" in names and names["sub_g_0/conv1_weight"].shape == (32, 3)
    assert names["sub_g_0/conv3/bn_moving_var"].shape == (64,)
    assert names["sub_g_0/update_att_mlp2d_frst/conv1_weight"].shape == (16, 10)
    assert names["sub_g_0/update_att_mlp2d_scnd/conv1_weight"].shape == (64, 16)


def test_ladders_match_reference_configs():
    s = stack.seg8192_shipped()  # segmentation/configs/configs.yaml:72-77
    assert [l.max_o_grid for l in s.layers] == [1024, 256, 24]
    assert [l.max_p_grid for l in s.layers] == [64, 32, 32]
    assert [l.grid_size for l in s.layers] == [40, 15, 5]
    assert stack.seg81920_shipped().layers[0].max_p_grid == 128  # configs.yaml:151
    s4 = stack.seg8192_4layer(64)
    assert [l.max_o_grid for l in s4.layers] == [1024, 256, 64, 16]
    assert [list(l.pt_mlp_lst) for l in s4.layers][3] == [256, 256, 512]
    params = stack.init_params(s4)
    assert [p["cin"] for p in params] == [0, 64, 128, 256]
    # per-edge MAC counts of SURVEY.md s8a
    macs = []
    for p in params:
        m = sum(st["weight"].size for st in p["feat"]) + sum(st["weight"].size for st in p["att"])
        macs.append(m)
    assert macs == [4352, 20800, 82560, 328960]


def test_synthetic_clouds():
    data, npts = synth.make_batch(3, 1024, seed0=0)
    assert data.shape == (3, 1024, 4) and data.dtype == np.float32 and npts.tolist() == [[1024]] * 3
    assert np.all(data[..., 3] == 1.0)
    assert np.abs(data[..., :3]).max() <= 1.0 + 1e-6
    for b in range(3):
        assert len(np.unique(data[b, :, :3], axis=0)) == 1024  # no duplicates
    q = (data[..., :3] + np.float32(1.0)) / np.float32(0.05)
    frac = q - np.floor(q)
    assert frac.min() > 5e-5 and frac.max() < 1 - 5e-5  # away from voxel faces
    again, _ = synth.make_batch(3, 1024, seed0=0)
    assert np.array_equal(data, again)


def test_reference_yaml_keys_are_understood():
    import os
    import pytest
    here = os.path.dirname(os.path.abspath(__file__))
    cfg, up = stack.load_reference_yaml(os.path.join(here, "golden", "seg8192_reference_keys.yaml"))
    ship = stack.seg8192_shipped()
    assert [(l.voxel_size, l.grid_size, l.max_o_grid, l.max_p_grid, l.kernel_size, list(l.pt_mlp_lst))
            for l in cfg.layers] == \
           [(l.voxel_size, l.grid_size, l.max_o_grid, l.max_p_grid, l.kernel_size, list(l.pt_mlp_lst))
            for l in ship.layers]
    assert cfg.num_points == 8192 and cfg.loc == 1 and cfg.attfdim == 10 and cfg.pre_relu
    assert (up.max_p_grid, up.kernel_size, up.neigh_fetch) == (5, 3, "ballknn")
    assert tuple(up.pt_mlp_lst) == (128,) and tuple(up.center_dim) == (128,) and tuple(up.out_dim) == (128,)
    ref = "/root/reference/segmentation/configs/configs.yaml"
    if os.path.exists(ref):  # the reference's own file (build container only)
        rcfg, rup = stack.load_reference_yaml(ref)
        assert [l.max_o_grid for l in rcfg.layers] == [1024, 256, 24] and rup.max_p_grid == 5
    with pytest.raises(NotImplementedError):
        stack.from_reference_config(dict(max_o_grid_lst=[8], voxel_size_lst=[[0.1, 0.2, 0.1]],
                                         grid_size_lst=[[4, 4, 4]], max_p_grid_lst=[4], kernel_size_lst=[3],
                                         pt_ele_dim=[[8]], num_points=16))
    # the decoder ladder is READ, not assumed: radius / GridifyUp grid come from the up_* lists
    assert tuple(up.voxel_size_lst) == (0.4, 0.133333, 0.05) and tuple(up.grid_size_lst) == (5, 15, 40)
    assert tuple(up.max_o_grid_lst) == (256, 1024, 8192)
    import yaml
    with open(os.path.join(here, "golden", "seg8192_reference_keys.yaml")) as f:
        conf = yaml.safe_load(f)
    knn_cfg, knn_up = stack.from_reference_config(dict(conf, real_knn=True))
    assert knn_up.neigh_fetch == "knn"
    assert stack.from_reference_config(dict(conf, up_neigh_fetch=False))[1].neigh_fetch == "gridifyup"
    for bad in (dict(up_max_o_grid_lst=[256, 1024, 4096]), dict(up_att_full="next"), dict(elevation=[1]),
                dict(up_cntxt_mlp_lst=[[8], [8], [8]]), dict(use_bn="f"), dict(up_center_inte="add"),
                dict(aggtype="agg_gcn"), dict(gcn_outDim=[[64], [], []])):
        with pytest.raises(NotImplementedError):  # unsupported keys are refused, never silently ignored
            stack.from_reference_config(dict(conf, **bad))
    # the classification keys (classification/configs/configs.yaml:44-68)
    cls_conf = dict(num_points=1024, voxel_size_lst=[[0.05] * 3, [0.25] * 3, [2.0] * 3],
                    grid_size_lst=[[40] * 3, [8] * 3, [1] * 3], lidar_coord=[1.0, 1.0, 1.0],
                    max_p_grid_lst=[64, 64, 128], max_o_grid_lst=[1024, 128, 1], kernel_size_lst=[7, 3, 1],
                    stride_lst=[1, 1, 1], aggtype="gcn", localfdim=3, attfdim=4, elevation=[],
                    pt_ele_dim=[[64, 64, 128], [128, 128, 256], [256, 256, 512]],
                    att_ele_dim=[[64, 128, 128], [128, 256, 256], [256, 512, 512]], cntxt_mlp_lst=[[], [], []],
                    gcn_outDim=[[], [], []], relu=True, agg="max_pooling", att_full="next", use_bn="t",
                    group_all=False, loc_within=True)
    ccfg, cup = stack.from_reference_config(cls_conf)
    ship = stack.cls1024_shipped()
    assert cup is None and (ccfg.att_full, ccfg.localfdim, ccfg.attfdim) == ("next", 3, 4)
    assert [(l.voxel_size, l.grid_size, l.max_o_grid, l.max_p_grid, l.kernel_size, list(l.pt_mlp_lst),
             list(l.att_ele_lst)) for l in ccfg.layers] == \
           [(l.voxel_size, l.grid_size, l.max_o_grid, l.max_p_grid, l.kernel_size, list(l.pt_mlp_lst),
             list(l.att_ele_lst)) for l in ship.layers]
    cref = "/root/reference/classification/configs/configs.yaml"
    if os.path.exists(cref):
        with open(cref) as f:
            rc = yaml.safe_load(f)
        rc.setdefault("num_points", 1024)
        assert [l.max_o_grid for l in stack.from_reference_config(rc)[0].layers] == [1024, 128, 1]
    head = stack.init_cls_params(ccfg, seed=0)["head"]
    assert [st["weight"].shape[0] for st in head] == [512, 256, 40] and head[0]["weight"].reshape(512, -1).shape[1] == 512


def test_cas_state_table_is_current():
    """grid-gcn_b200/csrc/cas_h_table.inc is generated: regenerate and compare; spot-check that walking
    the table reproduces the float/double arithmetic it stands for and that ids order like values."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_cas_table", os.path.join(root, "tools", "gen_cas_table.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    with open(gen.PATH) as f:
        assert f.read() == gen.render()
    values, nxt_a, nxt_ac = gen.build()
    assert values[0] == 0 and all(a < b for a, b in zip(values, values[1:]))
    rng = np.random.default_rng(0)
    for _ in range(200):
        s, h = 0, np.float32(0)
        for occ in rng.integers(0, 2, size=27):
            s = (nxt_ac if occ else nxt_a)[s]
            h = np.float32(np.float64(h) + 0.7)
            if occ:
                h = np.float32(np.float64(h) + 0.3)
            assert values[s] == h


def test_stack_query_selection():
    class Fake:
        Gridify, GridifyKNN = "g", "k"

        @staticmethod
        def Gridify_occaware(*a, **kw):
            return kw
    assert stack.query_fn(stack.seg8192_shipped(), Fake) == "g"
    assert stack.query_fn(stack.seg8192_4layer(64), Fake) == "k"
    cfg = stack.seg8192_4layer(64, "occaware_knn")
    cfg.cas_seed = 5
    assert stack.query_fn(cfg, Fake)(max_o_grid=1) == dict(seed=5, knn_query=True, max_o_grid=1)
    cfg.query = "nonsense"
    try:
        stack.query_fn(cfg, Fake)
        assert False
    except ValueError:
        pass
