"""The GridConv encoder stack: per layer  Gridify|GridifyKNN -> fused GridConv.

Mirrors the encoder loop of get_symbol_seg_ggcn (reference
segmentation/models/ggcn_models_g.py:152-187): each layer voxelises the previous layer's centres
(``data_loc = centers``, :160), chains ``actual_centnum`` (:166), gathers from the previous layer's
table ``concat(centers, center_feats)`` (:172,:186) and runs sub_g_update (:185).  The ladders below
are the reference's shipped configs (segmentation/configs/configs.yaml:72-77, :148-153;
classification/configs/configs.yaml:47-52) and the 4-layer layout its comments describe
(configs.yaml:45-48, segmentation/train_test/command:6).
"""
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np
import torch

from . import gridconv, ops


@dataclass
class LayerCfg:
    voxel_size: float
    grid_size: int
    max_o_grid: int
    max_p_grid: int
    kernel_size: int
    pt_mlp_lst: Sequence[int]
    stride: int = 1
    att_ele_lst: Sequence[int] = ()  # explicit attention widths (classification block); () = [C/4, C]


@dataclass
class StackCfg:
    name: str
    num_points: int
    layers: List[LayerCfg]
    coord_shift: Sequence[float] = (1.0, 1.0, 1.0)  # lidar_coord
    loc: int = 1                                     # loc_within: True
    attfdim: int = 10
    pre_relu: bool = True                            # configs["relu"]
    query: str = "gridifyknn"                        # "gridify" (what the seg graph calls), "gridifyknn",
                                                     # "occaware" / "occaware_knn" (coverage-aware sampling)
    cas_seed: int = 0                                # seed of the coverage-aware sampling
    att_full: str = ""                               # "next" / "last": classification block (fp32 precision)
    localfdim: int = 0                               # 3: geo vector in front of the gathered features
    voxels: Sequence[float] = field(default_factory=tuple)

    def __post_init__(self):
        self.voxels = tuple(l.voxel_size for l in self.layers)


_MLP4 = ([32, 32, 64], [64, 64, 128], [128, 128, 256], [256, 256, 512])


def seg8192_4layer(K=64, query="gridifyknn"):
    """N=8192; centres 1024/256/64/16 with K neighbours each and the MLP widths of
    configs.yaml:45-48.  Grids 40^3, 15^3 are the shipped ones; 8^3, 4^3 continue the ladder
    (SURVEY.md s8d config 3)."""
    vox, grid, O = [0.05, 0.133333, 0.25, 0.5], [40, 15, 8, 4], [1024, 256, 64, 16]
    return StackCfg("seg8192_4layer_K%d" % K, 8192,
                    [LayerCfg(vox[i], grid[i], O[i], K, 3, _MLP4[i]) for i in range(4)], query=query)


def cls1024_4layer(K=32, query="gridifyknn"):
    """ModelNet40-shaped input (N=1024) through a 4-layer ladder."""
    vox, grid, O = [0.05, 0.1, 0.25, 0.5], [40, 20, 8, 4], [512, 128, 32, 8]
    return StackCfg("cls1024_4layer_K%d" % K, 1024,
                    [LayerCfg(vox[i], grid[i], O[i], K, 3, _MLP4[i]) for i in range(4)], query=query)


def seg8192_shipped(query="gridify"):
    """segmentation/configs/configs.yaml:72-77 exactly (3 encoder layers)."""
    vox, grid, O, P = [0.05, 0.133333, 0.4], [40, 15, 5], [1024, 256, 24], [64, 32, 32]
    mlps = ([32, 32, 64], [64, 64, 128], [128, 128, 256])
    return StackCfg("seg8192_shipped", 8192,
                    [LayerCfg(vox[i], grid[i], O[i], P[i], 3, mlps[i]) for i in range(3)], query=query)


def seg81920_shipped(query="gridify"):
    """segmentation/configs/configs.yaml:148-153: as seg8192 but N=81920 and P0=128."""
    cfg = seg8192_shipped(query)
    cfg.name, cfg.num_points = "seg81920_shipped", 81920
    cfg.layers[0].max_p_grid = 128
    return cfg


def cls1024_shipped(query="gridify"):
    """classification/configs/configs.yaml:44-68 exactly: 3 layers, kernel 7/3/1, the last layer one voxel
    holding everything; attfdim 4, localfdim 3, att_full next, explicit attention widths -- the
    classification flavour of the block (classification/models/gcn_module_g.py), fp32 precision."""
    vox, grid, O, P, ks = [0.05, 0.25, 2.0], [40, 8, 1], [1024, 128, 1], [64, 64, 128], [7, 3, 1]
    pt = ([64, 64, 128], [128, 128, 256], [256, 256, 512])
    att = ([64, 128, 128], [128, 256, 256], [256, 512, 512])
    return StackCfg("cls1024_shipped", 1024,
                    [LayerCfg(vox[i], grid[i], O[i], P[i], ks[i], pt[i], att_ele_lst=att[i]) for i in range(3)],
                    attfdim=4, query=query, att_full="next", localfdim=3)


def tiny(K=8, query="gridifyknn"):
    """Small ladder for smoke tests."""
    return StackCfg("tiny_K%d" % K, 256,
                    [LayerCfg(0.25, 8, 64, K, 3, [16, 32]), LayerCfg(0.5, 4, 16, K, 3, [32, 64])],
                    query=query)


def init_params(cfg: StackCfg, seed=0):
    """Random parameters for every layer (there are no checkpoints in the container)."""
    rng = np.random.default_rng(seed)
    layers, cin = [], 0
    for l in cfg.layers:
        layers.append(gridconv.init_layer(rng, cin, list(l.pt_mlp_lst), cfg.attfdim,
                                          att_ele_lst=list(l.att_ele_lst) or None, att_full=cfg.att_full,
                                          localfdim=cfg.localfdim))
        cin = l.pt_mlp_lst[-1]
    return layers


QUERIES = ("gridify", "gridifyknn", "occaware", "occaware_knn")


def query_fn(cfg, module=ops):
    """The centre-sampling + neighbour-query operator a configuration names, bound to its extra
    arguments.  ``module`` is anything with Gridify / GridifyKNN / Gridify_occaware (the product ops, or
    an oracle front-end with snake_case names through ``oracle_query_fn`` in the tests)."""
    if cfg.query not in QUERIES:
        raise ValueError("query must be one of %s" % (QUERIES,))
    if cfg.query == "gridify":
        return module.Gridify
    if cfg.query == "gridifyknn":
        return module.GridifyKNN
    knn = cfg.query == "occaware_knn"
    return lambda *a, **kw: module.Gridify_occaware(*a, seed=cfg.cas_seed, knn_query=knn, **kw)


class GridGcnEncoder:
    """``feats_table = enc(data, actual_numpoints)``: data (B,N,4) f32 cuda, actual_numpoints
    (B,1) i32 -> last layer's table (B, O_last, 4+C_last).  ``enc.trace`` keeps every layer's
    outputs of the last call for parity checks."""

    def __init__(self, cfg: StackCfg, params, device, precision="fp32"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.convs = [gridconv.GridConv(p, device, pre_relu=cfg.pre_relu, precision=precision)
                      for p in params]
        self.trace = []

    def __call__(self, data, actual_numpoints, keep_trace=False):
        cfg = self.cfg
        query = query_fn(cfg)
        table, data_loc, num = data, data, actual_numpoints
        trace = []
        for l, conv in zip(cfg.layers, self.convs):
            nebidx, nebmsk, cent, centmsk, num = query(
                data_loc, num, max_o_grid=l.max_o_grid, max_p_grid=l.max_p_grid,
                kernel_size=l.kernel_size, stride=l.stride, coord_shift=cfg.coord_shift,
                voxel_size=[l.voxel_size] * 3, grid_size=[l.grid_size] * 3, loc=cfg.loc)
            table = conv(table, nebidx, cent, centmsk)
            data_loc = cent
            if keep_trace:
                trace.append(dict(nebidx=nebidx, nebidxmsk=nebmsk, cent=cent, centmsk=centmsk,
                                  actual_centnum=num, table=table))
        self.trace = trace
        return table


def capture_graph(module, data, actual_numpoints, warmup=2):
    """Captures ``module(data, actual_numpoints)`` (a GridGcnEncoder or GridGcnSeg forward: every kernel of
    every layer, and the workspace / output allocations, which then come from the graph's private pool)
    into ONE CUDA graph for these shapes.  Returns ``replay(data=None, actual_numpoints=None) -> out``:
    new inputs are copied into the captured input buffers, the graph is launched once, and the captured
    output tensor is returned (valid until the next replay).  Worth it when a step is launch bound
    (small batches: ~15 kernel launches and as many allocations per forward)."""
    static_data, static_num = data.clone(), actual_numpoints.clone()
    side = torch.cuda.Stream(device=data.device)
    side.wait_stream(torch.cuda.current_stream(data.device))
    with torch.cuda.stream(side):
        for _ in range(max(1, warmup)):  # first calls set kernel attributes: keep them out of the capture
            module(static_data, static_num)
    torch.cuda.current_stream(data.device).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = module(static_data, static_num)

    def replay(data=None, actual_numpoints=None):
        if data is not None:
            static_data.copy_(data, non_blocking=True)
        if actual_numpoints is not None:
            static_num.copy_(actual_numpoints, non_blocking=True)
        graph.replay()
        return out

    replay.graph = graph
    return replay


# ---------------------------------------------------------------------------------------------------
# Encoder + decoder + head of the segmentation graph (get_symbol_seg_ggcn, ggcn_models_g.py:110-237)
# ---------------------------------------------------------------------------------------------------
@dataclass
class UpCfg:
    max_p_grid: int = 5                       # up_max_p_grid_lst
    kernel_size: int = 3                      # up_kernel_size_lst
    pt_mlp_lst: Sequence[int] = (128,)        # up_pt_ele_dim
    center_dim: Sequence[int] = (128,)        # up_center_dim
    out_dim: Sequence[int] = (128,)           # up_gcn_outDim
    attfdim: int = 10                         # up_attfdim
    neigh_fetch: str = "ballknn"              # up_neigh_fetch: True -> BallKNN ("knn" when real_knn is set:
                                              # contrib.KNN, ggcn_models_g.py:74-83); False -> "gridifyup"
    # per decoder level (ggcn_models_g.py:204-210 index these with the decoder step i); () = the encoder ladder
    # reversed, which is what the shipped configs contain (configs.yaml:99-102)
    voxel_size_lst: Sequence[float] = ()      # up_voxel_size_lst (BallKNN radius = voxel * kernel * 1.7 / 2)
    grid_size_lst: Sequence[int] = ()         # up_grid_size_lst  (GridifyUp)
    max_o_grid_lst: Sequence[int] = ()        # up_max_o_grid_lst (GridifyUp; must equal the up level's row count)


def init_seg_params(cfg: StackCfg, up: UpCfg, seed=0, num_classes=21):
    """Encoder layers, one decoder layer per encoder layer (reverse order) and the head."""
    rng = np.random.default_rng(seed)
    enc, cin = [], 0
    for l in cfg.layers:
        enc.append(gridconv.init_layer(rng, cin, list(l.pt_mlp_lst), cfg.attfdim))
        cin = l.pt_mlp_lst[-1]
    widths = [0] + [l.pt_mlp_lst[-1] for l in cfg.layers]  # feature width of every level (input level: 0)
    dec, cd = [], widths[-1]
    for i in range(len(cfg.layers)):
        cu = widths[-i - 2]
        dec.append(gridconv.init_up_layer(rng, cd, cu, list(up.pt_mlp_lst), up.attfdim, list(up.center_dim),
                                          list(up.out_dim)))
        cd = up.out_dim[-1]
    head = [gridconv.init_stage(rng, cd, 128), gridconv.init_stage(rng, 128, num_classes)]
    return dict(enc=enc, dec=dec, head=head)


class GridGcnSeg:
    """``logits = net(data, actual_numpoints)``: (B,N,4) -> (B,N,21).  Mirrors get_symbol_seg_ggcn: the
    encoder loop (:152-187), then for every level in reverse the neighbour fetch
    (BallKNN with radius = up_voxel * up_kernel * 1.7 / 2, :201-204, or GridifyUp, :206-210), the gather
    from the coarser level's [cent | feat] table and sub_g_update with center_ori_feats (:212-231), and
    the head (:233-236)."""

    def __init__(self, cfg: StackCfg, up: UpCfg, params, device, precision="tf32x3"):
        self.cfg, self.up = cfg, up
        self.enc = GridGcnEncoder(cfg, params["enc"], device, precision=precision)
        self.dec = [gridconv.GridConvUp(p, device, pre_relu=cfg.pre_relu, precision=precision)
                    for p in params["dec"]]
        self.head = gridconv.SegHead(params["head"], device, precision=precision)
        self.trace = []

    def __call__(self, data, actual_numpoints, keep_trace=False):
        cfg, up = self.cfg, self.up
        self.enc(data, actual_numpoints, keep_trace=True)
        et = self.enc.trace
        # level 0 = input points, level i+1 = encoder layer i
        tables = [data] + [t["table"] for t in et]
        cents = [data] + [t["cent"] for t in et]
        nums = [actual_numpoints] + [t["actual_centnum"] for t in et]
        masks = [None] + [t["centmsk"] for t in et]
        f_last, trace = tables[-1], []
        nl = len(cfg.layers)
        for i, dec in enumerate(self.dec):
            lvl_dn, lvl_up = nl - i, nl - i - 1
            down, centers = cents[lvl_dn], cents[lvl_up]
            l = cfg.layers[lvl_up]  # default ladder: up_voxel_size_lst is the encoder ladder reversed
            voxel = float(up.voxel_size_lst[i]) if len(up.voxel_size_lst) else l.voxel_size
            grid = int(up.grid_size_lst[i]) if len(up.grid_size_lst) else l.grid_size
            if len(up.max_o_grid_lst) and int(up.max_o_grid_lst[i]) != centers.shape[1]:
                raise ValueError("up_max_o_grid_lst[%d] = %d but the up level has %d rows"
                                 % (i, up.max_o_grid_lst[i], centers.shape[1]))
            if up.neigh_fetch == "ballknn":
                radius = voxel * up.kernel_size * 1.7 / 2  # ggcn_models_g.py:204
                nebidx = ops.contrib.BallKNN(centers[:, :, :3].contiguous(), down[:, :, :3].contiguous(),
                                             nums[lvl_dn], nums[lvl_up], k=up.max_p_grid, radius=radius)
            elif up.neigh_fetch == "knn":  # real_knn (:74-83); ThreeNN (k == 3) is the same exact search
                nebidx = ops.contrib.KNN(centers[:, :, :3].contiguous(), down[:, :, :3].contiguous(),
                                         nums[lvl_dn], nums[lvl_up], k=up.max_p_grid)
            elif up.neigh_fetch == "gridifyup":
                nebidx, _ = ops.GridifyUp(down, centers, nums[lvl_dn], nums[lvl_up],
                                          max_p_grid=up.max_p_grid, max_o_grid=centers.shape[1],
                                          kernel_size=up.kernel_size, coord_shift=cfg.coord_shift,
                                          voxel_size=[voxel] * 3, grid_size=[grid] * 3)
            else:
                raise ValueError("UpCfg.neigh_fetch must be ballknn, knn or gridifyup, not %r" % (up.neigh_fetch,))
            mask = masks[lvl_up] if i != len(self.dec) - 1 else None  # ggcn_models_g.py:224
            f_last = dec(f_last, nebidx, centers, tables[lvl_up], mask)
            if keep_trace:
                trace.append(dict(nebidx=nebidx, table=f_last))
        logits = self.head(f_last[:, :, 4:])
        self.trace = trace
        return logits


# ---------------------------------------------------------------------------------------------------
# Classification graph (get_symbol_cls_ggcn, classification/models/ggcn_models_g.py:37-111) + head (:25-35)
# ---------------------------------------------------------------------------------------------------
def init_cls_params(cfg: StackCfg, seed=0, num_classes=40):
    """Encoder layers of ``cfg`` (the classification flavour when cfg.localfdim / att_full say so) and the
    FC 512 - 256 - num_classes head on the flattened features of the last layer."""
    rng = np.random.default_rng(seed)
    enc = init_params(cfg, seed)
    width = cfg.layers[-1].pt_mlp_lst[-1] * cfg.layers[-1].max_o_grid  # fully_connected(flatten=True)
    head = [gridconv.init_stage(rng, width, 512), gridconv.init_stage(rng, 512, 256),
            gridconv.init_stage(rng, 256, num_classes)]
    return dict(enc=enc, head=head)


class GridGcnCls:
    """``scores = net(data, actual_numpoints)``: (B,N,4) -> (B,num_classes).  The encoder loop of
    get_symbol_cls_ggcn (Gridify -> batch_take_g -> sub_g_update per layer, :66-106; group_all False as
    shipped) and get_cls_head (:25-35) in eval mode.  ``probs=True`` returns what SoftmaxOutput does."""

    def __init__(self, cfg: StackCfg, params, device, precision="fp32"):
        self.cfg = cfg
        self.enc = GridGcnEncoder(cfg, params["enc"], device, precision=precision)
        self.head = gridconv.ClsHead(params["head"], device, precision=precision)

    def __call__(self, data, actual_numpoints, probs=False, keep_trace=False):
        table = self.enc(data, actual_numpoints, keep_trace=keep_trace)
        B, O, _ = table.shape
        feats = table[:, :, 4:]                      # (B, O, C); the reference flattens (B, C, O)
        feats = feats.reshape(B, -1) if O == 1 else feats.transpose(1, 2).reshape(B, -1)
        return self.head(feats.contiguous(), probs=probs)


# ---------------------------------------------------------------------------------------------------
# Reference-compatible configuration: the keys of segmentation/configs/configs.yaml
# ---------------------------------------------------------------------------------------------------
def from_reference_config(conf, query="gridify"):
    """Builds (StackCfg, UpCfg) from a dict with the reference's YAML keys
    (segmentation/configs/configs.yaml:54-111): voxel_size_lst, grid_size_lst, max_p_grid_lst,
    max_o_grid_lst, kernel_size_lst, stride_lst, pt_ele_dim, lidar_coord, loc_within, attfdim, relu,
    num_points, up_max_p_grid_lst, up_kernel_size_lst, up_pt_ele_dim, up_center_dim, up_gcn_outDim,
    up_attfdim, up_neigh_fetch, real_knn, up_voxel_size_lst, up_grid_size_lst, up_max_o_grid_lst; and the
    classification keys att_ele_dim, att_full, localfdim (classification/configs/configs.yaml:44-68).  Only what the
    fused kernels implement is accepted -- cubic voxels/grids, aggtype gcn, max pooling, no context MLP, empty
    gcn_outDim / elevation, concat centre integration, use_bn t -- everything else raises NotImplementedError
    instead of being ignored."""
    def cubic(v):
        v = list(v)
        if len(set(v)) != 1:
            raise NotImplementedError("anisotropic voxel / grid sizes are not supported by StackCfg: %r" % (v,))
        return v[0]
    for key, ok in (("aggtype", ("gcn",)), ("agg", ("max_pooling", "max")), ("up_aggtype", ("gcn",)),
                    ("up_agg", ("max_pooling", "max")), ("up_center_inte", ("concat",)), ("use_bn", ("t", True))):
        if conf.get(key, ok[0]) not in ok:
            raise NotImplementedError("%s=%r is not supported" % (key, conf.get(key)))
    def empty(v):  # None, [], [[], [], []]
        return not v or all(not x for x in v)
    for key in ("cntxt_mlp_lst", "up_cntxt_mlp_lst", "elevation", "gcn_outDim"):
        if not empty(conf.get(key)):
            raise NotImplementedError("%s=%r is not supported by the fused kernels" % (key, conf.get(key)))
    if conf.get("up_att_full") or conf.get("group_all") or conf.get("fps") or conf.get("reverse_index"):
        raise NotImplementedError("up_att_full / group_all / fps / reverse_index are not supported")
    cls_block = bool(conf.get("att_ele_dim")) or bool(conf.get("att_full")) or conf.get("localfdim", 0) != 0
    if cls_block and conf.get("up_max_p_grid_lst"):
        raise NotImplementedError("the classification block (att_ele_dim / att_full / localfdim) has no decoder")
    n = len(conf["max_o_grid_lst"])
    att = conf.get("att_ele_dim") or [()] * n
    layers = [LayerCfg(float(cubic(conf["voxel_size_lst"][i])), int(cubic(conf["grid_size_lst"][i])),
                       int(conf["max_o_grid_lst"][i]), int(conf["max_p_grid_lst"][i]),
                       int(conf["kernel_size_lst"][i]), list(conf["pt_ele_dim"][i]),
                       int(conf.get("stride_lst", [1] * n)[i]), att_ele_lst=tuple(att[i])) for i in range(n)]
    cfg = StackCfg(str(conf.get("save_model_prefix", "reference_config")), int(conf["num_points"]), layers,
                   coord_shift=tuple(float(x) for x in conf.get("lidar_coord", (1.0, 1.0, 1.0))),
                   loc=1 if conf.get("loc_within", True) else 0, attfdim=int(conf.get("attfdim", 10)),
                   pre_relu=bool(conf.get("relu", True)), query=query,
                   att_full=str(conf.get("att_full") or ""), localfdim=int(conf.get("localfdim", 0)))
    up = None
    if conf.get("up_max_p_grid_lst"):
        def same(key):
            v = conf[key]
            if any(list(x) != list(v[0]) for x in v) if isinstance(v[0], (list, tuple)) else len(set(v)) != 1:
                raise NotImplementedError("%s must be the same on every decoder level" % key)
            return v[0]
        nu = len(conf["up_max_p_grid_lst"])
        if nu != n:
            raise NotImplementedError("one decoder level per encoder level is expected (%d vs %d)" % (nu, n))
        fetch = "gridifyup" if not conf.get("up_neigh_fetch", True) else ("knn" if conf.get("real_knn") else "ballknn")
        levels = [conf["num_points"]] + [int(o) for o in conf["max_o_grid_lst"]]  # rows of level 0 .. n
        up_o = [int(o) for o in conf.get("up_max_o_grid_lst", [])]
        if up_o and up_o != [levels[n - 1 - i] for i in range(n)]:
            raise NotImplementedError("up_max_o_grid_lst %r does not match the encoder levels %r" % (up_o, levels))
        up = UpCfg(max_p_grid=int(same("up_max_p_grid_lst")), kernel_size=int(same("up_kernel_size_lst")),
                   pt_mlp_lst=tuple(same("up_pt_ele_dim")), center_dim=tuple(same("up_center_dim")),
                   out_dim=tuple(same("up_gcn_outDim")), attfdim=int(conf.get("up_attfdim", 10)),
                   neigh_fetch=fetch,
                   voxel_size_lst=tuple(float(cubic(v)) for v in conf.get("up_voxel_size_lst", [])),
                   grid_size_lst=tuple(int(cubic(v)) for v in conf.get("up_grid_size_lst", [])),
                   max_o_grid_lst=tuple(up_o))
    return cfg, up


def load_reference_yaml(path, query="gridify"):
    """Reads a YAML file written for the reference (the ACTIVE, un-commented block) and returns
    (StackCfg, UpCfg)."""
    import yaml
    with open(path) as f:
        conf = yaml.safe_load(f)
    return from_reference_config(conf, query)
