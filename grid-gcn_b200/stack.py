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


@dataclass
class StackCfg:
    name: str
    num_points: int
    layers: List[LayerCfg]
    coord_shift: Sequence[float] = (1.0, 1.0, 1.0)  # lidar_coord
    loc: int = 1                                     # loc_within: True
    attfdim: int = 10
    pre_relu: bool = True                            # configs["relu"]
    query: str = "gridifyknn"                        # or "gridify" (what the seg graph calls)
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
        layers.append(gridconv.init_layer(rng, cin, list(l.pt_mlp_lst), cfg.attfdim))
        cin = l.pt_mlp_lst[-1]
    return layers


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
        query = ops.GridifyKNN if cfg.query == "gridifyknn" else ops.Gridify
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
