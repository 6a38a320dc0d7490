"""Host side of the fused GridConv block (reference: segmentation/models/gcn_module_g_att.py).

``GridConv`` holds one layer's parameters with eval-mode BatchNorm folded into the 1x1 convs and
launches ONE fused kernel (C-ABI ``gridgcn_gridconv_fwd``) that subsumes ``batch_take_g``
(utils/ops.py:78-93) and ``sub_g_update`` (gcn_module_g_att.py:172-287).  ``sub_g_update`` below
keeps the reference's call signature for callers that already hold gathered neighbours.

Parameter names follow the reference's scopes so a converted checkpoint can be loaded by name:
  {scope}/conv{j}_weight, _bias, {scope}/conv{j}/bn_gamma, _beta, _moving_mean, _moving_var
  {scope}/update_att_mlp2d_frst/conv1_*, {scope}/update_att_mlp2d_scnd/conv1_*
(utils/ops.py:149-158, gcn_module_g_att.py:135,141,152).
"""
import ctypes

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-3  # MXNet BatchNorm default eps (utils/ops.py:152 does not override it)

PRECISION = {"fp32": 0, "tf32": 1, "tf32x3": 2}
ATT_FULL = {"": 0, "next": 1, "last": 2}  # GRIDGCN_ATT_FULL_*


def init_stage(rng, cin, cout):
    """Random parameters of one conv1x1 + BN stage (Xavier-like weights, non-trivial BN stats)."""
    lim = np.sqrt(6.0 / (cin + cout))
    return dict(weight=rng.uniform(-lim, lim, size=(cout, cin)).astype(np.float32),
                bias=rng.uniform(-0.1, 0.1, size=cout).astype(np.float32),
                gamma=rng.uniform(0.8, 1.2, size=cout).astype(np.float32),
                beta=rng.uniform(-0.1, 0.1, size=cout).astype(np.float32),
                moving_mean=rng.uniform(-0.1, 0.1, size=cout).astype(np.float32),
                moving_var=rng.uniform(0.5, 1.5, size=cout).astype(np.float32))


def init_layer(rng, cin, pt_mlp_lst, attfdim, att_ele_lst=None, att_full="", localfdim=0):
    """Parameter dict of one GridConv layer, keyed like the reference's scopes.

    Defaults = the segmentation block (gcn_module_g_att.py: attention MLP att_vec -> C/4 -> C, features
    = gathered rows).  ``att_ele_lst`` (explicit attention widths, last one forced to C), ``att_full``
    ("next": the attention MLP's later stages also see the feature MLP's output; "last": its input) and
    ``localfdim`` = 3 (geo vector concatenated in front of the gathered features) give the
    classification block (classification/models/gcn_module_g.py:64-114,186-191)."""
    if att_full in ("off", None):
        att_full = ""
    feat_in = 3 if cin == 0 else cin + (3 if localfdim else 0)
    stages, w = [], feat_in
    for c in pt_mlp_lst:
        stages.append(init_stage(rng, w, c))
        w = c
    C = pt_mlp_lst[-1]
    att = []
    if attfdim > 0:
        ain = 3 if attfdim <= 3 else (4 if attfdim < 10 else 10)
        widths = [C // 4, C] if not att_ele_lst else list(att_ele_lst[:-1]) + [C]  # att_ele_lst[-1] = C, :85
        att = [init_stage(rng, ain, widths[0])]
        w = widths[0] + (C if att_full == "next" else feat_in if att_full == "last" else 0)
        for c in widths[1:]:
            att.append(init_stage(rng, w, c))
            w = c
    return dict(feat=stages, att=att, attfdim=attfdim, cin=cin, att_full=att_full, localfdim=int(localfdim))


def fold_bn(weight, bias, gamma, beta, moving_mean, moving_var, eps=BN_EPS):
    """conv1x1 -> BatchNorm(eval, fix_gamma=False) == conv1x1 with W' = W*s, b' = (b-mean)*s+beta,
    s = gamma / sqrt(var + eps).  Folded in float64, returned as float32."""
    s = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(moving_var, np.float64) + eps)
    w = np.asarray(weight, np.float64).reshape(len(s), -1) * s[:, None]
    b = (np.asarray(bias, np.float64) - np.asarray(moving_mean, np.float64)) * s + \
        np.asarray(beta, np.float64)
    return w.astype(np.float32), b.astype(np.float32)


def named_params(layer, scope):
    """Flatten a layer dict ({'feat': [...], 'att': [...]}) into reference-style names."""
    out = {}
    def put(prefix, st):
        out[prefix + "_weight"] = st["weight"]
        out[prefix + "_bias"] = st["bias"]
        for k in ("gamma", "beta", "moving_mean", "moving_var"):
            out[prefix + "/bn_" + k] = st[k]
    for j, st in enumerate(layer["feat"]):
        put("%s/conv%d" % (scope, j + 1), st)
    if layer["att"]:
        put(scope + "/update_att_mlp2d_frst/conv1", layer["att"][0])
        for j, st in enumerate(layer["att"][1:]):
            put(scope + "/update_att_mlp2d_scnd/conv%d" % (j + 1), st)
    return out


class GridConv:
    """One GridConv layer: ``out_table = layer(table, nebidx, cent, centmsk)``.

    table  (B, Nprev, 4+Cin) f32  rows [x y z w | feats]   nebidx (B, O, K) i32
    cent   (B, O, 4) f32                                   centmsk (B, O) f32
    out    (B, O, 4+Cout) f32     rows [cent | feats] -- the next layer's table
                                  (ggcn_models_g.py:186); ``features_nco(out)`` gives the
                                  reference's (B, Cout, O) layout.
    """

    def __init__(self, layer, device, pre_relu=True, precision="fp32"):
        self.cin = int(layer["cin"])
        self.attfdim = int(layer["attfdim"])
        self.pre_relu = bool(pre_relu)
        self.precision = precision
        self.device = torch.device(device)
        stages = list(layer["feat"]) + list(layer["att"])
        self.widths = [int(st["weight"].shape[0]) for st in stages]
        self.cout = int(layer["feat"][-1]["weight"].shape[0])
        self._w, self._b = [], []
        for st in stages:
            w, b = fold_bn(st["weight"], st["bias"], st["gamma"], st["beta"], st["moving_mean"],
                           st["moving_var"])
            self._w.append(torch.from_numpy(w).to(self.device).contiguous())
            self._b.append(torch.from_numpy(b).to(self.device).contiguous())
        self.localfdim = int(layer.get("localfdim", 0)) if self.cin > 0 else 0
        self.att_full = (layer.get("att_full", "") or "") if layer["att"] else ""
        if self.att_full == "off":
            self.att_full = ""
        if self.att_full not in ATT_FULL or self.localfdim not in (0, 3):
            raise ValueError("unsupported att_full %r / localfdim %r" % (self.att_full, self.localfdim))
        d = _lib.MlpDesc()
        d.n_feat_stages = len(layer["feat"])
        d.attfdim = self.attfdim
        d.feat_in = 3 if self.cin == 0 else self.cin + self.localfdim
        d.pre_relu = 1 if self.pre_relu else 0
        d.n_att_stages = len(layer["att"])
        d.localfdim = self.localfdim
        d.att_full = ATT_FULL[self.att_full]
        for i, (w, b) in enumerate(zip(self._w, self._b)):
            d.widths[i] = self.widths[i]
            d.weight[i] = w.data_ptr()
            d.bias[i] = b.data_ptr()
        self._desc = d
        self._packed = None
        if precision != "fp32":  # tensor-core paths: one-time re-layout of the weights on the device
            L = _lib.lib()
            nbytes = L.gridgcn_gridconv_packed_bytes(ctypes.byref(d), self.cin)
            if nbytes == 0:
                raise _lib.GridGcnError("this MLP shape is not supported by the tensor-core GridConv (widths that "
                                        "are not multiples of 4 run with precision='fp32')")
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                rc = L.gridgcn_gridconv_pack(ctypes.byref(d), self.cin, self._packed.data_ptr(), nbytes,
                                             torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(rc, "gridgcn_gridconv_pack")

    def __call__(self, table, nebidx, cent, centmsk, out=None):
        L = _lib.lib()
        if not table.is_cuda:
            raise _lib.GridGcnError("GridConv needs CUDA tensors: there is no CPU path")
        table = table.contiguous()
        nebidx = nebidx.contiguous()
        cent = cent.contiguous()
        centmsk = centmsk.contiguous()
        B, Nprev, roww = table.shape
        if roww != 4 + self.cin:
            raise ValueError("table rows should be 4+%d wide, got %d" % (self.cin, roww))
        _, O, K = nebidx.shape
        if out is None:
            out = torch.empty((B, O, 4 + self.cout), dtype=torch.float32, device=table.device)
        ws, ws_bytes, packed = None, 0, None
        if self._packed is not None:
            packed = self._packed.data_ptr()
            ws_bytes = max(L.gridgcn_gridconv_workspace_bytes(ctypes.byref(self._desc), B, Nprev, self.cin),
                           L.gridgcn_gridconv_edge_workspace_bytes(ctypes.byref(self._desc), B, self.cin, O, K))
        else:  # fp32: activation scratch for layers too wide for shared memory
            ws_bytes = L.gridgcn_gridconv_fp32_scratch_bytes(ctypes.byref(self._desc), self.cin, K)
        if ws_bytes:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=table.device)
        with torch.cuda.device(table.device):
            rc = L.gridgcn_gridconv_fwd(
                table.data_ptr(), nebidx.data_ptr(), cent.data_ptr(), centmsk.data_ptr(), B, Nprev,
                self.cin, O, K, ctypes.byref(self._desc), PRECISION[self.precision], packed,
                ws.data_ptr() if ws is not None else None, ws_bytes, out.data_ptr(),
                torch.cuda.current_stream(table.device).cuda_stream)
        _lib.check(rc, "gridgcn_gridconv_fwd")
        return out


_SHIM_CACHE = {}  # sub_g_update(): (id(layer), device, pre_relu, precision) -> (layer, GridConv)


def features_nco(out_table):
    """(B, O, 4+C) [cent | feats] -> (B, C, O), the layout sub_g_update returns (BN=True path)."""
    return out_table[:, :, 4:].transpose(1, 2)


def sub_g_update(centers_xyz, center_den, neighbors, has_feats, center_masks, neighbor_masks,
                 attfdim, center_ori_feats=None, pt_mlp_lst=None, outDim=(), cntxt_mlp=None,
                 shape=None, scope="layer", aggtype="gcn", pool_type="max_pooling", att_full="",
                 center_dim=(), recalden=False, bn_decay=0.9, *, layer=None, pre_relu=True,
                 precision="fp32"):
    """Reference call signature (gcn_module_g_att.py:172-174) on already gathered neighbours.

    centers_xyz (B,3,O), center_den (B,1,O), neighbors (B,4+C,O,P), center_masks (B,O) -> (B,C,O).
    ``layer`` carries the parameters.  The gathered tensor is viewed as a table with an identity
    index so that the same fused kernel runs.  Supported: aggtype gcn, max pooling, no context MLP, empty
    outDim -- what the shipped seg and cls configs use; ``att_full`` ("next" / "last", the classification
    block, fp32 precision) must agree with the one ``layer`` was initialised with."""
    if aggtype != "gcn" or pool_type not in ("max_pooling", "max") or cntxt_mlp is not None \
            or len(outDim) != 0 or center_ori_feats is not None:
        raise NotImplementedError("sub_g_update: unsupported configuration for the fused kernel")
    if (att_full or "") not in ("", "off") and att_full != layer.get("att_full", ""):
        raise ValueError("att_full=%r but the layer was initialised with %r" % (att_full, layer.get("att_full", "")))
    B, C4, O, P = neighbors.shape
    table = neighbors.permute(0, 2, 3, 1).reshape(B, O * P, C4).contiguous()
    idx = torch.arange(O * P, dtype=torch.int32, device=neighbors.device).reshape(1, O, P)
    idx = idx.expand(B, O, P).contiguous()
    cent = torch.cat([centers_xyz.transpose(1, 2), center_den.transpose(1, 2)], dim=2).contiguous()
    if center_masks is None:
        center_masks = torch.ones((B, O), dtype=torch.float32, device=neighbors.device)
    key = (id(layer), str(neighbors.device), bool(pre_relu), precision)
    conv = _SHIM_CACHE.get(key)
    if conv is None or conv[0] is not layer:  # BN folding + weight packing happen once per layer, not per call
        conv = (layer, GridConv(layer, neighbors.device, pre_relu=pre_relu, precision=precision))
        _SHIM_CACHE[key] = conv
    return features_nco(conv[1](table, idx, cent, center_masks))


def init_up_layer(rng, cd, cu, pt_mlp_lst, attfdim, center_dim, out_dim):
    """Parameters of one decoder layer: neighbour features Cd -> pt_mlp, attention, centre branch on the
    finer level's [cent | feat] rows (4+Cu -> center_dim), update MLP (center_dim[-1]+pt[-1] -> out_dim)."""
    layer = init_layer(rng, cd, pt_mlp_lst, attfdim)
    w, stages = 4 + cu, []
    for c in center_dim:
        stages.append(init_stage(rng, w, c))
        w = c
    layer["center"] = stages
    w, stages = w + pt_mlp_lst[-1], []
    for c in out_dim:
        stages.append(init_stage(rng, w, c))
        w = c
    layer["update"] = stages
    return layer


def _folded(st, device, bn=True):
    if bn:
        w, b = fold_bn(st["weight"], st["bias"], st["gamma"], st["beta"], st["moving_mean"], st["moving_var"])
    else:
        w, b = np.asarray(st["weight"], np.float32).reshape(len(st["bias"]), -1), np.asarray(st["bias"], np.float32)
    return torch.from_numpy(w).to(device).contiguous(), torch.from_numpy(b).to(device).contiguous()


def rowmlp(in1, in2, w, b, *, relu_in=False, relu_out=True, row_scale=None, out=None, out_col=0, cent=None, tc=False):
    """out[..., out_col:out_col+Cout] = act(W act_in([in1 | in2]) + b) * row_scale   (C-ABI gridgcn_rowmlp_fwd;
    ``tc=True``: gridgcn_rowmlp_tc_fwd, the tf32x3 tensor-core GEMM).
    in1 / in2: (..., C) views whose last dim is contiguous and whose rows are equally strided."""
    L = _lib.lib()
    rows = int(np.prod(in1.shape[:-1]))
    c1, ld1 = in1.shape[-1], in1.stride(-2)
    c2, ld2, p2 = 0, 0, None
    if in2 is not None:
        c2, ld2, p2 = in2.shape[-1], in2.stride(-2), in2.data_ptr()
    cout = w.shape[0]
    if out is None:
        out = torch.empty(tuple(in1.shape[:-1]) + (out_col + cout,), dtype=torch.float32, device=in1.device)
    ldo = out.stride(-2)
    with torch.cuda.device(in1.device):
        fn = L.gridgcn_rowmlp_tc_fwd if tc else L.gridgcn_rowmlp_fwd
        rc = fn(in1.data_ptr(), ld1, c1, p2, ld2, c2, w.data_ptr(), b.data_ptr(), cout,
                                  1 if relu_in else 0, 1 if relu_out else 0,
                                  row_scale.data_ptr() if row_scale is not None else None,
                                  out.data_ptr() + 4 * out_col, ldo,
                                  cent.data_ptr() if cent is not None else None,
                                  out.data_ptr() if cent is not None else None, rows,
                                  torch.cuda.current_stream(in1.device).cuda_stream)
    _lib.check(rc, "gridgcn_rowmlp_tc_fwd" if tc else "gridgcn_rowmlp_fwd")
    return out


class GridConvUp:
    """One decoder GridConv layer (ggcn_models_g.py:191-231): aggregated neighbour features from the
    fused GridConv kernels (no pre-ReLU / mask: they follow the concat), centre branch, concat, update MLP
    and centre mask as row-MLP stages.  ``out = layer(f_last, nebidx, cent_up, f_this, centmsk)``."""

    def __init__(self, layer, device, pre_relu=True, precision="tf32x3"):
        self.device = torch.device(device)
        self.core = GridConv(layer, device, pre_relu=False, precision=precision)
        self.center = [_folded(st, self.device) for st in layer["center"]]
        self.update = [_folded(st, self.device) for st in layer["update"]]
        self.pre_relu = bool(pre_relu)
        self.tc = precision != "fp32"  # per-centre stages on the tensor cores too (exact fp32 FMA for "fp32")
        self.cout = int(layer["update"][-1]["weight"].shape[0]) if layer["update"] else None

    def __call__(self, f_last, nebidx, cent_up, f_this, centmsk=None):
        B, O, _ = nebidx.shape
        ones = torch.ones((B, O), dtype=torch.float32, device=f_last.device)
        agg = self.core(f_last, nebidx, cent_up, ones)[:, :, 4:]          # (B, O, C) strided view
        cf = f_this
        for w, b in self.center:
            cf = rowmlp(cf, None, w, b, tc=self.tc)
        x, x2 = cf, agg
        for i, (w, b) in enumerate(self.update):
            last = i + 1 == len(self.update)
            x = rowmlp(x, x2, w, b, relu_in=self.pre_relu and i == 0, row_scale=centmsk if last else None,
                       out_col=4 if last else 0, cent=cent_up.contiguous() if last else None, tc=self.tc)
            x2 = None
        return x


class SegHead:
    """get_seg_head up to the logits (ggcn_models_g.py:30-36, eval mode)."""

    def __init__(self, head, device, precision="fp32"):
        self.l1 = _folded(head[0], device)
        self.l2 = _folded(head[1], device, bn=False)
        self.tc = precision != "fp32"

    def __call__(self, feats):
        x = rowmlp(feats, None, *self.l1, tc=self.tc)
        return rowmlp(x, None, *self.l2, relu_out=False, tc=self.tc)


class ClsHead:
    """get_cls_head up to the class scores (classification/models/ggcn_models_g.py:25-35, eval mode: BatchNorm
    folded, Dropout = identity): FC 512 -> BN -> ReLU -> FC 256 -> BN -> ReLU -> FC num_classes; ``probs=True``
    adds the softmax SoftmaxOutput applies at inference.  Runs on the row-MLP kernel (gridgcn_rowmlp_fwd)."""

    def __init__(self, head, device, precision="fp32"):
        self.l1 = _folded(head[0], device)
        self.l2 = _folded(head[1], device)
        self.l3 = _folded(head[2], device, bn=False)
        self.tc = precision != "fp32"

    def __call__(self, feats, probs=False):
        x = rowmlp(feats, None, *self.l1, tc=self.tc)
        x = rowmlp(x, None, *self.l2, tc=self.tc)
        x = rowmlp(x, None, *self.l3, relu_out=False, tc=self.tc)
        return torch.softmax(x, dim=-1) if probs else x
