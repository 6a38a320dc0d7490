"""CPU restatement of the GridConv block -- TEST INFRASTRUCTURE (numpy fp32, op by op).

Follows, op for op, the symbol graph the reference builds for one encoder layer:
  batch_take_g                       utils/ops.py:78-93   (take mode "clip" after the arange(B)*N offset)
  transpose to (B, 4+C, O, P)        segmentation/models/ggcn_models_g.py:172-175
  sub_g_update                       segmentation/models/gcn_module_g_att.py:172-287
    geo_vec, geo_dist                :189-194
    att_vec by attfdim               :209-222
    neighbor_feats (geo when !has_feats) :242-250
    verts_pair_func                  :120-170  (mlp2d_c = conv1x1 -> BatchNorm(eval) -> relu,
                                                utils/ops.py:149-158,254-260)
    max pooling over P               :57-59  (no neighbour mask for max pooling, :259)
    update_func pre-relu             :31-32
    centre mask                      :284-285
  concat(centers, center_feats)      ggcn_models_g.py:186
BatchNorm runs in eval mode (moving statistics), eps = 1e-3 (MXNet BatchNorm default), fix_gamma
False -- the mode the reference's own timing harness uses (train_gpu_speed_profiling.py:108).

Parity status: "parity unpinned" (the reference ships no test or golden value for this block and
MXNet cannot be imported here); this restatement is the pin.  Only tests/, smoke() and bench.py's
CPU-baseline legs may import it.
"""
import numpy as np

F = np.float32
BN_EPS = 1e-3


def _conv_bn_relu(x, st):
    """x: (B, Cin, O, P) -> (B, Cout, O, P); Convolution(kernel 1x1) -> BatchNorm eval -> relu."""
    y = np.einsum("oc,bcnp->bonp", st["weight"], x, optimize=True).astype(F)
    y = (y + st["bias"][None, :, None, None]).astype(F)
    inv = (F(1.0) / np.sqrt(st["moving_var"] + F(BN_EPS))).astype(F)
    y = ((y - st["moving_mean"][None, :, None, None]) * inv[None, :, None, None]).astype(F)
    y = (y * st["gamma"][None, :, None, None] + st["beta"][None, :, None, None]).astype(F)
    return np.maximum(y, F(0))


def batch_take_g(data, index):
    """data (B, N, C), index (B, O, P) -> (B, O, P, C); utils/ops.py:78-93."""
    B, N, C = data.shape
    flat = data.reshape(B * N, C)
    gi = index.astype(np.int64) + (np.arange(B, dtype=np.int64) * N)[:, None, None]
    gi = np.clip(gi, 0, B * N - 1)  # mx.symbol.take default mode = clip
    return flat[gi]


def gridconv_layer(table, nebidx, cent, centmsk, layer, pre_relu=True):
    """One encoder GridConv layer.  table (B,Nprev,4+Cin), nebidx (B,O,P), cent (B,O,4),
    centmsk (B,O) -> next table (B,O,4+Cout) = concat(cent, feats)."""
    table = np.asarray(table, F)
    cent = np.asarray(cent, F)
    B, O, P = nebidx.shape
    has_feats = layer["cin"] > 0
    neighbors = batch_take_g(table, nebidx).transpose(0, 3, 1, 2)  # (B, 4+C, O, P)
    centers_xyz = cent[:, :, :3].transpose(0, 2, 1)                # (B, 3, O)
    centers_expand = np.broadcast_to(centers_xyz[:, :, :, None], (B, 3, O, P))
    nloc = neighbors[:, 0:3]
    geo_vec = (nloc - centers_expand).astype(F)
    geo_dist = np.sqrt(np.sum(np.square(geo_vec), axis=1, keepdims=True)).astype(F)
    attfdim = layer["attfdim"]
    if attfdim <= 3:
        att_vec = geo_vec
    elif attfdim == 4:
        att_vec = np.concatenate([geo_dist, geo_vec], axis=1)
    elif attfdim == 10:
        att_vec = np.concatenate([geo_dist, geo_vec, centers_expand, nloc], axis=1)
    else:
        raise NotImplementedError("attfdim %d" % attfdim)
    # neighbour features: the geo vector when the layer has no input features; with localfdim != 0 (the
    # classification config, classification/models/gcn_module_g.py:165-171,186-191) the geo features are
    # concatenated in front of the gathered ones; localfdim == 0 is the seg config
    localfdim = int(layer.get("localfdim", 0))
    if localfdim > 3:
        raise NotImplementedError("localfdim %d" % localfdim)
    if not has_feats:
        feats = geo_vec
    elif localfdim != 0:
        feats = np.concatenate([geo_vec, neighbors[:, 4:]], axis=1)
    else:
        feats = neighbors[:, 4:]
    ori_feats = feats
    for st in layer["feat"]:
        feats = _conv_bn_relu(feats, st)
    if attfdim > 0:
        # verts_pair_func: first attention stage alone ("update_att_mlp2d_frst"), optional concat with the
        # MLP output (att_full "next") or its input ("last"), remaining stages ("update_att_mlp2d_scnd")
        # -- classification/models/gcn_module_g.py:83-97; the seg module (gcn_module_g_att.py:141-152)
        # is the two-stage special case without concat
        att_full = layer.get("att_full", "") or ""
        a = _conv_bn_relu(att_vec, layer["att"][0])
        if att_full == "last":
            a = np.concatenate([a, ori_feats], axis=1)
        elif att_full == "next":
            a = np.concatenate([a, feats], axis=1)
        elif att_full not in ("", "off"):
            raise NotImplementedError("att_full %r" % att_full)
        for st in layer["att"][1:]:
            a = _conv_bn_relu(a, st)
        pair = (a * feats).astype(F)
    else:
        pair = feats
    agg = pair.max(axis=3)  # (B, C, O)
    if pre_relu:
        agg = np.maximum(agg, F(0))
    agg = (agg * np.asarray(centmsk, F)[:, None, :]).astype(F)
    return np.concatenate([cent, agg.transpose(0, 2, 1)], axis=2).astype(F)


def _conv1d_bn_relu(x, st, relu=True, bn=True):
    """x: (B, Cin, O) -> (B, Cout, O); Convolution(kernel 1) -> BatchNorm eval -> relu
    (utils/ops.py:141-147, mlp1d_c :236-242)."""
    y = np.einsum("oc,bcn->bon", st["weight"], x, optimize=True).astype(F)
    y = (y + st["bias"][None, :, None]).astype(F)
    if bn:
        inv = (F(1.0) / np.sqrt(st["moving_var"] + F(BN_EPS))).astype(F)
        y = ((y - st["moving_mean"][None, :, None]) * inv[None, :, None]).astype(F)
        y = (y * st["gamma"][None, :, None] + st["beta"][None, :, None]).astype(F)
    return np.maximum(y, F(0)) if relu else y


def gridconv_up_layer(f_last, nebidx, cent_up, f_this, centmsk, layer, pre_relu=True):
    """One DECODER GridConv layer (segmentation/models/ggcn_models_g.py:191-231 + sub_g_update with
    center_ori_feats, gcn_module_g_att.py:267-285).

    f_last  (B, Nd, 4+Cd)  coarser level's [cent | feat] rows (gathered through nebidx, BallKNN/GridifyUp)
    nebidx  (B, O, K)      cent_up (B, O, 4)  centres of the finer level
    f_this  (B, O, 4+Cu)   finer level's own [cent | feat] rows (`center_ori_feats`)
    centmsk (B, O) or None
    layer: dict(feat=[stage], att=[stage, stage], center=[stage], update=[stage], attfdim, cin=Cd)
    -> (B, O, 4+Cout) = concat(cent_up, feats)  (ggcn_models_g.py:231)"""
    f_last = np.asarray(f_last, F)
    cent_up = np.asarray(cent_up, F)
    B, O, K = nebidx.shape
    ones = np.ones((B, O), F)
    # aggregated neighbour features: the encoder block without pre-ReLU / mask (they come after the concat)
    agg = gridconv_layer(f_last, nebidx, cent_up, ones, layer, pre_relu=False)[..., 4:]   # (B, O, C)
    agg = agg.transpose(0, 2, 1)                                                          # (B, C, O)
    center_ori = np.asarray(f_this, F).transpose(0, 2, 1)                                 # (B, 4+Cu, O)
    center_feats = center_ori
    for st in layer["center"]:                                  # mlp1d_c(center_ori_feats, center_dim) :269
        center_feats = _conv1d_bn_relu(center_feats, st)
    x = np.concatenate([center_feats, agg], axis=1)             # up_center_inte == "concat" :281-282
    if pre_relu:
        x = np.maximum(x, F(0))                                 # update_func :31-32
    for st in layer["update"]:                                  # mlp1d_c(outDim) :33-36
        x = _conv1d_bn_relu(x, st)
    if centmsk is not None:
        x = (x * np.asarray(centmsk, F)[:, None, :]).astype(F)  # :284-285
    return np.concatenate([cent_up, x.transpose(0, 2, 1)], axis=2).astype(F)


def seg_head(feats, head):
    """get_seg_head (ggcn_models_g.py:30-36) in eval mode up to the logits: conv1d 128 + BN + relu,
    dropout = identity, conv1d 21 without BN / relu.  feats (B, O, C) -> logits (B, O, 21)."""
    x = np.asarray(feats, F).transpose(0, 2, 1)
    x = _conv1d_bn_relu(x, head[0])
    x = _conv1d_bn_relu(x, head[1], relu=False, bn=False)
    return x.transpose(0, 2, 1).astype(F)
