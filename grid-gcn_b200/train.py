"""Training-mode GridConv block and the data-parallel training step.

What this is: the block of segmentation/models/gcn_module_g_att.py:172-287 (and the classification
variants of classification/models/gcn_module_g.py) with BatchNorm in TRAINING mode -- batch statistics
over all B*O*P edges between consecutive 1x1 convs (utils/ops.py:149-158, `use_global_stats: False`,
segmentation/configs/configs.yaml:27), moving statistics updated with momentum bn_decay -- written
op by op on the reference's (B, C, O, P) layout with PyTorch tensor ops, so that autograd provides the
backward of gather / MLP / attention product / max-pool.  The index operators in front of it
(Gridify, GridifyKNN, Gridify_occaware, BallKNN) are this library's CUDA kernels; like the reference's
(gridify-inl.h:227-231) they have no gradient.  One step = forward, backward, ONE all-reduce of the
flattened gradient bucket (shard.allreduce_gradients; the north star's only collective), optimiser update.

This module is the autograd statement of the block (PyTorch's library kernels do the math).  The same block with
forward AND backward on this library's own kernels is train_cuda.GridConvTrainCuda (`block="cuda"` below; csrc/train_ops.cu
+ the tcgen05 row GEMM), held to this module within 1e-3 by tests/test_gpu_train.py.  `GridConvTrain.export_layer()` hands
the trained parameters to the fused inference kernels (BatchNorm folded into the convs), and the eval-mode forward of
this module is held to the same oracle as they are (tests), which is what ties the forms together.  `GraphedTrainStep`
replays a whole training step as CUDA graphs around the one gradient all-reduce.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import gridconv, shard

BN_EPS = gridconv.BN_EPS


class ConvBnRelu(nn.Module):
    """mlp2d_c stage: Convolution(kernel 1x1) -> BatchNorm(axis=1, fix_gamma=False, momentum=bn_decay)
    -> relu (utils/ops.py:149-158).  MXNet's `momentum` weighs the OLD moving value; torch's the new one."""

    def __init__(self, stage, bn_decay=0.9):
        super().__init__()
        w = torch.as_tensor(np.asarray(stage["weight"], np.float32))
        self.weight = nn.Parameter(w.reshape(w.shape[0], -1).clone())
        self.bias = nn.Parameter(torch.as_tensor(np.asarray(stage["bias"], np.float32)).clone())
        self.bn = nn.BatchNorm2d(w.shape[0], eps=BN_EPS, momentum=1.0 - bn_decay)
        with torch.no_grad():
            self.bn.weight.copy_(torch.as_tensor(stage["gamma"]))
            self.bn.bias.copy_(torch.as_tensor(stage["beta"]))
            self.bn.running_mean.copy_(torch.as_tensor(stage["moving_mean"]))
            self.bn.running_var.copy_(torch.as_tensor(stage["moving_var"]))

    def forward(self, x):  # (B, Cin, O, P)
        y = torch.einsum("oc,bcnp->bonp", self.weight, x) + self.bias[None, :, None, None]
        return F.relu(self.bn(y))

    def export(self):
        g = lambda t: t.detach().cpu().numpy().astype(np.float32)  # noqa: E731
        return dict(weight=g(self.weight), bias=g(self.bias), gamma=g(self.bn.weight), beta=g(self.bn.bias),
                    moving_mean=g(self.bn.running_mean), moving_var=g(self.bn.running_var))


class GridConvTrain(nn.Module):
    """One GridConv layer with trainable parameters.  ``forward(table, nebidx, cent, centmsk)`` has the
    fused layer's signature: table (B,Nprev,4+Cin), nebidx (B,O,K) int, cent (B,O,4), centmsk (B,O) ->
    (B,O,4+Cout) = concat(cent, feats) (ggcn_models_g.py:186)."""

    def __init__(self, layer, pre_relu=True, bn_decay=0.9):
        super().__init__()
        self.cin, self.attfdim = int(layer["cin"]), int(layer["attfdim"])
        self.localfdim = int(layer.get("localfdim", 0))
        self.att_full = layer.get("att_full", "") or ""
        self.pre_relu = bool(pre_relu)
        self.feat = nn.ModuleList([ConvBnRelu(st, bn_decay) for st in layer["feat"]])
        self.att = nn.ModuleList([ConvBnRelu(st, bn_decay) for st in layer["att"]])

    def forward(self, table, nebidx, cent, centmsk):
        B, Nprev, W = table.shape
        _, O, P = nebidx.shape
        # batch_take_g: take with mode "clip" AFTER the per-batch offset (utils/ops.py:90-92)
        gi = nebidx.long() + (torch.arange(B, device=table.device) * Nprev)[:, None, None]
        gi = gi.clamp_(0, B * Nprev - 1)
        nb = table.reshape(B * Nprev, W)[gi].permute(0, 3, 1, 2)       # (B, 4+C, O, P)
        cxyz = cent[:, :, :3].permute(0, 2, 1)[:, :, :, None].expand(B, 3, O, P)
        nloc = nb[:, 0:3]
        geo = nloc - cxyz
        dist = geo.square().sum(1, keepdim=True).sqrt()
        if self.attfdim <= 3:
            att_vec = geo
        elif self.attfdim == 4:
            att_vec = torch.cat([dist, geo], 1)
        else:
            att_vec = torch.cat([dist, geo, cxyz, nloc], 1)
        if self.cin == 0:
            feats = geo
        elif self.localfdim:
            feats = torch.cat([geo, nb[:, 4:]], 1)
        else:
            feats = nb[:, 4:]
        ori = feats
        for st in self.feat:
            feats = st(feats)
        if len(self.att):
            a = self.att[0](att_vec)
            if self.att_full == "next":
                a = torch.cat([a, feats], 1)
            elif self.att_full == "last":
                a = torch.cat([a, ori], 1)
            for st in self.att[1:]:
                a = st(a)
            feats = a * feats                                         # gcn_module_g_att.py:167
        agg = feats.amax(dim=3)                                       # max pooling over P (:57-59)
        if self.pre_relu:
            agg = F.relu(agg)
        agg = agg * centmsk[:, None, :]
        return torch.cat([cent, agg.permute(0, 2, 1)], 2)

    def export_layer(self):
        """Parameter dict for the fused inference kernels (gridconv.GridConv folds the BatchNorms)."""
        return dict(feat=[st.export() for st in self.feat], att=[st.export() for st in self.att],
                    attfdim=self.attfdim, cin=self.cin, att_full=self.att_full, localfdim=self.localfdim)


class GridGcnClassifier(nn.Module):
    """Encoder ladder (index operator -> GridConvTrain per layer) + global max pool + linear head: the shape
    of get_symbol_cls_ggcn (classification/models/ggcn_models_g.py:37-111) reduced to what a training step
    needs.  ``query(data_loc, num, layer_cfg) -> (nebidx, nebidxmsk, cent, centmsk, num)`` supplies the
    indices: the CUDA operators on a GPU (default), anything with the same outputs in CPU tests."""

    def __init__(self, cfg, params, num_classes=40, query=None, bn_decay=0.9, block="torch"):
        super().__init__()
        self.cfg = cfg
        if block == "cuda":  # forward + backward on the library's own kernels (train_cuda.py, csrc/train_ops.cu)
            from .train_cuda import GridConvTrainCuda as Block
        elif block == "torch":
            Block = GridConvTrain
        else:
            raise ValueError("block must be 'torch' or 'cuda'")
        self.layers = nn.ModuleList([Block(p, cfg.pre_relu, bn_decay) for p in params])
        self.head = nn.Linear(cfg.layers[-1].pt_mlp_lst[-1], num_classes)
        self._query = query

    def _default_query(self, loc, num, l):
        from . import stack
        fn = stack.query_fn(self.cfg)
        return fn(loc, num, max_o_grid=l.max_o_grid, max_p_grid=l.max_p_grid, kernel_size=l.kernel_size,
                  stride=l.stride, coord_shift=self.cfg.coord_shift, voxel_size=[l.voxel_size] * 3,
                  grid_size=[l.grid_size] * 3, loc=self.cfg.loc)

    def forward(self, data, actual_numpoints):
        q = self._query or self._default_query
        table, loc, num = data, data, actual_numpoints
        for l, conv in zip(self.cfg.layers, self.layers):
            with torch.no_grad():  # index operators: no gradient (gridify-inl.h:227-231)
                nebidx, _, cent, centmsk, num = q(loc, num, l)
            table = conv(table, nebidx, cent, centmsk)
            loc = cent
        feats = table[:, :, 4:]
        mask = table.new_ones(feats.shape[:2]) if centmsk is None else centmsk
        pooled = (feats - (1.0 - mask[:, :, None]) * 1e30).amax(dim=1)  # masked global max pool
        return self.head(pooled)


def train_step(model, optimizer, data, actual_numpoints, labels):
    """forward -> cross entropy -> backward -> ONE flat gradient all-reduce (mean over ranks) -> update.
    Returns the local loss (python float).  With equal shard sizes the averaged gradient is the gradient of
    the global mean loss; BatchNorm statistics stay per rank (the reference has no SyncBN either)."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(data, actual_numpoints), labels)
    loss.backward()
    shard.allreduce_gradients([p for p in model.parameters()])
    optimizer.step()
    return float(loss.detach())


class GraphedTrainStep:
    """The data-parallel training step replayed as CUDA graphs (the step is launch bound at the reference's batch
    sizes: ~300 small kernels for 3 clouds of 81920 points).

        graph 1:  zero the flat gradient bucket, forward, cross entropy, backward   (every p.grad is a VIEW of the bucket)
        eager  :  ONE all-reduce of the bucket over torch.distributed (NCCL), world > 1 only
        graph 2:  bucket /= world, optimizer.step()

    Same arithmetic as ``train_step``.  Inputs are copied into static buffers before each replay; shapes are fixed at
    construction.  Autograd graphs of earlier eager steps of the same model must be gone (``del loss``): their
    AccumulateGrad nodes are tied to the stream they ran on and would invalidate the capture.  The caller must not call ``optimizer.zero_grad(set_to_none=True)`` afterwards (it would detach the
    gradients from the bucket).  Parameters, buffers and optimiser state are restored after the warm-up iterations, so
    constructing this object does not train the model.  Stateful optimisers must be capture-safe (e.g.
    ``torch.optim.Adam(..., capturable=True)``)."""

    def __init__(self, model, optimizer, data, actual_numpoints, labels, warmup=3):
        if not data.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors")
        self.model, self.optimizer = model, optimizer
        self.data, self.num, self.labels = data.clone(), actual_numpoints.clone(), labels.clone()
        self.params = [p for p in model.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        self.bucket = torch.zeros(total, dtype=torch.float32, device=data.device)
        off = 0
        for p in self.params:
            p.grad = self.bucket[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.loss = None
        model.train()
        saved_model = {k: v.clone() for k, v in model.state_dict().items()}
        saved_opt = {p: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in optimizer.state.items()}
        side = torch.cuda.Stream(device=data.device)
        side.wait_stream(torch.cuda.current_stream(data.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._fwd_bwd()
                self._reduce()
                self._update()
        torch.cuda.current_stream(data.device).wait_stream(side)
        torch.cuda.synchronize(data.device)
        self.loss = None  # drops the warm-up autograd graph (its AccumulateGrad nodes are tied to the warm-up stream)
        with torch.no_grad():
            model.load_state_dict(saved_model)
        self._restore_optimizer(saved_opt)
        self.g1, self.g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g1):
            self._fwd_bwd()
        with torch.cuda.graph(self.g2, pool=self.g1.pool()):
            self._update()
        # capture executed nothing, but BatchNorm's python-side counters / moving statistics must match "no step yet"
        with torch.no_grad():
            model.load_state_dict(saved_model)

    def _restore_optimizer(self, saved):
        """Optimiser state back to what it was before the warm-up steps, IN PLACE (the graphs must see the same tensors
        on every replay): saved values where the state existed, zeros where the warm-up created it (the initial state
        of Adam / AdamW / RMSprop / momentum SGD without dampening)."""
        with torch.no_grad():
            for p, st in self.optimizer.state.items():
                old = saved.get(p, {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in old and torch.is_tensor(old[k]):
                            v.copy_(old[k])
                        else:
                            v.zero_()
                    elif k in old:
                        st[k] = old[k]

    def _fwd_bwd(self):
        self.bucket.zero_()
        self.loss = F.cross_entropy(self.model(self.data, self.num), self.labels)
        self.loss.backward()

    def _reduce(self):
        if self.world > 1:
            dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM)

    def _update(self):
        if self.world > 1:
            self.bucket.mul_(1.0 / self.world)
        self.optimizer.step()

    def __call__(self, data=None, actual_numpoints=None, labels=None):
        """One step; returns the (static) loss tensor of this rank -- no host synchronisation."""
        if data is not None:
            self.data.copy_(data, non_blocking=True)
        if actual_numpoints is not None:
            self.num.copy_(actual_numpoints, non_blocking=True)
        if labels is not None:
            self.labels.copy_(labels, non_blocking=True)
        self.g1.replay()
        self._reduce()
        self.g2.replay()
        return self.loss.detach()
