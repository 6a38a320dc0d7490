"""Training-mode GridConv block on this library's own kernels.

Same math and parameters as ``train.GridConvTrain`` (the block of segmentation/models/gcn_module_g_att.py:172-287 with
BatchNorm in TRAINING mode, utils/ops.py:149-158, ``use_global_stats: False`` configs.yaml:27) -- but the forward and
the backward run on hand-written CUDA kernels instead of PyTorch ops + autograd:

  * every 1x1 convolution of the forward (Z = X W^T + b) and every input gradient of the backward (dX = dZ W) is one
    launch of the persistent tcgen05 row GEMM (csrc/rowgemm_tc.cu, 3-pass tf32 = fp32-class accuracy);
  * csrc/train_ops.cu holds the rest: edge rows + gather indices, per-channel batch statistics (+ moving statistics),
    normalise + ReLU, max pool with arg-max, its routing backward, ReLU backward fused with the BatchNorm-backward
    sums, BatchNorm backward, weight + bias gradients, gather scatter-add.

``torch.autograd.Function`` is only the seam that hands ``grad_output`` in and the parameter gradients out; PyTorch
supplies device memory and the optimiser.  Supported: the segmentation flavour (localfdim 0, att_full off, two
attention stages) -- what the shipped ScanNet configs train.  The un-fused layout (one HBM round trip per operator,
like the reference's MXNet graph) is deliberate for this first version; tests/test_gpu_train.py holds it to the
autograd module within 1e-3.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib, gridconv
from .train import GridConvTrain

BN_EPS = gridconv.BN_EPS


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _gemm(x, w, b, relu):
    """rows x Cin @ (Cout x Cin)^T + b on the tensor-core row GEMM."""
    return gridconv.rowmlp(x, None, w, b, relu_out=relu, tc=True)


_ZEROS = {}


def _zeros(n, dev):
    """cached read-only zero vector (the bias of the backward GEMMs)"""
    key = (n, str(dev))
    if key not in _ZEROS:
        _ZEROS[key] = torch.zeros(n, dtype=torch.float32, device=dev)
    return _ZEROS[key]


class _Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, table, nebidx, cent, centmsk, *params):
        L = _lib.lib()
        dev = table.device
        B, Nprev, roww = table.shape
        _, O, K = nebidx.shape
        cin, C = mod.cin, mod.cout
        edges = B * O * K
        table, nebidx, cent, centmsk = table.contiguous(), nebidx.contiguous(), cent.contiguous(), centmsk.contiguous()
        fin_p = cin if cin > 0 else 4
        ain_p = mod.ain_p
        xf = torch.empty((edges, fin_p), dtype=torch.float32, device=dev)
        xa = torch.empty((edges, max(ain_p, 4)), dtype=torch.float32, device=dev)
        rowidx = torch.empty(edges, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.gridgcn_train_edge_rows(table.data_ptr(), nebidx.data_ptr(), cent.data_ptr(), B, Nprev, cin, O, K,
                                                 mod.attfdim, xf.data_ptr(), xa.data_ptr(), rowidx.data_ptr(), _stream(table)),
                       "gridgcn_train_edge_rows")
            saved = []
            widths = [int(st_.weight.shape[0]) for st_ in mod._stages()]
            pool = torch.zeros(4 * sum(widths), dtype=torch.float32, device=dev)   # ONE fill for every stage's statistics
            pool_off = [0]

            def stage(x, idx):
                w, b, gamma, beta = params[4 * idx:4 * idx + 4]
                wp = mod._pad_w(idx, w)                       # zero columns for the padded input layouts
                z = _gemm(x, wp, b, False)
                rows, Cout = z.shape
                st = pool[pool_off[0]:pool_off[0] + 4 * Cout].view(4, Cout)   # sum, sum of squares -> mean, invstd
                pool_off[0] += 4 * Cout
                mean, invstd = st[2], st[3]
                _lib.check(L.gridgcn_train_col_sums(z.data_ptr(), None, rows, Cout, st[0].data_ptr(), st[1].data_ptr(),
                                                    _stream(z)), "gridgcn_train_col_sums")
                bn = mod._stages()[idx].bn   # biased variance normalises; the moving one is unbiased (torch convention)
                _lib.check(L.gridgcn_train_bn_finalize(st[0].data_ptr(), st[1].data_ptr(), rows, Cout, BN_EPS, float(bn.momentum),
                                                       mean.data_ptr(), invstd.data_ptr(), bn.running_mean.data_ptr(),
                                                       bn.running_var.data_ptr(), bn.num_batches_tracked.data_ptr(), _stream(z)),
                           "gridgcn_train_bn_finalize")
                y = torch.empty_like(z)
                _lib.check(L.gridgcn_train_bn_relu_fwd(z.data_ptr(), rows, Cout, mean.data_ptr(), invstd.data_ptr(),
                                                       gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _stream(z)),
                           "gridgcn_train_bn_relu_fwd")
                saved.append((x, z, y, mean, invstd))
                return y

            x = xf
            for s in range(mod.n_feat):
                x = stage(x, s)
            F = x
            A = None
            if mod.n_att:
                a = xa
                for s in range(mod.n_att):
                    a = stage(a, mod.n_feat + s)
                A = a
            out = torch.empty((B, O, 4 + C), dtype=torch.float32, device=dev)
            out[:, :, :4] = cent
            argmax = torch.empty((B * O, C), dtype=torch.int32, device=dev)
            feats = out[:, :, 4:]
            _lib.check(L.gridgcn_train_pool_fwd(F.data_ptr(), A.data_ptr() if A is not None else None, B * O, K, C,
                                                1 if mod.pre_relu else 0, centmsk.data_ptr(), feats.data_ptr(), 4 + C,
                                                argmax.data_ptr(), _stream(F)), "gridgcn_train_pool_fwd")
        ctx.mod, ctx.saved, ctx.misc = mod, saved, (F, A, argmax, centmsk, rowidx, (B, Nprev, roww, O, K), params)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        mod, saved = ctx.mod, ctx.saved
        F, A, argmax, centmsk, rowidx, (B, Nprev, roww, O, K), params = ctx.misc
        dev = dout.device
        cin, C = mod.cin, mod.cout
        edges = B * O * K
        dout = dout.contiguous()
        grads = [None] * len(params)
        with torch.cuda.device(dev):
            dF = torch.empty_like(F)
            dA = torch.empty_like(A) if A is not None else None
            _lib.check(L.gridgcn_train_pool_bwd(dout.data_ptr() + 16, 4 + C, F.data_ptr(), A.data_ptr() if A is not None else None,
                                                argmax.data_ptr(), centmsk.data_ptr(), B * O, K, C, dF.data_ptr(),
                                                dA.data_ptr() if dA is not None else None, _stream(dout)),
                       "gridgcn_train_pool_bwd")

            wps = [mod._pad_w(i, params[4 * i]) for i in range(len(saved))]   # padded (Cout, Cin) weights
            need = 0   # ONE fill for every stage's sums and weight / bias gradients
            for i, wp_ in enumerate(wps):
                need += 3 * wp_.shape[0] + wp_.numel()
            gpool = torch.zeros(need, dtype=torch.float32, device=dev)
            gpool_off = [0]

            def take(n):
                v = gpool[gpool_off[0]:gpool_off[0] + n]
                gpool_off[0] += n
                return v

            def stage_bwd(dy, idx, need_dx):
                x, z, y, mean, invstd = saved[idx]
                w, b, gamma, beta = params[4 * idx:4 * idx + 4]
                rows, Cout = z.shape
                # dy <- dz = dy * (y > 0);  z <- xhat;  per-channel sum dz, sum dz * xhat in the same pass
                sum_dz, sum_dzx = take(Cout), take(Cout)
                _lib.check(L.gridgcn_train_relu_bwd_xhat(dy.data_ptr(), y.data_ptr(), z.data_ptr(), rows, Cout, mean.data_ptr(),
                                                         invstd.data_ptr(), sum_dz.data_ptr(), sum_dzx.data_ptr(), _stream(dy)),
                           "gridgcn_train_relu_bwd_xhat")
                dzpre = torch.empty_like(dy)
                _lib.check(L.gridgcn_train_bn_bwd(dy.data_ptr(), z.data_ptr(), rows, Cout, gamma.data_ptr(), invstd.data_ptr(),
                                                  sum_dz.data_ptr(), sum_dzx.data_ptr(), dzpre.data_ptr(), _stream(dy)),
                           "gridgcn_train_bn_bwd")
                wp = wps[idx]
                dwp, dbias = take(wp.numel()).view_as(wp), take(Cout)   # dW | db
                _lib.check(L.gridgcn_train_wgrad(dzpre.data_ptr(), Cout, x.data_ptr(), x.stride(0), x.shape[1], None, 0, 0, rows,
                                                 dwp.data_ptr(), dbias.data_ptr(), _stream(dy)), "gridgcn_train_wgrad")
                grads[4 * idx] = mod._unpad_w(idx, dwp).reshape(w.shape)
                grads[4 * idx + 1] = dbias
                grads[4 * idx + 2] = sum_dzx   # d gamma
                grads[4 * idx + 3] = sum_dz    # d beta
                if not need_dx:
                    return None
                return _gemm(dzpre, wp.t().contiguous(), _zeros(wp.shape[1], dev), False)   # dX = dZ W

            if A is not None:
                d = dA
                for s in range(mod.n_att - 1, -1, -1):
                    d = stage_bwd(d, mod.n_feat + s, need_dx=s > 0)
            d = dF
            for s in range(mod.n_feat - 1, -1, -1):
                d = stage_bwd(d, s, need_dx=(s > 0 or cin > 0))
            dtable = None
            if cin > 0 and ctx.needs_input_grad[1]:
                dtable = torch.zeros((B, Nprev, roww), dtype=torch.float32, device=dev)
                _lib.check(L.gridgcn_train_scatter_add(d.data_ptr(), rowidx.data_ptr(), edges, cin, roww, dtable.data_ptr(),
                                                       _stream(d)), "gridgcn_train_scatter_add")
        return (None, dtable, None, None, None) + tuple(grads)


class GridConvTrainCuda(GridConvTrain):
    """Drop-in for ``train.GridConvTrain`` (same constructor, parameters, state dict and ``export_layer``): forward and
    backward on the library's CUDA kernels.  In eval mode (``module.eval()``) it defers to the parent's op-by-op
    forward with the moving statistics."""

    def __init__(self, layer, pre_relu=True, bn_decay=0.9):
        super().__init__(layer, pre_relu, bn_decay)
        if self.localfdim or self.att_full or len(self.att) not in (0, 2):
            raise NotImplementedError("GridConvTrainCuda implements the segmentation flavour of the block")
        self.n_feat, self.n_att = len(self.feat), len(self.att)
        self.cout = int(self.feat[-1].weight.shape[0])
        aw = 0 if self.attfdim <= 0 else (3 if self.attfdim <= 3 else (4 if self.attfdim < 10 else 10))
        self.ain, self.ain_p = aw, (aw + 3) // 4 * 4
        self.bn_decay = bn_decay

    def _stages(self):
        return list(self.feat) + list(self.att)

    def _pad_w(self, idx, w):
        """(Cout, Cin) -> (Cout, Cin padded): [geo, 0] for the first feature stage without input features, trailing
        zeros of att_vec for the first attention stage."""
        w = w.reshape(w.shape[0], -1)
        want = None
        if idx == 0 and self.cin == 0:
            want = 4
        elif idx == self.n_feat and self.n_att:
            want = self.ain_p
        if want is None or w.shape[1] == want:
            return w.contiguous()
        return torch.cat([w, w.new_zeros(w.shape[0], want - w.shape[1])], 1).contiguous()

    def _unpad_w(self, idx, dwp):
        st = self._stages()[idx]
        return dwp[:, :st.weight.reshape(st.weight.shape[0], -1).shape[1]].contiguous()

    def forward(self, table, nebidx, cent, centmsk):
        if not self.training or not table.is_cuda:
            return super().forward(table, nebidx, cent, centmsk)
        params = []
        for st in self._stages():
            params += [st.weight, st.bias, st.bn.weight, st.bn.bias]
        return _Fn.apply(self, table, nebidx.int(), cent, centmsk, *params)
