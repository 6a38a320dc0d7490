"""GPU suite: the training-mode block (grid-gcn_b200/train.py) on the device -- its eval-mode forward against the
fused inference kernels, a few data-parallel-style steps with the CUDA index operators in front, and the trained
parameters handed back to the fused kernels."""
import numpy as np
import pytest
import torch

from gridgcn_b200 import gridconv, stack, synth, train

pytestmark = pytest.mark.gpu


def _rel_err(got, want):
    rms = float(np.sqrt(np.mean(want.astype(np.float64) ** 2))) + 1e-30
    return float(np.max(np.abs(got - want) / (np.abs(want) + rms)))


def test_train_module_eval_matches_fused_kernels(gg, cuda_dev):
    cfg = stack.tiny(8)
    params = stack.init_params(cfg, seed=6)
    data, npts = synth.make_batch(3, cfg.num_points, seed0=30, voxels=cfg.voxels)
    d, n = torch.from_numpy(data).to(cuda_dev), torch.from_numpy(npts).to(cuda_dev)
    enc = stack.GridGcnEncoder(cfg, params, cuda_dev, precision="tf32x3")
    enc(d, n, keep_trace=True)
    table = d
    for i, p in enumerate(params):
        tr = enc.trace[i]
        m = train.GridConvTrain(p).to(cuda_dev).eval()
        with torch.no_grad():
            got = m(table, tr["nebidx"], tr["cent"], tr["centmsk"])
        err = _rel_err(got[..., 4:].cpu().numpy(), tr["table"][..., 4:].cpu().numpy())
        assert err <= 1e-3, "layer %d: rel err %.3g" % (i, err)
        table = tr["table"]


def test_train_steps_reduce_loss_and_export_to_fused_kernel(gg, cuda_dev):
    torch.manual_seed(0)
    cfg = stack.tiny(8)
    model = train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=2), num_classes=4).to(cuda_dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    data, npts = synth.make_batch(8, cfg.num_points, seed0=40, voxels=cfg.voxels)
    d, n = torch.from_numpy(data).to(cuda_dev), torch.from_numpy(npts).to(cuda_dev)
    labels = torch.arange(8, device=cuda_dev) % 4
    losses = [train.train_step(model, opt, d, n, labels) for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses[::5]
    # trained layer 0 -> fused inference kernel (BatchNorm folded) == the module's eval-mode forward
    model.eval()
    l0 = cfg.layers[0]
    nebidx, _, cent, centmsk, _ = gg.GridifyKNN(d, n, max_o_grid=l0.max_o_grid, max_p_grid=l0.max_p_grid,
                                                kernel_size=l0.kernel_size, stride=1, coord_shift=cfg.coord_shift,
                                                voxel_size=[l0.voxel_size] * 3, grid_size=[l0.grid_size] * 3,
                                                loc=cfg.loc)
    with torch.no_grad():
        want = model.layers[0](d, nebidx, cent, centmsk)
    fused = gridconv.GridConv(model.layers[0].export_layer(), cuda_dev, pre_relu=cfg.pre_relu, precision="tf32x3")
    got = fused(d, nebidx, cent, centmsk)
    err = _rel_err(got[..., 4:].cpu().numpy(), want[..., 4:].cpu().numpy())
    assert err <= 1e-3, "exported layer: rel err %.3g" % err


@pytest.mark.parametrize("cin,mlp,K", [(0, [16, 32], 8), (16, [32, 64], 8), (64, [64, 64, 128], 32)],
                         ids=["first_layer_geo", "with_features", "seg_layer1_shape"])
def test_training_kernels_match_autograd(gg, cuda_dev, cin, mlp, K):
    """The hand-written training kernels (csrc/train_ops.cu + the tcgen05 row GEMM; train_cuda.GridConvTrainCuda):
    forward with batch-statistic BatchNorm, and the backward of gather / MLP / attention product / max pool -- against
    the same block on PyTorch ops + autograd (train.GridConvTrain), outputs, every parameter gradient, the gradient
    of the input features and the updated moving statistics within 1e-3."""
    from gridgcn_b200 import train_cuda
    rng = np.random.default_rng(cin + K)
    B, N, O = 3, 96, 24
    layer = gridconv.init_layer(np.random.default_rng(9), cin, mlp, 10)
    table = torch.from_numpy(rng.uniform(-1, 1, size=(B, N, 4 + cin)).astype(np.float32)).to(cuda_dev)
    nebidx = torch.from_numpy(rng.integers(-1, N, size=(B, O, K)).astype(np.int32)).to(cuda_dev)  # -1: BallKNN miss
    cent = torch.from_numpy(rng.uniform(-1, 1, size=(B, O, 4)).astype(np.float32)).to(cuda_dev)
    centmsk = torch.from_numpy((rng.uniform(size=(B, O)) > 0.2).astype(np.float32)).to(cuda_dev)
    wout = torch.from_numpy(rng.normal(size=(B, O, mlp[-1])).astype(np.float32)).to(cuda_dev)

    def run(mod):
        mod = mod.to(cuda_dev).train()
        t = table.clone().requires_grad_(cin > 0)
        out = mod(t, nebidx, cent, centmsk)
        loss = (out[..., 4:] * wout).sum()
        loss.backward()
        grads = {n: p.grad.detach().cpu().numpy() for n, p in mod.named_parameters()}
        stats = {n: b.detach().cpu().numpy() for n, b in mod.named_buffers() if "running" in n}
        return out.detach().cpu().numpy(), grads, stats, (t.grad.detach().cpu().numpy() if cin > 0 else None)

    want = run(train.GridConvTrain(layer))
    got = run(train_cuda.GridConvTrainCuda(layer))
    assert np.array_equal(got[0][..., :4], want[0][..., :4])
    assert _rel_err(got[0][..., 4:], want[0][..., 4:]) <= 1e-3
    for name in want[1]:
        if name.endswith(".bias") and ".bn." not in name:
            # a convolution bias in front of a batch-statistic BatchNorm has an analytically ZERO gradient (the mean
            # is subtracted): both implementations return rounding noise -- require it to be noise, not to agree
            scale = float(np.abs(want[1][name.replace(".bias", ".bn.bias")]).max())
            assert np.abs(got[1][name]).max() <= 1e-3 * scale and np.abs(want[1][name]).max() <= 1e-3 * scale, name
            continue
        assert _rel_err(got[1][name], want[1][name]) <= 1e-3, name
    for name in want[2]:
        assert _rel_err(got[2][name], want[2][name]) <= 1e-3, name
    if cin > 0:
        assert np.array_equal(got[3][..., :4], np.zeros_like(got[3][..., :4]))  # coordinates carry no gradient
        assert _rel_err(got[3][..., 4:], want[3][..., 4:]) <= 1e-3


def test_training_step_on_cuda_kernels(gg, cuda_dev):
    """A few optimiser steps of the classifier ladder with the CUDA-kernel block: the loss falls, and the trained
    layer exported to the fused inference kernel reproduces the module's eval-mode forward."""
    torch.manual_seed(0)
    cfg = stack.tiny(8)
    model = train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=2), num_classes=4, block="cuda").to(cuda_dev)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    data, npts = synth.make_batch(8, cfg.num_points, seed0=40, voxels=cfg.voxels)
    d, n = torch.from_numpy(data).to(cuda_dev), torch.from_numpy(npts).to(cuda_dev)
    labels = torch.arange(8, device=cuda_dev) % 4
    losses = [train.train_step(model, opt, d, n, labels) for _ in range(30)]
    assert losses[-1] < 0.5 * losses[0], losses[::5]


@pytest.mark.parametrize("block,optim", [("cuda", "sgd"), ("torch", "sgd"), ("cuda", "adam")])
def test_graphed_train_step_matches_eager(gg, cuda_dev, block, optim):
    """train.GraphedTrainStep (forward + backward and the update replayed as CUDA graphs, gradients in one flat bucket)
    is the same arithmetic as train.train_step: after 5 SGD steps from the same start the parameters, the BatchNorm
    moving statistics and the losses agree (1e-4: the weight-gradient kernels sum with atomics), and constructing the
    graphed step does not move the parameters."""
    cfg = stack.tiny(8)
    data, npts = synth.make_batch(8, cfg.num_points, seed0=40, voxels=cfg.voxels)
    d, n = torch.from_numpy(data).to(cuda_dev), torch.from_numpy(npts).to(cuda_dev)
    labels = torch.arange(8, device=cuda_dev) % 4

    def make():
        torch.manual_seed(0)
        m = train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=2), num_classes=4, block=block).to(cuda_dev)
        if optim == "adam":  # stateful: the state created by the warm-up steps must be back at its initial value
            return m, torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
        return m, torch.optim.SGD(m.parameters(), lr=1e-2)

    m0, o0 = make()
    eager = [train.train_step(m0, o0, d, n, labels) for _ in range(5)]
    m1, o1 = make()
    before = {k: v.clone() for k, v in m1.state_dict().items()}
    step = train.GraphedTrainStep(m1, o1, d, n, labels)
    for k, v in m1.state_dict().items():
        assert torch.equal(v, before[k]), "construction changed " + k
    for st in o1.state.values():  # state the warm-up steps created is back at its initial value (zeros)
        for k, v in st.items():
            assert not torch.is_tensor(v) or float(v.abs().max()) == 0.0, k
    graphed = [float(step(d, n, labels)) for _ in range(5)]
    tol = 1e-4 if optim == "sgd" else 2e-3  # Adam divides by sqrt(v): the atomics' rounding noise is amplified where g ~ 0
    assert np.allclose(graphed, eager, rtol=tol, atol=1e-6), (graphed, eager)
    s0, s1 = m0.state_dict(), m1.state_dict()
    if optim == "adam":
        return  # Adam turns the rounding noise of near-zero gradients (conv biases in front of a batch-stat BN, ...) into +-lr steps
    for k in s0:
        a, b = s0[k].float().cpu().numpy(), s1[k].float().cpu().numpy()
        assert _rel_err(b, a) <= tol, k
