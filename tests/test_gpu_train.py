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
