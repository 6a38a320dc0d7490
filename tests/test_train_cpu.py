"""CPU suite, part 5: the training-mode GridConv block (grid-gcn_b200/train.py) -- eval-mode forward against the
oracle, batch-statistic BatchNorm, parameter export to the fused kernels, and the data-parallel step (one flat
gradient all-reduce) on two gloo ranks.  Indices come from the CPU oracle here (test infrastructure); on a GPU
the same module takes them from the CUDA operators (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridgcn_b200 import gridconv, stack, synth, train


def _oracle_query(cfg):
    from oracle import oracle

    def q(loc, num, l):
        fn = oracle.gridify_knn if cfg.query == "gridifyknn" else oracle.gridify
        out = fn(loc.detach().numpy(), num.numpy(), max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid,
                 kernel_size=l.kernel_size, loc=cfg.loc, coord_shift=cfg.coord_shift,
                 voxel_size=(l.voxel_size,) * 3, grid_size=(l.grid_size,) * 3)
        return tuple(torch.from_numpy(o) for o in out)
    return q


def _case(cin, pt, **kw):
    from oracle import oracle
    rng = np.random.default_rng(3)
    data, npts = synth.make_batch(2, 256, seed0=5, voxels=(0.25,))
    nebidx, _, cent, centmsk, _ = oracle.gridify_knn(
        data, npts, max_p_grid=8, max_o_grid=32, kernel_size=3, loc=1, coord_shift=(1, 1, 1),
        voxel_size=(0.25,) * 3, grid_size=(8,) * 3)
    table = data if cin == 0 else np.concatenate([data, rng.uniform(0, 1, (2, 256, cin)).astype(np.float32)], 2)
    layer = gridconv.init_layer(np.random.default_rng(7), cin, pt, **kw)
    return layer, table, nebidx, cent, centmsk


def test_eval_forward_matches_oracle_and_export_round_trips():
    from oracle import gridconv_oracle
    for cin, pt, kw in ((0, [16, 32], dict(attfdim=10)), (16, [16, 32], dict(attfdim=10)),
                        (16, [16, 32], dict(attfdim=4, att_ele_lst=[8, 16, 32], att_full="next", localfdim=3))):
        layer, table, nebidx, cent, centmsk = _case(cin, pt, **kw)
        m = train.GridConvTrain(layer).eval()
        args = [torch.from_numpy(a) for a in (table, nebidx, cent, centmsk)]
        got = m(*args).detach().numpy()
        want = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, layer)
        assert np.allclose(got, want, rtol=1e-4, atol=1e-5)
        # a few training-mode forwards move the BatchNorm statistics; the exported layer reproduces the
        # module's new eval-mode forward through the oracle (the path a trained model takes to the fused kernels)
        m.train()
        for _ in range(3):
            m(*args)
        m.eval()
        exported = m.export_layer()
        assert not np.allclose(exported["feat"][0]["moving_mean"], layer["feat"][0]["moving_mean"])
        want2 = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, exported)
        assert np.allclose(m(*args).detach().numpy(), want2, rtol=1e-4, atol=1e-5)


def test_training_mode_uses_batch_statistics():
    layer, table, nebidx, cent, centmsk = _case(0, [16, 32], attfdim=10)
    st = train.ConvBnRelu(layer["feat"][0], bn_decay=0.9).train()
    x = torch.randn(2, 3, 5, 7)
    y = st(x)
    z = torch.einsum("oc,bcnp->bonp", st.weight, x) + st.bias[None, :, None, None]
    mu, var = z.mean((0, 2, 3)), z.var((0, 2, 3), unbiased=False)
    ref = torch.relu((z - mu[None, :, None, None]) / torch.sqrt(var + 1e-3)[None, :, None, None]
                     * st.bn.weight[None, :, None, None] + st.bn.bias[None, :, None, None])
    assert torch.allclose(y, ref, atol=1e-5)
    # moving = moving * bn_decay + batch * (1 - bn_decay)   (MXNet momentum semantics, utils/ops.py:152)
    want_mean = torch.as_tensor(layer["feat"][0]["moving_mean"]) * 0.9 + mu.detach() * 0.1
    assert torch.allclose(st.bn.running_mean, want_mean, atol=1e-6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_model(cfg):
    torch.manual_seed(0)
    return train.GridGcnClassifier(cfg, stack.init_params(cfg, seed=2), num_classes=5, query=_oracle_query(cfg))


def _shard(rank):
    data, npts = synth.make_batch(2, 256, seed0=100 + 2 * rank, voxels=(0.25, 0.5))
    labels = torch.tensor([rank, (rank + 2) % 5])
    return torch.from_numpy(data), torch.from_numpy(npts), labels


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = stack.tiny(8)
        model = _make_model(cfg)
        opt = torch.optim.SGD(model.parameters(), lr=0.1)
        loss = train.train_step(model, opt, *_shard(rank))
        out[rank] = (loss, [p.detach().numpy().copy() for p in model.parameters()])
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_two_gloo_ranks():
    """Each rank steps on its own clouds; the ONE flat all-reduce makes both apply the mean gradient: parameters
    end identical on both ranks and equal to p - lr * (g_0 + g_1) / 2 with g_r computed in this process."""
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
    torch.set_num_threads(1)
    cfg = stack.tiny(8)
    grads, p0 = [], None
    for rank in range(world):
        model = _make_model(cfg).train()
        p0 = [p.detach().clone() for p in model.parameters()]
        d, n, y = _shard(rank)
        torch.nn.functional.cross_entropy(model(d, n), y).backward()
        grads.append([p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in model.parameters()])
    for k, p in enumerate(p0):
        want = (p - 0.1 * (grads[0][k] + grads[1][k]) / 2).numpy()
        assert np.allclose(res[0][1][k], want, atol=1e-6), k
        assert np.array_equal(res[0][1][k], res[1][1][k]), k
    assert res[0][0] != res[1][0]  # different shards, different local losses


def test_cuda_block_host_logic_without_gpu():
    """train_cuda.GridConvTrainCuda: the host side that needs no GPU -- padded weight layouts ([geo, 0] for the first
    feature stage, trailing zeros of att_vec) round-trip, the classification flavours are refused, and on CPU tensors
    the module is its parent (same parameters, same forward), so a model built with block="cuda" still runs the CPU
    tests' op-by-op path."""
    import pytest
    from gridgcn_b200 import train_cuda
    cfg = stack.tiny(8)
    params = stack.init_params(cfg, seed=4)
    m = train_cuda.GridConvTrainCuda(params[0])       # first layer: no input features
    w0 = m.feat[0].weight
    wp = m._pad_w(0, w0)
    assert wp.shape == (w0.shape[0], 4) and torch.equal(wp[:, :3], w0) and float(wp[:, 3].abs().max()) == 0.0
    assert torch.equal(m._unpad_w(0, wp), w0)
    ia = m.n_feat                                      # first attention stage: att_vec padded to a multiple of 4
    wa = m._pad_w(ia, m.att[0].weight)
    assert wa.shape[1] == m.ain_p and m.ain_p % 4 == 0
    assert torch.equal(m._unpad_w(ia, wa), m.att[0].weight.reshape(wa.shape[0], -1))
    m1 = train_cuda.GridConvTrainCuda(params[1])       # layer with input features: nothing to pad in the feature chain
    assert m1._pad_w(0, m1.feat[0].weight).shape == m1.feat[0].weight.reshape(m1.feat[0].weight.shape[0], -1).shape
    with pytest.raises(NotImplementedError):
        train_cuda.GridConvTrainCuda(dict(params[1], localfdim=3))
    # CPU tensors: the parent's forward, bit for bit
    ref = train.GridConvTrain(params[0])
    ref.load_state_dict(m.state_dict())
    B, Np, O, K = 2, 32, 8, 4
    g = torch.Generator().manual_seed(0)
    table = torch.rand((B, Np, 4), generator=g)
    nebidx = torch.randint(0, Np, (B, O, K), generator=g, dtype=torch.int32)
    cent = torch.rand((B, O, 4), generator=g)
    msk = torch.ones((B, O))
    m.train(), ref.train()
    assert torch.equal(m(table, nebidx, cent, msk), ref(table, nebidx, cent, msk))
