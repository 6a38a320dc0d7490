#!/usr/bin/env python
"""bench.py -- points/sec through the 4-layer GridConv stack @ N=8192, K=64 (BASELINE.json metric).

A "step" is one pass of the hot path (per layer: GridifyKNN voxel-hash build + query, fused
GridConv) over one batch of synthetic 8192-point clouds.  One process per GPU; clouds are
independent, so ranks shard the batch with no data-path collective ("scaling": "weak").

  python bench.py --gpus 1 --steps 20 --warmup 5            # this framework (CUDA, sm_100a)
  python bench.py --impl reference ...                      # CPU arm: the oracle port of the
                                                            # reference's algorithm on host cores

Prints ONE JSON line on rank 0 (see the task contract): value = whole-job points/sec with inputs
resident in HBM; e2e = the same through the public API with pinned HOST buffers (H2D + D2H inside
the timed region); roofline = dominant kernel vs MEASURED_PEAKS.json; roofline_hbm = the
gridifyknn operator vs the measured HBM copy bandwidth; cpu_baseline = the oracle on host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


import contextlib


@contextlib.contextmanager
def stdout_to_stderr():
    """File-descriptor level redirect: NCCL prints its version banner on stdout while it initialises, and
    stdout is reserved for the ONE JSON line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            if "hbm_gbs" in p and "bf16_tflops" in p:
                p["_source"] = "measured"
                return p
        except Exception:
            pass
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback"
    return p


def gridify_bytes(N, O, P):
    """Algorithmic HBM bytes of one Gridify/GridifyKNN call per cloud:
    read 16N, write nebidx + mask 8*O*P, cent 16*O, centmsk 4*O, centnum 4."""
    return 16 * N + 8 * O * P + 16 * O + 4 * O + 4


def layer_macs(params):
    return [sum(st["weight"].size for st in p["feat"]) + sum(st["weight"].size for st in p["att"])
            for p in params]


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons while the timed region runs.  In-process NVML (pynvml) when it is there:
    a query costs microseconds.  Spawning nvidia-smi every few milliseconds instead -- the fallback -- attaches a new
    process to the driver each time and was seen to stall the GPU for milliseconds inside a 66 ms timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]  # nvmlClocksEventReason HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap

    def __init__(self, index, uuid=None):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml, self.handle, self.source = None, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)  # probe
            self.nvml, self.handle, self.source = pynvml, h, "nvml"
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(h))
        return [str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for b in self.BITS]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True,
                                         text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.25)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(s[2 + i].lower() == "active" for s in self.samples)]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples), "source": self.source}


def cpu_stack(oracle, gridconv_oracle, cfg, params, data, npts, pool=None):
    """The reference algorithm's CPU port over a batch: oracle query (C, OpenMP over clouds) + numpy
    GridConv per layer (one cloud per worker thread; numpy releases the GIL)."""
    if cfg.query.startswith("occaware"):
        q = lambda *a, **kw: oracle.gridify_occaware(*a, seed=cfg.cas_seed, knn_query=cfg.query.endswith("knn"), **kw)
    else:
        q = oracle.gridify_knn if cfg.query == "gridifyknn" else oracle.gridify
    table, loc, num = data, data, npts
    for l, p in zip(cfg.layers, params):
        nebidx, _, cent, centmsk, num = q(loc, num, max_p_grid=l.max_p_grid, max_o_grid=l.max_o_grid,
                                          kernel_size=l.kernel_size, loc=cfg.loc,
                                          coord_shift=cfg.coord_shift, voxel_size=(l.voxel_size,) * 3,
                                          grid_size=(l.grid_size,) * 3)
        if pool is None:
            table = gridconv_oracle.gridconv_layer(table, nebidx, cent, centmsk, p, pre_relu=cfg.pre_relu)
        else:
            def one(b, table=table, nebidx=nebidx, cent=cent, centmsk=centmsk, p=p):
                return gridconv_oracle.gridconv_layer(table[b:b + 1], nebidx[b:b + 1], cent[b:b + 1],
                                                      centmsk[b:b + 1], p, pre_relu=cfg.pre_relu)
            table = np.concatenate(list(pool.map(one, range(len(data)))), axis=0)
        loc = cent
    return table


def time_cpu(cfg, params, clouds, steps, warmup):
    """Times the CPU port on `clouds` clouds per step with every host thread it can use: OpenMP over
    clouds in the C oracle, a thread per cloud (BLAS pinned to 1 thread each) for the numpy GridConv."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle, gridconv_oracle
    from gridgcn_b200 import synth
    oracle.build()
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, clouds))
    oracle.set_threads(min(oracle.max_threads(), workers))
    data, npts = synth.make_batch(min(clouds, 8), cfg.num_points, seed0=0, voxels=cfg.voxels)
    reps = (clouds + len(data) - 1) // len(data)
    data = np.tile(data, (reps, 1, 1))[:clouds].copy()
    npts = np.full((clouds, 1), cfg.num_points, np.int32)
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)
    except Exception:
        limiter = None
    with ThreadPoolExecutor(max_workers=workers) as pool:
        for _ in range(warmup):
            cpu_stack(oracle, gridconv_oracle, cfg, params, data, npts, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_stack(oracle, gridconv_oracle, cfg, params, data, npts, pool)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    if limiter is not None:
        limiter.unregister() if hasattr(limiter, "unregister") else None
    return clouds * cfg.num_points / dt, dt, workers


def _timed(torch, flush, fn, steps):
    """Mean device time (ms) of fn over `steps` calls, L2 flushed before each, CUDA events on the current stream."""
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    return sum(s.elapsed_time(e) for s, e in evs) / max(steps, 1)


def measure_encoder(torch, flush, peaks, cfg, B, precision, steps=5, graph=False, seed=0):
    """One BASELINE configuration measured the way the headline is: device-resident points/s of the encoder
    ladder, per-operator times (query = voxel-hash build + neighbour gather; gridconv = the per-edge MLP
    kernels) and the two roofline fractions."""
    from gridgcn_b200 import stack, synth
    params = stack.init_params(cfg, seed=seed)
    dev = flush.device
    pool = min(B, 8)
    base, _ = synth.make_batch(pool, cfg.num_points, seed0=1000, voxels=cfg.voxels)
    data = torch.from_numpy(np.tile(base, ((B + pool - 1) // pool, 1, 1))[:B].copy()).to(dev)
    npts = torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
    enc = stack.GridGcnEncoder(cfg, params, dev, precision=precision)
    for _ in range(3):
        enc(data, npts)
    out = {"clouds": B, "points_per_cloud": cfg.num_points, "layers": len(cfg.layers),
           "K": [l.max_p_grid for l in cfg.layers], "O": [l.max_o_grid for l in cfg.layers],
           "query": cfg.query, "precision": precision}
    ms = _timed(torch, flush, lambda: enc(data, npts), steps)
    out["ms_per_step"] = ms
    out["points_per_s"] = B * cfg.num_points / (ms * 1e-3)
    if graph:
        replay = stack.capture_graph(enc, data, npts)
        replay()
        gms = _timed(torch, flush, replay, steps)
        out["cuda_graph"] = {"ms_per_step": gms, "points_per_s": B * cfg.num_points / (gms * 1e-3)}
    enc(data, npts, keep_trace=True)
    tr = enc.trace
    qf = stack.query_fn(cfg)
    q_ms, c_ms = [], []
    for i, (l, conv) in enumerate(zip(cfg.layers, enc.convs)):
        loc_in = data if i == 0 else tr[i - 1]["cent"]
        num_in = npts if i == 0 else tr[i - 1]["actual_centnum"]
        tin = data if i == 0 else tr[i - 1]["table"]
        kwl = dict(max_o_grid=l.max_o_grid, max_p_grid=l.max_p_grid, kernel_size=l.kernel_size, stride=1,
                   coord_shift=cfg.coord_shift, voxel_size=[l.voxel_size] * 3, grid_size=[l.grid_size] * 3, loc=cfg.loc)
        fq = lambda: qf(loc_in, num_in, **kwl)  # noqa: E731
        fc = lambda: conv(tin, tr[i]["nebidx"], tr[i]["cent"], tr[i]["centmsk"])  # noqa: E731
        for _ in range(2):
            fq(); fc()
        q_ms.append(_timed(torch, flush, fq, steps))
        c_ms.append(_timed(torch, flush, fc, steps))
    macs = layer_macs(params)
    flops = [2 * l.max_o_grid * l.max_p_grid * m for l, m in zip(cfg.layers, macs)]
    l0 = cfg.layers[0]
    q_gbs = B * gridify_bytes(cfg.num_points, l0.max_o_grid, l0.max_p_grid) / (q_ms[0] * 1e-3) / 1e9
    dom = int(np.argmax(c_ms))
    tf = B * flops[dom] / (c_ms[dom] * 1e-3) / 1e12
    tensor_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    out["breakdown_ms"] = {"gather": q_ms, "mlp": c_ms, "gather_total": sum(q_ms), "mlp_total": sum(c_ms),
                           "gather_share": sum(q_ms) / (sum(q_ms) + sum(c_ms))}
    out["roofline_hbm"] = {"kernel": "%s layer 0 (build + query)" % cfg.query, "achieved_gbs": q_gbs,
                           "frac": q_gbs / peaks["hbm_gbs"]}
    out["roofline_tensor"] = {"kernel": "gridconv layer %d" % dom, "algorithmic_tflops": tf, "frac": tf / tensor_peak}
    return out


def measure_seg_graph(torch, flush, B, precision, steps=5):
    """BASELINE config 3: the shipped ScanNet-8192 graph (3 Gridify + GridConv encoder layers, 3 BallKNN + GridConv
    decoder layers, head) through stack.GridGcnSeg at the shipped batch."""
    from gridgcn_b200 import stack, synth
    cfg, up = stack.seg8192_shipped(), stack.UpCfg()
    params = stack.init_seg_params(cfg, up, seed=0)
    dev = flush.device
    pool = min(B, 8)
    base, _ = synth.make_batch(pool, cfg.num_points, seed0=2000, voxels=cfg.voxels)
    data = torch.from_numpy(np.tile(base, ((B + pool - 1) // pool, 1, 1))[:B].copy()).to(dev)
    npts = torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
    net = stack.GridGcnSeg(cfg, up, params, dev, precision=precision)
    for _ in range(3):
        net(data, npts)
    ms = _timed(torch, flush, lambda: net(data, npts), steps)
    enc_ms = _timed(torch, flush, lambda: net.enc(data, npts), steps)
    replay = stack.capture_graph(net, data, npts)
    replay()
    gms = _timed(torch, flush, replay, steps)
    return {"clouds": B, "points_per_cloud": cfg.num_points, "graph": "encoder (Gridify, strict reservoir) + decoder "
            "(BallKNN k=5) + head, 21 classes", "precision": precision, "ms_per_step": ms,
            "points_per_s": B * cfg.num_points / (ms * 1e-3), "encoder_ms": enc_ms, "decoder_head_ms": ms - enc_ms,
            "cuda_graph": {"ms_per_step": gms, "points_per_s": B * cfg.num_points / (gms * 1e-3)}}


def measure_train_step(torch, dist, dev, world, B=3, steps=4, block="cuda"):
    """BASELINE config 4: one data-parallel TRAINING step of the seg81920 ladder, B clouds per GPU: forward
    (index operators = this library's kernels; training-mode block on torch ops, train.py), backward, ONE flat
    NCCL all-reduce of the gradient bucket, update.  Every rank takes part; times are MAX over ranks."""
    from gridgcn_b200 import stack, synth, train, shard
    cfg = stack.seg81920_shipped("gridify")
    params = stack.init_params(cfg, seed=0)
    model = train.GridGcnClassifier(cfg, params, num_classes=21, block=block).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=1e-3)
    base, _ = synth.make_batch(1, cfg.num_points, seed0=3000 + (dist.get_rank() if world > 1 else 0), voxels=cfg.voxels)
    data = torch.from_numpy(np.tile(base, (B, 1, 1)).copy()).to(dev)
    npts = torch.full((B, 1), cfg.num_points, dtype=torch.int32, device=dev)
    labels = torch.arange(B, device=dev) % 21
    plist = [p for p in model.parameters()]
    grad_bytes = sum(p.numel() for p in plist) * 4
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_step, t_ar = [], []
    for it in range(steps + 2):
        model.train()
        opt.zero_grad(set_to_none=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        loss = torch.nn.functional.cross_entropy(model(data, npts), labels)
        loss.backward()
        e1.record()
        shard.allreduce_gradients(plist)
        e2.record()
        opt.step()
        e3.record()
        torch.cuda.synchronize()
        if it >= 2:
            t_step.append(e0.elapsed_time(e3))
            t_ar.append(e1.elapsed_time(e2))
    eager_ms = shard.max_over_ranks([sum(t_step) / len(t_step)], device=dev)[0]
    # the same step replayed as CUDA graphs (train.GraphedTrainStep): the eager step is launch bound
    del loss  # its autograd graph pins AccumulateGrad nodes to the eager stream
    gstep = train.GraphedTrainStep(model, opt, data, npts, labels)
    t_step, t_ar = [], []
    for it in range(4 * steps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        gstep.g1.replay()
        e1.record()
        gstep._reduce()
        e2.record()
        gstep.g2.replay()
        e3.record()
        torch.cuda.synchronize()
        if it >= 2:
            t_step.append(e0.elapsed_time(e3))
            t_ar.append(e1.elapsed_time(e2))
    step_ms, ar_ms = shard.max_over_ranks([sum(t_step) / len(t_step), sum(t_ar) / len(t_ar)], device=dev)
    return {"workload": "seg81920 ladder, %d clouds per GPU, training step (forward+backward and the update replayed as CUDA graphs, "
                        "all-reduce between them; eager_step_ms = the same step launched op by op); block = %s" % (B, "hand-written training kernels "
                        "(csrc/train_ops.cu + tcgen05 row GEMM)" if block == "cuda" else "torch ops + autograd"), "n_gpus": world, "step_ms": step_ms, "cuda_graph": True,
            "eager_step_ms": eager_ms, "allreduce_ms": ar_ms,
            "allreduce_share": ar_ms / step_ms, "gradient_bytes": grad_bytes,
            "points_per_s": world * B * cfg.num_points / (step_ms * 1e-3),
            "collective": "one flat NCCL all-reduce (SUM, then / world)" if world > 1 else "none (single rank)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=296,
                    help="clouds per GPU per step (default: two per SM of a 148-SM B200; measured 192 -> 4.19e8, "
                         "296 -> 4.41e8, 444 -> 4.45e8, 592 -> 4.47e8 points/s)")
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--query", default="gridifyknn", choices=["gridifyknn", "gridify", "occaware", "occaware_knn"],
                    help="centre sampling + neighbour query operator (occaware = coverage-aware sampling)")
    ap.add_argument("--workload", default="seg8192", choices=["seg8192", "cls1024", "seg81920", "cls1024_shipped"],
                    help="seg8192: BASELINE.json's metric configuration (N=8192, 4 layers); cls1024: N=1024 4-layer "
                         "ladder (config 2); seg81920: the shipped 81920-point ladder (config 4, 3 layers, P0=128); "
                         "cls1024_shipped: the shipped ModelNet40 ladder with the classification flavour of the block "
                         "(fp32 CUDA-core kernel, Gridify query)")
    ap.add_argument("--precision", default=os.environ.get("GRIDGCN_PRECISION", "tf32x3"),
                    choices=["tf32x3", "tf32", "fp32"])
    ap.add_argument("--graph", action="store_true",
                    help="replay the forward as ONE CUDA graph (stack.capture_graph) for the device-resident `value`; "
                         "pays off when the step is launch bound (small batches)")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the `configs` block (the other BASELINE configurations measured in the same run)")
    ap.add_argument("--cpu-clouds", type=int, default=0,
                    help="clouds per CPU-baseline step (default: one per host core, at most 64)")
    args = ap.parse_args()
    if args.cpu_clouds <= 0:
        args.cpu_clouds = max(1, min(64, os.cpu_count() or 1))

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from gridgcn_b200 import stack
    if args.workload == "seg8192":
        cfg = stack.seg8192_4layer(args.K, args.query)
        wl_name = "seg8192 4-layer GridConv encoder (O=1024/256/64/16, K=%d, N=8192)" % args.K
    elif args.workload == "cls1024":
        cfg = stack.cls1024_4layer(args.K, args.query)
        wl_name = "cls1024 4-layer GridConv encoder (O=512/128/32/8, K=%d, N=1024)" % args.K
    elif args.workload == "cls1024_shipped":
        if args.query == "gridifyknn":
            args.query = "gridify"
        cfg = stack.cls1024_shipped(args.query)
        wl_name = "cls1024 shipped 3-layer ladder, classification block (O=1024/128/1, P=64/64/128, kernel 7/3/1, N=1024; " \
                  "tf32x3 = chain of tensor-core row GEMMs, fp32 = CUDA-core kernel)"
    else:
        cfg = stack.seg81920_shipped(args.query)
        wl_name = "seg81920 shipped 3-layer GridConv encoder (O=1024/256/24, P=128/32/32, N=81920)"
    params = stack.init_params(cfg, seed=0)
    macs = layer_macs(params)
    flops_cloud = [2 * l.max_o_grid * l.max_p_grid * m for l, m in zip(cfg.layers, macs)]
    config = {"workload": "%s, query=%s, synthetic surface clouds" % (wl_name, args.query),
              "clouds_per_gpu": args.batch, "points_per_cloud": cfg.num_points, "K": args.K,
              "precision": args.precision, "cuda_graph": bool(args.graph), "parallelism": "batch-sharded clouds x%d, no collective" % world,
              "l2": "value: L2 flushed (256 MiB write) between timed steps; e2e: inputs rewritten by H2D every step and "
                    "the per-step working set (index + feature tables) exceeds the 126 MB L2"}

    metric_name = "points/sec through %d-layer GridConv @ N=%d, K=%d" % (len(cfg.layers), cfg.num_points,
                                                                        cfg.layers[0].max_p_grid)

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        # every step is a bounded sample of the workload (cpu_clouds clouds, ~0.5 s of CPU work at 3e5 points/s), so
        # the driver's K and W are honoured as given
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        val, dt, cores = time_cpu(cfg, params, args.cpu_clouds, steps, warmup)
        ref_config = dict(config, clouds_per_gpu=args.cpu_clouds, precision="fp32 (C / numpy port of the reference algorithm)",
                          cuda_graph=False, parallelism="host threads, one cloud each",
                          l2="n/a (CPU)", sample_of="the b200 arm's workload at clouds_per_gpu=%d, precision=%s"
                                                    % (args.batch, args.precision))
        line = {"impl": "reference", "metric": metric_name,
                "value": val, "unit": "points/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": ref_config,
                "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": "port",
                                 "sample": "%d clouds of %d points per step, %d steps (oracle C port: OpenMP over "
                                           "clouds for the grid ops, one thread per cloud for the numpy GridConv); "
                                           "the reference has no CPU Gridify (gridify.cc:30-39)"
                                           % (args.cpu_clouds, cfg.num_points, steps)},
                "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ this framework
    import torch
    import torch.distributed as dist
    from gridgcn_b200 import synth, shard
    import gridgcn_b200 as gg
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: gridgcn_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():  # communicator set-up happens here and at the first collective
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
    gg.build()

    B = args.batch
    # rank r owns clouds [r*B, (r+1)*B): seeds are global cloud ids, so any N sees the same data
    pool = min(B, 32)  # distinct clouds generated per rank, tiled to B (host generation cost)
    base, npts1 = synth.make_batch(pool, cfg.num_points, seed0=shard.cloud_seeds(B, rank)[0],
                                   voxels=cfg.voxels)
    reps = (B + pool - 1) // pool
    data_h = torch.from_numpy(np.tile(base, (reps, 1, 1))[:B].copy()).pin_memory()
    npts_h = torch.full((B, 1), cfg.num_points, dtype=torch.int32).pin_memory()
    data_d, npts_d = data_h.to(dev), npts_h.to(dev)
    enc = stack.GridGcnEncoder(cfg, params, dev, precision=args.precision)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out_h = torch.empty((B, cfg.layers[-1].max_o_grid, 4 + cfg.layers[-1].pt_mlp_lst[-1]),
                        dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.fill_(1)  # evict L2 between timed iterations (outside the event pair)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in evs)

    replay = stack.capture_graph(enc, data_d, npts_d) if args.graph else None

    def step_device():
        return replay() if replay is not None else enc(data_d, npts_d)

    # End-to-end: every step copies its inputs from pinned host memory and reads its result back.  Copies
    # run on their own stream, double buffered, so the H2D of step i+1 and the D2H of step i-1 overlap
    # the compute of step i (all of it inside the timed region).
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(data_d), torch.empty_like(npts_d)) for _ in range(2)]
    out_hs = [out_h, torch.empty_like(out_h).pin_memory()]

    def run_e2e(steps):
        comp = torch.cuda.current_stream(dev)
        ev_h2d = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def h2d(i):
            b = i & 1
            with torch.cuda.stream(h2d_stream):
                if i >= 2:
                    h2d_stream.wait_event(ev_comp[b])  # step i-2 no longer reads in_bufs[b]
                in_bufs[b][0].copy_(data_h, non_blocking=True)
                in_bufs[b][1].copy_(npts_h, non_blocking=True)
                ev_h2d[b].record(h2d_stream)

        start.record(comp)
        h2d_stream.wait_event(start)
        d2h_stream.wait_event(start)
        h2d(0)
        for i in range(steps):
            b = i & 1
            if i + 1 < steps:
                h2d(i + 1)  # next step's inputs travel while this step computes
            comp.wait_event(ev_h2d[b])
            out = enc(in_bufs[b][0], in_bufs[b][1])
            ev_comp[b].record(comp)
            out.record_stream(d2h_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ev_comp[b])
                out_hs[b].copy_(out, non_blocking=True)
        end.record(d2h_stream)
        torch.cuda.synchronize()
        return start.elapsed_time(end)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank, getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    barrier()
    run_e2e(2)
    barrier()
    ms_e2e = run_e2e(args.steps)
    barrier()
    sampler.stop_flag = True

    ms_total, ms_e2e = shard.max_over_ranks([ms_total, ms_e2e], device=dev)  # MAX over ranks
    points_job = world * B * cfg.num_points
    value = points_job * args.steps / (ms_total * 1e-3)
    e2e_value = points_job * args.steps / (ms_e2e * 1e-3)

    # BASELINE config 4 (training step with the gradient all-reduce): every rank takes part
    train_cfg = None
    if not args.no_configs and args.workload == "seg8192":
        try:
            train_cfg = measure_train_step(torch, dist, dev, world, block="cuda")
            tt = measure_train_step(torch, dist, dev, world, block="torch")
            train_cfg["torch_autograd_block_step_ms"] = tt["step_ms"]
            train_cfg["torch_autograd_block_eager_step_ms"] = tt["eager_step_ms"]
        except Exception as e:  # never lose the headline line to a side measurement
            train_cfg = {"error": "%s: %s" % (type(e).__name__, e)}
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------------ per-kernel rooflines (rank 0)
    peaks = load_peaks()
    enc(data_d, npts_d, keep_trace=True)
    trace = enc.trace
    reps_k = max(5, min(args.steps, 20))
    conv_ms = []
    table, prev = data_d, None
    for i, (l, conv) in enumerate(zip(cfg.layers, enc.convs)):
        tr = trace[i]
        tin = data_d if i == 0 else trace[i - 1]["table"]
        f = lambda: conv(tin, tr["nebidx"], tr["cent"], tr["centmsk"])
        for _ in range(3):
            f()
        conv_ms.append(timed(f, reps_k) / reps_k)
    query_ms = []
    for i, l in enumerate(cfg.layers):
        loc_in = data_d if i == 0 else trace[i - 1]["cent"]
        num_in = npts_d if i == 0 else trace[i - 1]["actual_centnum"]
        qf = stack.query_fn(cfg)
        kwl = dict(max_o_grid=l.max_o_grid, max_p_grid=l.max_p_grid, kernel_size=l.kernel_size, stride=1,
                   coord_shift=cfg.coord_shift, voxel_size=[l.voxel_size] * 3,
                   grid_size=[l.grid_size] * 3, loc=cfg.loc)
        f = lambda: qf(loc_in, num_in, **kwl)
        for _ in range(3):
            f()
        query_ms.append(timed(f, reps_k) / reps_k)
    dom = int(np.argmax(conv_ms))
    # executed (useful) flops of the restructured layer: the feature MLP runs once per source point
    # for layers with input features, the attention MLP once per edge ("hoisting")
    exec_cloud = []
    nprev = cfg.num_points
    for i, (l, pr) in enumerate(zip(cfg.layers, params)):
        mf = sum(st["weight"].size for st in pr["feat"])
        ma = sum(st["weight"].size for st in pr["att"])
        edges = l.max_o_grid * l.max_p_grid
        if args.precision == "fp32" or i == 0:
            exec_cloud.append(2 * edges * (mf + ma))
        else:
            exec_cloud.append(2 * (nprev * mf + edges * ma))
        nprev = l.max_o_grid
    tflops = B * exec_cloud[dom] / (conv_ms[dom] * 1e-3) / 1e12
    tflops_alg = B * flops_cloud[dom] / (conv_ms[dom] * 1e-3) / 1e12
    l0 = cfg.layers[0]
    qfn = gg.GridifyKNN
    kwq = dict(max_o_grid=l0.max_o_grid, max_p_grid=l0.max_p_grid, kernel_size=l0.kernel_size,
               stride=1, coord_shift=cfg.coord_shift, voxel_size=[l0.voxel_size] * 3,
               grid_size=[l0.grid_size] * 3, loc=cfg.loc)
    fq = lambda: qfn(data_d, npts_d, **kwq)
    for _ in range(3):
        fq()
    q_ms = timed(fq, reps_k) / reps_k
    q_bytes = B * gridify_bytes(cfg.num_points, l0.max_o_grid, l0.max_p_grid)
    q_gbs = q_bytes / (q_ms * 1e-3) / 1e9

    # DRAM traffic per launch from the committed ncu --set full capture (profiles/r02_traffic.json),
    # scaled to this run's batch; null when no capture exists for the kernel
    traffic = {}
    try:
        if not (args.workload == "seg8192" and args.K == 64 and args.query == "gridifyknn"):
            raise KeyError("the committed capture is of the default workload only")
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)["per_launch"]
        for k, v in tj.items():
            traffic[k] = (v["dram_read_bytes"] + v["dram_write_bytes"]) * B / v["clouds"]
    except Exception:
        pass
    tensor_peak = peaks["bf16_tflops_sustained"] if "bf16_tflops_sustained" in peaks else peaks["bf16_tflops"]
    roofline = {"kernel": "gridconv layer %d (%s)" % (dom, args.precision), "bound": "tensor",
                "achieved": tflops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tflops / tensor_peak,
                "traffic": traffic.get("edge_L%d" % dom), "ms_per_launch": conv_ms[dom], "layer_ms": conv_ms,
                "algorithmic_tflops": tflops_alg,
                "note": "achieved = flops the restructured layer executes (x1, split passes not counted) / "
                        "duration; algorithmic_tflops uses SURVEY s8d's per-edge formula 2*O*K*MAC",
                "peak_source": peaks["_source"] + " dense bf16 cuBLAS (sustained); tf32 nominal peak is half of bf16"}
    roofline_hbm = {"kernel": "gridifyknn (build + query) N=%d O=%d P=%d" % (cfg.num_points, l0.max_o_grid, l0.max_p_grid),
                    "bound": "hbm", "achieved": q_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": q_gbs / peaks["hbm_gbs"],
                    "traffic": (traffic["knn_query"] + traffic.get("build", 0.0)) if "knn_query" in traffic else None,
                    "ms_per_launch": q_ms,
                    "algorithmic_bytes": q_bytes, "peak_source": peaks["_source"]}

    # ------------------------------------------------------------------ the other BASELINE configurations
    configs_block = None
    if not args.no_configs and args.workload == "seg8192" and args.K == 64 and args.query == "gridifyknn":
        configs_block = {}
        def side(name, fn):
            try:
                configs_block[name] = fn()
            except Exception as e:
                configs_block[name] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
        side("config2_cls1024_B32_K32", lambda: measure_encoder(torch, flush, peaks, stack.cls1024_4layer(32), 32,
                                                                 args.precision, graph=True))
        side("config2_cls1024_shipped_cls_block_B32", lambda: measure_encoder(torch, flush, peaks, stack.cls1024_shipped(), 32,
                                                                               args.precision, steps=3))
        side("config3_seg8192_shipped_graph_B12", lambda: measure_seg_graph(torch, flush, 12, args.precision))
        side("config4_seg81920_B3", lambda: measure_encoder(torch, flush, peaks, stack.seg81920_shipped("gridify"), 3,
                                                             args.precision, graph=True))
        for k in (16, 32, 128):
            side("config5_seg8192_K%d_B%d" % (k, B), lambda k=k: measure_encoder(torch, flush, peaks, stack.seg8192_4layer(k),
                                                                                 B, args.precision, steps=3))
        configs_block["config5_seg8192_K64_B%d" % B] = {
            "note": "the headline line itself", "ms_per_step": ms_total / args.steps, "points_per_s": value,
            "breakdown_ms": {"gather": query_ms, "mlp": conv_ms, "gather_total": sum(query_ms), "mlp_total": sum(conv_ms),
                             "gather_share": sum(query_ms) / (sum(query_ms) + sum(conv_ms))}}
        configs_block["config4_train_step"] = train_cfg
    cpu_val, cpu_dt, cores = time_cpu(cfg, params, args.cpu_clouds, 2, 1)
    line = {"metric": metric_name,
            "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "tf32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(world * (data_h.numel() * 4 + npts_h.numel() * 4)),
                    "d2h_bytes_per_step": int(world * out_h.numel() * 4)},
            # per layer: voxel-table build, query (+ the sampling kernel for occaware), edge kernel; layers with input
            # features run the per-point MLP kernel first in the tensor-core precisions
            "gpu_launches": args.steps * (len(cfg.layers) * (4 if args.query.startswith("occaware") else 3) +
                                          (0 if args.precision == "fp32" else len(cfg.layers) - 1)),
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": {"value": cpu_val, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": "%d clouds of %d points, 2 steps (oracle C port: OpenMP over clouds for the grid "
                                       "ops, one thread per cloud for the numpy GridConv)" % (args.cpu_clouds, cfg.num_points)},
            "breakdown_ms": {"query": query_ms, "gridconv": conv_ms,
                             "note": "each operator timed alone, L2 flushed before every call"},
            "clocks": sampler.summary()}
    if configs_block is not None:
        line["configs"] = configs_block
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
