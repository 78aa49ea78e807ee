#!/usr/bin/env python
"""Headline benchmark: implicit-decoder queries/sec on the GREATER-shape synthetic workload
(BASELINE.json configs[1]: 14336 points, 12 frames, 524288 grid queries -> 534,528 query
points, 6 MLP blocks, 2 cross-attention layers, implicit_batch_size 32768).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # this repo (CUDA, sm_100a)
    python bench.py --impl reference ...                            # CPU arm (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU

A step = one pass of the decoder hot path over all query mini-batches of one frame with
the scene encoding resident (the eval/inference.py:204-246 loop).  Weak scaling: every
rank decodes one whole frame (frames are independent: eval/test.py:67 loops over them)
and the ranks all-gather their (N_q, d_out) outputs over NCCL.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'occlusions-4d_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'implicit_queries_per_sec'
UNIT = 'queries/s'
FLOP_PER_QUERY = 47.89e6     # SURVEY.md section 8d / BASELINE.md section 2 (G = 9)
FAMILIES = ['dense_layer', 'knn', 'fps', 'attn_gather_softmax', 'misc', 'fused_attn_mlp']


def workload_config(batch):
    return {'workload': 'GREATER synthetic: n_points=14336, video_len=12, 524288 grid queries (534528 points), '
                        'attention mode, 6 MLP blocks, 2 cross-attn layers (K=14), d_hidden=416, d_out=9, '
                        'seeded random-init weights',
            'implicit_batch_size': batch,
            'l2': 'a 256 MiB buffer is overwritten before every timed step (L2 flush); per-step activations '
                  'are > 1 GiB anyway',
            'parallelism': 'one frame of queries per rank, NCCL all-gather of outputs'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def golden_scene():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c2_greater_seeded.npz'))
    return torch.from_numpy(z['abstract']), torch.from_numpy(z['glob'])


def cpu_reference_rate(sample, threads=None):
    """Oracle port (torch CPU fp32) of the decoder on `sample` queries of the workload; returns
    (queries/s, seconds, threads)."""
    from oracle import o4d_oracle as orc
    from tests import configs
    cfg = configs.C2_GREATER
    if threads:
        torch.set_num_threads(threads)
    _, dec = configs.build_modules(cfg)
    sd = orc.cast_state(dec.state_dict(), torch.float32)
    abstract, glob = golden_scene()
    q = configs.synthetic_queries(cfg)
    sel = torch.linspace(0, q.shape[0] - 1, sample).long()
    q = q[sel]
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=4096)
        dt = time.perf_counter() - t0
    return sample / dt, dt, torch.get_num_threads()


def torch_eager_gpu_rate(dev, dec, abstract, glob, q_dev, batch, batches=2):
    """The north star's denominator on the SAME GPU: the reference's PyTorch eager formulation of the decoder
    (oracle port on `dev`, fp32, TF32 off: full distance matrix + sort per kNN, materialised (N, K, D) gathers,
    one cuBLAS SGEMM per nn.Linear) on `batches` mini-batches of the workload.  Baseline measurement only."""
    from oracle import o4d_oracle as orc
    from tests import configs
    cfg = configs.C2_GREATER

    def knn_on_device(query_xyz, ref_xyz, k, sqrt=False, chunk=2048):
        idx, dist = [], []
        for s in range(0, query_xyz.shape[0], chunk):
            d2 = orc._pair_sqdist(query_xyz[s:s + chunk], ref_xyz)
            if sqrt:
                d2 = d2.sqrt()
            val, order = torch.sort(d2, dim=1, stable=True)
            idx.append(order[:, :k])
            dist.append(val[:, :k])
        return torch.cat(idx), torch.cat(dist)

    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    saved = orc.knn_indices
    orc.knn_indices = knn_on_device          # the oracle's own helper allocates on the CPU and uses numpy's sqrt
    try:
        sd = {k: v.detach().float().to(dev) for k, v in dec.state_dict().items()}
        n = min(q_dev.shape[0], batches * batch)
        q = q_dev[torch.linspace(0, q_dev.shape[0] - 1, n, device=q_dev.device).long()]
        with torch.no_grad():
            ref_out = orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=batch)[0]   # warm-up
            if dev.type == 'cuda':
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=batch)
            if dev.type == 'cuda':
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return {'value': n / dt, 'unit': UNIT, 'kind': 'port', 'seconds': dt, 'tf32': False,
                'sample': '%d of %d grid queries (evenly strided) in mini-batches of %d, decoder only, oracle port '
                          'as torch eager fp32 on the same GPU' % (n, q_dev.shape[0], batch)}, q, ref_out
    finally:
        orc.knn_indices = saved
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port; the reference is Python and cannot travel
    to the GPU box) on the host cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import o4d_oracle as orc
    from tests import configs
    cfg = configs.C2_GREATER
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    _, dec = configs.build_modules(cfg)
    sd = orc.cast_state(dec.state_dict(), torch.float32)
    abstract, glob = golden_scene()
    q_all = configs.synthetic_queries(cfg)
    # calibrate on 1024 queries, then size a step so that warmup+steps take about two minutes
    with torch.no_grad():
        t0 = time.perf_counter()
        orc.decoder_forward(sd, cfg['implicit_args'], q_all[:1024], abstract, glob, chunk=1024)
        rate = 1024 / (time.perf_counter() - t0)
    total = args.steps + args.warmup
    sample = int(max(1024, min(q_all.shape[0], rate * 120.0 / total)))
    sample = sample // 1024 * 1024
    sel = torch.linspace(0, q_all.shape[0] - 1, sample).long()
    q = q_all[sel]
    with torch.no_grad():
        for _ in range(args.warmup):
            orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=4096)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=4096)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = '%d of %d grid queries per step (evenly strided), decoder only, scene encoding resident' % (
        sample, q_all.shape[0])
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args.batch),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                             'sample': desc},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def sampler_and_loss_timing(dev):
    """SURVEY.md 8f rows 1 and 3 at the config-5 shape, timed on their own (the sampler is excluded from the
    train-step timing): GuidedImplicitPointSampler on 28,672-point CARLA target frames
    (7,168 solid + 10,035 air queries per frame) and the fused loss heads on 17,203 x 18 logits."""
    try:
        import logging
        from o4d import geometry as geo, loss as o4d_loss
        g = torch.Generator().manual_seed(1830)
        lo, hi = torch.tensor([0.0, -16.0, -1.0]), torch.tensor([40.0, 16.0, 6.4])
        frames = []
        for _ in range(4):
            f = torch.rand(1, 28672, 11, generator=g)
            f[0, :, :3] = f[0, :, :3] * (hi - lo) + lo
            f[0, :, 5] = torch.randint(0, 13, (28672,), generator=g).float()
            frames.append(f.to(dev))
        sizes = [torch.tensor([28672]) for _ in range(4)]
        valo, num_valo = torch.zeros(1, 4), torch.zeros(1, dtype=torch.int64)
        smp = geo.GuidedImplicitPointSampler(
            logging.getLogger('bench'), min_z=-1.0, cube_bounds=16.0, point_occupancy_radius=0.2, num_solid=7168,
            num_air=10035, predict_segmentation=True, semantic_classes=13, data_kind='carla',
            point_sample_bias='none', cube_mode=4, device_rng=True)

        def timed(fn, iters):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters

        sampler_ms = timed(lambda: smp(frames, sizes, valo, num_valo, 1), 10)
        solid_in, air_in, solid_tgt, air_tgt, _, _ = smp(frames, sizes, valo, num_valo, 1)
        target = torch.cat([solid_tgt[0], air_tgt[0]], dim=0)
        out = torch.randn(target.shape[0], 18, generator=g).to(dev).requires_grad_(True)
        w = torch.ones(4, device=dev)

        def heads():
            out.grad = None
            (o4d_loss.implicit_loss_heads(out, target, 'rgb', 13, True) * w).sum().backward()

        return {'sampler_ms_per_frame': sampler_ms, 'loss_heads_fwd_bwd_ms_per_frame': timed(heads, 20),
                'queries_per_frame': int(target.shape[0]), 'target_points_per_frame': 28672,
                'what': 'GuidedImplicitPointSampler (bias none, GPU generator) and fused loss heads (rgb + density + '
                        '13-class segmentation + tracking), timed on their own; not part of ms_per_step'}
    except Exception as exc:                   # noqa: BLE001 -- a side measurement must not take the bench line down
        return {'error': '%s: %s' % (type(exc).__name__, exc)}


def train_step_bench(dev, world, rank, steps):
    """BASELINE.json configs[4] per GPU: one CARLA-4D sample (14336 points), 4 frames x 17,203 query
    points (args.py:254,257 / train.py:270-274), forward + backward through the nn.Module API
    (o4d/autograd.py -> backward kernels), gradient averaging over ranks, AdamW.  The loss heads and the
    optimizer are the caller's torch code (out of scope); the query sampler is
    excluded.  Returns a dict for the bench line."""
    import torch.distributed as dist
    from o4d import parallel
    from tests import configs
    cfg = configs.C3_CARLA
    enc, dec = configs.build_modules(cfg, dev)
    enc.train()
    dec.train()
    params = list(enc.parameters()) + list(dec.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4)
    g = torch.Generator().manual_seed(1830 + rank)
    pcl = configs.synthetic_cloud(cfg).to(dev)
    frames, per_frame = 4, 17203
    lo = torch.tensor([0.0, -16.0, -1.0])
    hi = torch.tensor([40.0, 16.0, 6.4])
    queries = []
    for f in range(frames):
        q = torch.rand(per_frame, 4, generator=g)
        q[:, :3] = q[:, :3] * (hi - lo) + lo
        q[:, 3] = float(f)
        queries.append(q.to(dev))
    target = torch.rand(frames, per_frame, 6, generator=g).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        abstract, glob, _ = enc(pcl[None], False)
        total = 0.0
        for f in range(frames):
            out, _ = dec(queries[f], abstract[0], glob[0], None)
            t = target[f]
            loss = torch.nn.functional.binary_cross_entropy_with_logits(out[:, 0], (t[:, 0] > 0.5).float()) + \
                (out[:, 1:4] - t[:, 1:4]).abs().mean() + \
                torch.nn.functional.cross_entropy(out[:, 5:18], (t[:, 5] * 12.99).long())
            total = total + loss
        (total / frames).backward()
        if world > 1:
            parallel.allreduce_gradients([enc, dec])
        opt.step()
        return total

    step()                                   # warm-up (workspaces, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        total = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    del opt
    extras = sampler_and_loss_timing(dev) if rank == 0 else None
    return {'sampler_and_loss_heads': extras, 'ms_per_step': ms, 'samples_per_step': world, 'queries_per_sample': frames * per_frame,
            'points_per_sample': cfg['n_points'], 'query_grads_per_s': world * frames * per_frame / (ms / 1e3),
            'loss': float(total.detach()) / frames, 'steps': steps, 'precision': 'bf16x3 (fp32-grade) forward and backward',
            'what': 'CARLA config 5 shape: encoder + 4 decoder frames forward/backward, grad all-reduce, AdamW; '
                    'sampler excluded'}


def run_o4d(args):
    import torch.distributed as dist
    import o4d
    from o4d import _lib, ops, parallel
    from tests import configs

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device for the o4d arm (there is no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.lib()
    cfg = configs.C2_GREATER
    batch = args.batch
    enc, dec = configs.build_modules(cfg, dev)
    if args.precision is not None:
        enc.o4d_precision = dec.o4d_precision = args.precision
    precision = ops.default_precision() if args.precision is None else args.precision
    pcl = configs.synthetic_cloud(cfg).to(dev)
    # frame (time index) per rank: frames are the independent units sharded across GPUs
    q_host = configs.synthetic_queries(cfg, time_idx=rank % cfg['video_len']).contiguous().pin_memory()
    nq = q_host.shape[0]
    q_dev = q_host.to(dev)
    d_out = cfg['implicit_args']['d_out']

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ---- encoder: once per scene; timed on its own (pts/s), outside the query metric
        for _ in range(max(1, args.warmup)):
            abstract, glob, _ = enc(pcl[None], False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        enc_iters = max(2, min(args.steps, 5))
        for _ in range(enc_iters):
            abstract, glob, _ = enc(pcl[None], False)
        e1.record()
        torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1) / enc_iters
        abstract, glob = abstract[0].contiguous(), glob[0].contiguous()
        scene = dec.o4d_scene(abstract, glob)
        dcfg, dparams = dec.o4d_config(), dec.o4d_params()
        out_dev = torch.empty((nq, d_out), dtype=torch.float32, device=dev)
        gathered = torch.empty((world * nq, d_out), dtype=torch.float32, device=dev) if world > 1 else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step_device(comm=True):
            flush.fill_(1)                                   # L2 flush (torch fill, not an o4d kernel)
            for s in range(0, nq, batch):
                ops.decoder_forward(dcfg, dparams, scene, q_dev[s:s + batch], want_penult=False,
                                    out=out_dev[s:s + batch])
            if world > 1 and comm:
                dist.all_gather_into_tensor(gathered, out_dev)

        for _ in range(args.warmup):
            step_device()
        sampler = ClockSampler(local)
        barrier()
        if rank == 0:
            sampler.start()
        launches0 = lib.o4d_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            step_device()
        t1.record()
        barrier()
        launches = lib.o4d_launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        ms = max_over_ranks(t0.elapsed_time(t1))
        value = world * nq * args.steps / (ms / 1e3)

        # ---- end to end: HOST query buffer -> C-ABI host loop (H2D, forward, D2H per mini-batch)
        out_host = torch.empty((nq, d_out), dtype=torch.float32).pin_memory()
        ops.decoder_run_host(dcfg, dparams, scene, q_host, batch, out_host)
        barrier()
        w0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            ops.decoder_run_host(dcfg, dparams, scene, q_host, batch, out_host)   # synchronises inside
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - w0) * 1e3)
        e2e_value = world * nq * e2e_steps / (e2e_ms / 1e3)
        e2e_matches = bool(torch.equal(out_host, out_dev.cpu()))

        # ---- roofline of the dominant kernel family: one extra profiled step right after
        roofline, families = None, None
        if rank == 0:
            lib.o4d_profile_enable(1)
            step_device(comm=False)                          # rank 0 only: no collective here
            torch.cuda.synchronize()
            import ctypes
            n = len(FAMILIES)
            ms_a, fl_a, ct_a = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
            lib.o4d_profile_read(n, ms_a, fl_a, ct_a)
            lib.o4d_profile_enable(0)
            families = {FAMILIES[i]: {'ms': ms_a[i], 'launches': ct_a[i], 'tflops': (fl_a[i] / ms_a[i] / 1e9) if ms_a[i] > 0 else 0.0}
                        for i in range(n) if ct_a[i]}
            top = max(families, key=lambda k: families[k]['ms'])
            peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
            if os.path.isfile(peaks_path):
                pk = json.load(open(peaks_path))
                peak, which = float(pk['bf16_tflops_sustained']), 'MEASURED_PEAKS.json bf16_tflops_sustained'
            else:
                peak, which = 1400.0, 'fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)'
            step_total = sum(f['ms'] for f in families.values())
            ach = families[top]['tflops']
            traffic = None
            tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
            if os.path.isfile(tpath):      # dram bytes per launch of the dominant kernel, from the committed ncu capture
                traffic = json.load(open(tpath)).get(top, {}).get('dram_bytes_per_launch')
            roofline = {'bound': 'tensor', 'kernel': top, 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': ach / peak, 'traffic': traffic,
                        'traffic_source': 'profiles/ncu_traffic.json (ncu --set full, dram__bytes_read+write per launch)',
                        'peak_source': which,
                        'share_of_step': families[top]['ms'] / step_total if step_total else None,
                        'avg_launch_ms': families[top]['ms'] / families[top]['launches'],
                        'measured': 'CUDA events around every launch of the family, one extra step after the timed region',
                        'whole_step_algorithmic_tflops': FLOP_PER_QUERY * value / world / 1e12}

    train = None
    if not args.no_train_step:
        torch.cuda.empty_cache()
        train = train_step_bench(dev, world, rank, max(1, min(args.steps, 3)))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, secs, threads = cpu_reference_rate(args.cpu_sample)
        cpu_base = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                    'sample': '%d of %d grid queries (evenly strided), decoder only, oracle port torch-CPU fp32, '
                              '%.1f s' % (args.cpu_sample, nq, secs)}
    torch_gpu = None
    if rank == 0 and world == 1 and not args.no_torch_gpu_baseline:
        try:                                   # a reported baseline: never allowed to take the bench line down
            with torch.no_grad():
                torch_gpu, q_s, ref_s = torch_eager_gpu_rate(dev, dec, abstract, glob, q_dev, batch)
                ours = torch.cat([ops.decoder_forward(dcfg, dparams, scene, q_s[s:s + batch], want_penult=False)[0]
                                  for s in range(0, q_s.shape[0], batch)])
            torch_gpu['o4d_over_torch_eager'] = value / torch_gpu['value']
            torch_gpu['max_rel_err_vs_torch_eager'] = float((ours - ref_s).abs().max() / ref_s.abs().max())
            del q_s, ref_s, ours
        except Exception as exc:               # noqa: BLE001
            torch_gpu = {'error': '%s: %s' % (type(exc).__name__, exc)}
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': {0: 'f32', 1: 'bf16x3', 2: 'bf16'}[precision],
                'data': 'synthetic', 'config': workload_config(batch), 'queries_per_step_per_gpu': nq,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': nq * 16,
                        'd2h_bytes_per_step': nq * d_out * 4, 'steps': e2e_steps,
                        'bit_identical_to_device_path': e2e_matches},
                'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'kernel_families': families,
                'cpu_baseline': cpu_base, 'torch_gpu_baseline': torch_gpu,
                'encoder': {'pts_per_s': cfg['n_points'] / (enc_ms / 1e3), 'ms': enc_ms, 'n_points': cfg['n_points']},
                'train_step': train, 'tcgen05': bool(lib.o4d_has_tcgen05())}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='o4d', choices=['o4d', 'reference'])
    ap.add_argument('--batch', type=int, default=32768, help='implicit_batch_size')
    ap.add_argument('--precision', type=int, default=None, help='0 fp32 CUDA cores, 1 tcgen05 bf16x3, 2 bf16')
    ap.add_argument('--cpu-sample', type=int, default=196608)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train-step', action='store_true', help='skip the config-5 training-step timing')
    ap.add_argument('--no-torch-gpu-baseline', action='store_true',
                    help='skip the PyTorch-eager timing of the same decoder on the same GPU')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_o4d(args)


if __name__ == '__main__':
    main()
