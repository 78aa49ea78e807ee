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
import gc
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
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def mark_end(self):
        self.t_end = time.monotonic()

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
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(',')]))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        inside = [r for t, r in self.rows if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.05)]
        if not inside:                      # timed region shorter than one sampling period: nearest samples under load
            inside = [r for _, r in self.rows[-3:]]
        for r in inside:
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


def reference_modules(cfg, dev):
    """The UNMODIFIED reference modules (oracle/ref_loader.py: /root/reference, or its byte-for-byte copy
    oracle/_ref made by oracle/make_ref.py) with the config's seeded default init -- the same weights
    configs.build_modules gives the o4d modules (same creation order, same RNG stream).  None if unavailable."""
    from oracle import ref_loader
    if not ref_loader.available():
        return None
    ref = ref_loader.load()
    torch.manual_seed(cfg['seed'])
    with ref_loader.quiet():
        enc = ref['model'].PointCompletionNetV3(**cfg['pcl_args']).eval()
        dec = ref['implicit'].LocalPclResnetFC(**cfg['implicit_args']).eval()
    return ref_loader, enc.to(dev), dec.to(dev)


def _median(xs):
    return float(np.median(np.asarray(xs, dtype=np.float64)))


def cpu_reference_rate(sample, threads=None):
    """The reference decoder (implicit.LocalPclResnetFC.forward, torch CPU fp32) on `sample` evenly strided queries
    of the workload in mini-batches of 4096; falls back to the oracle port when the reference copy is absent.
    Returns (queries/s, seconds, threads, kind)."""
    from tests import configs
    cfg = configs.C2_GREATER
    if threads:
        torch.set_num_threads(threads)
    abstract, glob = golden_scene()
    q = configs.synthetic_queries(cfg)
    sel = torch.linspace(0, q.shape[0] - 1, sample).long()
    q = q[sel]
    mods = reference_modules(cfg, 'cpu')
    with torch.no_grad():
        if mods is not None:
            loader, _, dec = mods
            with loader.quiet():
                dec(q[:256], abstract, glob, None)                       # first-call overheads out of the timing
                t0 = time.perf_counter()
                for s0 in range(0, q.shape[0], 4096):
                    dec(q[s0:s0 + 4096], abstract, glob, None)
                dt = time.perf_counter() - t0
            kind = 'reference'
        else:
            from oracle import o4d_oracle as orc
            _, dec = configs.build_modules(cfg)
            sd = orc.cast_state(dec.state_dict(), torch.float32)
            t0 = time.perf_counter()
            orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=4096)
            dt = time.perf_counter() - t0
            kind = 'port'
    return sample / dt, dt, torch.get_num_threads(), kind


def reference_decoder_gpu(dev, cfg, abstract, glob, q_host, batch, o4d_value, o4d_out, warm=10, iters=20):
    """SURVEY 8d "Reference timing (1)" = the north star's denominator: implicit.LocalPclResnetFC itself, PyTorch
    eager fp32 with TF32 off, on the SAME GPU, same weights, same scene, the eval/inference.py:204-246 loop over
    ALL mini-batches of the frame.  10 warm-up passes, median of 20 (device-resident queries, no per-batch D2H);
    then the loop exactly as the reference runs it (numpy slice -> .to(device) -> forward -> .cpu().numpy() per
    mini-batch), 3 + 5 passes.  Also the largest relative difference between the reference's and o4d's outputs
    over the whole frame."""
    mods = reference_modules(cfg, dev)
    if mods is None:
        return {'error': 'reference copy (oracle/_ref) not present'}
    loader, _, dec = mods
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        q_np = q_host.numpy()
        q_dev = q_host.to(dev)
        nq = q_dev.shape[0]

        def pass_device():
            return [dec(q_dev[s:s + batch], abstract, glob, None)[0] for s in range(0, nq, batch)]

        def pass_host():
            outs = []
            for s in range(0, nq, batch):
                qb = torch.from_numpy(q_np[s:s + batch]).to(dev)                  # inference.py:206
                o = dec(qb, abstract, glob, None)[0]
                o[..., 0] = torch.sigmoid(o[..., 0])                              # :218
                outs.append(o.detach().cpu().numpy())                             # :245
            return outs

        with torch.no_grad(), loader.quiet():
            ref_out = torch.cat(pass_device())
            row_err = (o4d_out - ref_out).abs().max(dim=1)[0] / ref_out.abs().max()
            err = {'max': float(row_err.max()), 'p9999': float(torch.quantile(row_err[::8].float(), 0.9999)),
                   'median': float(row_err.median()), 'rows_above_1e-3': int((row_err > 1e-3).sum()),
                   'rows_above_1e-4': int((row_err > 1e-4).sum()), 'rows': int(row_err.numel()),
                   'note': 'per-row max |o4d - reference| / max |reference|; the reference on the GPU computes kNN '
                           'distances with its own rounding, so rows whose k-th / (k+1)-th neighbours are a near tie '
                           'pick a different neighbour -- those rows carry the maximum'}
            del ref_out, row_err
            for _ in range(warm - 1):
                pass_device()
            times = []
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                pass_device()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            host_times = []
            for i in range(3 + 5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pass_host()
                torch.cuda.synchronize()
                if i >= 3:
                    host_times.append((time.perf_counter() - t0) * 1e3)
        rate = nq / (_median(times) / 1e3)
        rate_host = nq / (_median(host_times) / 1e3)
        return {'value': rate, 'unit': UNIT, 'kind': 'reference', 'ms_per_frame': _median(times), 'tf32': False,
                'warmup_passes': warm, 'timed_passes': iters, 'statistic': 'median',
                'with_per_batch_h2d_d2h': {'value': rate_host, 'ms_per_frame': _median(host_times), 'timed_passes': 5},
                'o4d_over_torch_eager': o4d_value / rate,
                'rel_err_o4d_vs_reference_all_queries': err,
                'sample': 'all %d grid queries of the frame in mini-batches of %d: the unmodified reference module '
                          '(model/implicit.py LocalPclResnetFC, copy in oracle/_ref) as torch eager fp32 on the same '
                          'GPU, scene encoding resident' % (nq, batch)}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


def reference_encoder_baselines(dev, cfg, pcl, cpu=True):
    """Encoder pts/s of the unmodified reference module on the same GPU (torch eager fp32) and on the host cores.
    torch_cluster is not installable: its two ops run through the CPU restatement oracle/cluster_ops.py,
    whose time is measured separately and SUBTRACTED in `*_excl_cluster_stub` -- an upper bound on what the
    reference with the real extension could reach."""
    from oracle import cluster_ops
    import torch_cluster as stub                      # the stub module ref_loader put on sys.path
    out = {}
    spent = [0.0]
    saved = (stub.fps, stub.knn)

    def timed(fn):
        def wrap(*a, **k):
            if dev != 'cpu':
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **k)
            spent[0] += time.perf_counter() - t0
            return r
        return wrap
    try:
        for where in (['gpu_eager', 'cpu'] if cpu else ['gpu_eager']):
            d = dev if where == 'gpu_eager' else 'cpu'
            mods = reference_modules(cfg, d)
            if mods is None:
                return {'error': 'reference copy (oracle/_ref) not present'}
            loader, enc, _ = mods
            import modules as ref_modules            # flat module of the reference (model/modules.py)
            ref_modules.torch_cluster.fps, ref_modules.torch_cluster.knn = timed(cluster_ops.fps), timed(cluster_ops.knn)
            x = pcl.to(d)[None]
            reps = 3 if where == 'gpu_eager' else 1
            with torch.no_grad(), loader.quiet():
                if where == 'gpu_eager':
                    enc(x, False)
                    torch.cuda.synchronize()
                spent[0] = 0.0
                t0 = time.perf_counter()
                for _ in range(reps):
                    enc(x, False)
                if where == 'gpu_eager':
                    torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / reps
            stub_s = spent[0] / reps
            n = pcl.shape[0]
            out[where] = {'pts_per_s': n / dt, 'ms': dt * 1e3, 'cluster_stub_ms': stub_s * 1e3,
                          'pts_per_s_excl_cluster_stub': n / max(dt - stub_s, 1e-9), 'kind': 'reference',
                          'cores': torch.get_num_threads() if where == 'cpu' else None}
            del enc
            if where == 'gpu_eager':
                torch.cuda.empty_cache()
    finally:
        stub.fps, stub.knn = saved
    return out


def run_reference(args):
    """CPU arm: the reference's own decoder module (oracle/_ref copy of model/implicit.py; oracle port if the copy
    is absent) on the host cores with all threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from tests import configs
    cfg = configs.C2_GREATER
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    abstract, glob = golden_scene()
    q_all = configs.synthetic_queries(cfg)
    mods = reference_modules(cfg, 'cpu')
    if mods is not None:
        loader, _, dec = mods
        kind = 'reference'

        def forward(q):
            with loader.quiet():
                for s0 in range(0, q.shape[0], 4096):
                    dec(q[s0:s0 + 4096], abstract, glob, None)
    else:
        from oracle import o4d_oracle as orc
        _, dec = configs.build_modules(cfg)
        sd = orc.cast_state(dec.state_dict(), torch.float32)
        kind = 'port'

        def forward(q):
            orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=4096)
    # calibrate on 2048 queries, then size a step so that warmup+steps take about two minutes
    with torch.no_grad():
        forward(q_all[:1024])
        t0 = time.perf_counter()
        forward(q_all[:2048])
        rate = 2048 / (time.perf_counter() - t0)
    total = args.steps + args.warmup
    sample = int(max(1024, min(q_all.shape[0], rate * 120.0 / total)))
    sample = sample // 1024 * 1024
    sel = torch.linspace(0, q_all.shape[0] - 1, sample).long()
    q = q_all[sel]
    with torch.no_grad():
        for _ in range(args.warmup):
            forward(q)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            forward(q)
        dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = '%d of %d grid queries per step (evenly strided, mini-batches of 4096), decoder only, scene encoding ' \
           'resident; %s' % (sample, q_all.shape[0], 'unmodified reference module (oracle/_ref copy of model/implicit.py)'
                             if kind == 'reference' else 'oracle port')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args.batch),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind,
                             'sample': desc},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def family_profile(lib, step_fn):
    """Per-kernel-family device time of one call of step_fn: library-side CUDA-event pairs on the launching stream
    around every launch (o4d_profile_enable)."""
    import ctypes
    lib.o4d_profile_enable(1)
    step_fn()
    torch.cuda.synchronize()
    n = len(FAMILIES)
    ms_a, fl_a, ct_a = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
    lib.o4d_profile_read(n, ms_a, fl_a, ct_a)
    lib.o4d_profile_enable(0)
    return {FAMILIES[i]: {'ms': ms_a[i], 'launches': ct_a[i], 'tflops': (fl_a[i] / ms_a[i] / 1e9) if ms_a[i] > 0 else 0.0}
            for i in range(n) if ct_a[i]}


def measured_peak():
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(peaks_path):
        pk = json.load(open(peaks_path))
        return float(pk['bf16_tflops_sustained']), 'MEASURED_PEAKS.json bf16_tflops_sustained'
    return 1400.0, 'fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)'


def carla_config3(dev, lib, batch, steps):
    """BASELINE.json configs[2] throughput: CARLA-4D shape, abstract_levels=2 -> M = 2124 abstract points (4x the
    GREATER cloud), d_out = 18 (13-class segmentation head), 541,314 grid queries, same decoder loop."""
    from o4d import ops
    from tests import configs
    cfg = configs.C3_CARLA
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c3_carla_seeded.npz'))
    abstract, glob = torch.from_numpy(z['abstract']).to(dev), torch.from_numpy(z['glob']).to(dev)
    _, dec = configs.build_modules(cfg, dev)
    q = configs.synthetic_queries(cfg).to(dev)
    nq, d_out = q.shape[0], cfg['implicit_args']['d_out']
    scene = dec.o4d_scene(abstract, glob)
    dcfg, dparams = dec.o4d_config(), dec.o4d_params()
    out = torch.empty((nq, d_out), dtype=torch.float32, device=dev)

    def step():
        for s0 in range(0, nq, batch):
            ops.decoder_forward(dcfg, dparams, scene, q[s0:s0 + batch], want_penult=False, out=out[s0:s0 + batch])

    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    fam = family_profile(lib, step)
    top = max(fam, key=lambda k: fam[k]['ms'])
    peak, _ = measured_peak()
    sel = torch.linspace(0, nq - 1, 4096).long().to(dev)        # the golden subset of the reference (make_golden.py)
    gold = torch.from_numpy(z['out']).to(dev)
    # abstract_levels=2 duplicates every level-2 position in level 1: exact distance ties, where the reference's topk
    # choice is implementation-defined.  Compare on rows whose neighbour sets are unambiguous (as the tests do).
    from tests.test_oracle import boundary_tie_free
    ok = boundary_tie_free(torch.from_numpy(z['query'])[:, :3], torch.from_numpy(z['abstract'])[:, :3],
                           [cfg['implicit_args']['num_local_features'], cfg['implicit_args']['cross_attn_neighbors']]).to(dev)
    return {'queries_per_s': nq / (ms / 1e3), 'ms_per_step': ms, 'queries_per_step': nq, 'm_abstract': int(abstract.shape[0]),
            'd_out': d_out, 'kernel_families': fam,
            'roofline': {'bound': 'tensor', 'kernel': top, 'achieved': fam[top]['tflops'], 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': fam[top]['tflops'] / peak},
            'max_rel_err_vs_reference_golden': float((out[sel][ok] - gold[ok]).abs().max() / gold[ok].abs().max()),
            'golden_rows_compared': int(ok.sum()), 'golden_rows_with_exact_distance_ties': int((~ok).sum())}


def strong_scaling(dev, world, rank, enc, dec, pcl, cfg, frames=2):
    """BASELINE.json configs[3] / SURVEY 8e: ONE frame of 2,097,152 `random` queries split contiguously over the
    ranks (o4d.parallel.decode_sharded: no data-path collective, one all_gather_into_tensor of the output shards),
    implicit_batch_size in {8192, 32768, 131072}.  The timed region is the whole frame as eval/inference.py runs
    it: encoder (replicated on every rank) + prepare_scene + this rank's mini-batches + the all-gather.  Rank 0 also
    decodes the whole frame alone and compares bit for bit."""
    import torch.distributed as dist
    from o4d import geometry, ops, parallel
    state = np.random.get_state()
    np.random.seed(cfg['seed'])
    q_np = geometry.sample_implicit_points_blind_numpy(2097152, cfg['min_z'], cfg['cr_cube_bounds'], 3, cfg['kind'],
                                                       cfg['cube_mode'], 'random')
    np.random.set_state(state)
    q_all = torch.from_numpy(q_np).to(dev)
    dcfg, dparams = dec.o4d_config(), dec.o4d_params()
    res = {'queries': int(q_all.shape[0]), 'frames_timed': frames, 'by_batch': {}}
    gathered = None
    enc_ms = None
    for bs in (8192, 32768, 131072):
        def frame():
            abstract, glob, _ = enc(pcl[None], False)
            a, g = abstract[0].contiguous(), glob[0].contiguous()
            scene = dec.o4d_scene(a, g)                               # new tensors -> prepare_scene runs every frame

            def fn(b):
                return ops.decoder_forward(dcfg, dparams, scene, b, want_penult=False)[0]
            return parallel.decode_sharded(fn, q_all, bs), (a, g, scene)

        frame()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            gathered, keep = frame()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / frames], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        res['by_batch'][str(bs)] = {'ms_per_frame': ms, 'queries_per_s': q_all.shape[0] / (ms / 1e3)}
    # replicated (un-sharded) part of a frame, timed on its own: the limiter of the 1 -> 8 curve
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        abstract, glob, _ = enc(pcl[None], False)
        dec.o4d_scene(abstract[0].contiguous(), glob[0].contiguous())
    e1.record()
    torch.cuda.synchronize()
    res['replicated_encoder_plus_prepare_ms'] = e0.elapsed_time(e1) / 3
    if rank == 0:
        a, g, scene = keep
        single = torch.cat([ops.decoder_forward(dcfg, dparams, scene, q_all[s0:s0 + 131072], want_penult=False)[0]
                            for s0 in range(0, q_all.shape[0], 131072)])
        res['sharded_equals_single_rank_bitwise'] = bool(torch.equal(gathered, single))
        res['max_abs_diff_sharded_vs_single'] = float((gathered - single).abs().max())
    res['what'] = ('strong scaling: 2,097,152 random queries of ONE frame, contiguous shards over %d rank(s), encoder + '
                   'prepare_scene replicated and inside the timed region, one all_gather_into_tensor' % world)
    return res


def second_device_check(dev):
    """One process, two GPUs (the nn.DataParallel caller of train.py:305): a small decoder forward on another visible
    device after this rank's own must give the same bits (per-device kernel attributes, per-device workspaces)."""
    from tests import configs
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        return None
    other = torch.device('cuda', (dev.index + 1) % n_dev)
    cfg = configs.C1_GREATER
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_greater_seeded.npz'))
    outs = []
    for d in (dev, other):
        _, dec = configs.build_modules(cfg, d)
        with torch.cuda.device(d), torch.no_grad():
            o, _ = dec(torch.from_numpy(z['query']).to(d), torch.from_numpy(z['abstract']).to(d),
                       torch.from_numpy(z['glob']).to(d), None)
            torch.cuda.synchronize(d)
        outs.append(o.cpu())
    gold = torch.from_numpy(z['out'])
    return {'devices': [str(dev), str(other)], 'bit_identical': bool(torch.equal(outs[0], outs[1])),
            'max_rel_err_vs_reference_golden': float((outs[1] - gold).abs().max() / gold.abs().max())}


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


def reference_train_frame(dev, cfg, abstract, glob, query, target, weights, iters=5):
    """One decoder frame of the training step -- forward, fused loss heads, backward -- through the UNMODIFIED reference
    module (implicit.LocalPclResnetFC.forward under torch.autograd, eager fp32 and bf16 autocast as train.py:282-296 runs
    it) and through o4d on the same weights, inputs and loss; median of `iters` after 2 warm-ups.  The encoder is left
    out on both sides (the reference's needs torch_cluster, absent here).  None without the reference copy."""
    mods = reference_modules(cfg, dev)
    if mods is None:
        return None
    from o4d import loss as o4d_loss
    from tests import configs
    loader, _, rdec = mods
    _, odec = configs.build_modules(cfg, dev)
    rdec.train()
    odec.train()
    a0, g0 = abstract.detach(), glob.detach()

    def frame(dec, autocast):
        a, g = a0.clone().requires_grad_(True), g0.clone().requires_grad_(True)
        for p in dec.parameters():
            p.grad = None
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
            out, _ = dec(query, a, g, None)
        heads = o4d_loss.implicit_loss_heads(out.float().reshape(query.shape[0], -1), target, 'rgb', 13, True)
        (heads * weights).sum().backward()
        # feature columns only: the xyz columns are FPS-selected input coordinates with no parameter upstream, and the
        # o4d training path does not differentiate through them (autograd.decoder_train detaches them)
        return torch.cat([a.grad[:, 3:].reshape(-1), g.grad.reshape(-1)])

    def timed(dec, autocast):
        ts = []
        for i in range(iters + 2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ga = frame(dec, autocast)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        return _median(ts), ga

    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with loader.quiet():
            ref_ms, ref_ga = timed(rdec, False)
            ref_bf16_ms, _ = timed(rdec, True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    o4d_ms, o4d_ga = timed(odec, False)
    err = float((o4d_ga - ref_ga).norm() / ref_ga.norm())
    torch.cuda.empty_cache()
    return {'what': 'one decoder frame (%d query points, %d abstract points): forward + loss heads + backward' % (query.shape[0], a0.shape[0]),
            'reference_eager_fp32_ms': ref_ms, 'reference_eager_bf16_autocast_ms': ref_bf16_ms, 'o4d_ms': o4d_ms,
            'speedup_vs_reference_fp32': ref_ms / o4d_ms, 'speedup_vs_reference_bf16_autocast': ref_bf16_ms / o4d_ms,
            'abstract_feature_and_global_gradient_rel_l2_vs_reference_autograd': err, 'kind': 'reference (oracle/_ref copy of the unmodified modules)'}


def warm_up_until_quiet(step, alloc_count, world, dev, min_steps=3, max_steps=12):
    """Runs step() until one step makes at most one device allocation (alloc_count() = running number of cudaMalloc
    calls of this rank), at least min_steps and at most max_steps times; returns the number of steps run.  Every step
    contains a collective when world > 1 (the gradient all-reduce), so the decision to go on is itself agreed across
    the ranks (MAX): all ranks run the SAME number of steps -- ranks leaving the loop at different times deadlock
    the next collective (seen once on 2 GPUs: an all-reduce of the gradients against the barrier of the faster rank)."""
    import torch.distributed as dist
    n = 0
    while n < max_steps:
        before = alloc_count()
        step()
        n += 1
        more = 1 if (n < min_steps or alloc_count() - before > 1) else 0
        if world > 1:
            flag = torch.tensor([more], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            more = int(flag.item())
        if not more:
            break
    return n


def train_step_bench(dev, world, rank, steps):
    """BASELINE.json configs[4] per GPU: one CARLA-4D sample (14336 points), 4 frames x 17,203 query
    points (args.py:254,257 / train.py:270-274), forward + backward through the nn.Module API
    (o4d/autograd.py -> backward kernels), the fused loss heads (o4d.loss), gradient averaging over
    ranks, AdamW (the caller's torch optimizer, out of scope); the query sampler is excluded.  Timed twice: bf16x3
    operands (fp32-grade, what the parity tests check) and single-pass bf16 operands with fp32 accumulation and fp32
    master weights -- the "bf16" of BASELINE.json's config (the reference's mixed_precision = torch autocast).
    Returns a dict for the bench line."""
    import torch.distributed as dist
    from o4d import loss as o4d_loss, parallel
    from tests import configs
    cfg = configs.C3_CARLA
    frames, per_frame = 4, 17203
    g = torch.Generator().manual_seed(1830 + rank)
    pcl = configs.synthetic_cloud(cfg).to(dev)
    lo = torch.tensor([0.0, -16.0, -1.0])
    hi = torch.tensor([40.0, 16.0, 6.4])
    queries = []
    for f in range(frames):
        q = torch.rand(per_frame, 4, generator=g)
        q[:, :3] = q[:, :3] * (hi - lo) + lo
        q[:, 3] = float(f)
        queries.append(q.to(dev))
    # targets (B, n, 6) as pipeline.py:184 shapes them: density, RGB, mark_track, semantic tag
    target = torch.rand(frames, per_frame, 6, generator=g)
    target[..., 0] = (target[..., 0] > 0.5).float()
    target[..., 4] = 0.0
    target[..., 5] = torch.randint(0, 13, (frames, per_frame), generator=g).float()
    target = target.to(dev)
    weights = torch.tensor([1.0, 1.0, 0.6, 1.0], device=dev)        # color, density, segmentation, tracking

    each, diag, means = [], [], []

    def run(precision):
        # the inference legs before this one leave the caching allocator full of differently sized blocks; without this
        # the first precision timed pays cudaMalloc / cudaFree churn inside its steps (seen as 110-200 ms instead of 84)
        torch.cuda.empty_cache()
        enc, dec = configs.build_modules(cfg, dev)
        enc.train()
        dec.train()
        enc.o4d_precision = dec.o4d_precision = precision
        for m in list(enc.modules()) + list(dec.modules()):
            if hasattr(m, 'o4d_precision'):
                m.o4d_precision = precision
        params = list(enc.parameters()) + list(dec.parameters())
        opt = torch.optim.AdamW(params, lr=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            abstract, glob, _ = enc(pcl[None], False)
            total = 0.0
            for f in range(frames):
                out, _ = dec(queries[f], abstract[0], glob[0], None)
                heads = o4d_loss.implicit_loss_heads(out, target[f], 'rgb', 13, True)
                total = total + (heads * weights).sum()
            (total / frames).backward()
            if world > 1:
                parallel.allreduce_gradients([enc, dec])
            opt.step()
            return total

        # warm-up (workspaces, packed-weight cache, cold host pages) until the caching allocator has stopped growing:
        # after the inference legs it needs ~7 steps of ~20 cudaMalloc calls each to settle on this step's block sizes,
        # and a cudaMalloc of a multi-GB segment blocks the host for 100-300 ms (seen as single 150-420 ms steps)
        warm_steps = warm_up_until_quiet(step, lambda: torch.cuda.memory_stats(dev).get('num_device_alloc', 0), world, dev)
        gc.collect()                             # ... and no cycle-collector pass inside the timed steps either
        gc.disable()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        evs[0].record()
        host, allocs = [], []
        for i in range(steps):
            t_h = time.perf_counter()
            n_a = torch.cuda.memory_stats(dev).get('num_device_alloc', 0)
            total = step()
            evs[i + 1].record()
            host.append(round((time.perf_counter() - t_h) * 1e3, 2))
            allocs.append(int(torch.cuda.memory_stats(dev).get('num_device_alloc', 0) - n_a))
        torch.cuda.synchronize()
        diag.append({'warmup_steps': warm_steps, 'host_enqueue_ms_each_step': host, 'cudaMalloc_calls_each_step': allocs})
        # The MEDIAN step is reported: on a freshly started box single steps show host stalls of 100-300 ms (the host
        # blocked in the driver, not in this code: `host_enqueue_ms_each_step`; the second bench process on the same box
        # shows none).  Mean and every step's time are kept beside it.
        step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
        ms = float(np.median(step_ms))
        means.append(evs[0].elapsed_time(evs[steps]) / steps)
        each.append([round(x, 2) for x in step_ms])
        gc.enable()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        loss = float(total.detach()) / frames
        del opt, enc, dec
        torch.cuda.empty_cache()
        return ms, loss

    ms, loss = run(1)
    ms_bf16, loss_bf16 = run(2)
    ref_frame = None
    if rank == 0:
        try:
            with torch.no_grad():
                enc0, _ = configs.build_modules(cfg, dev)
                abstract0, glob0, _ = enc0(pcl[None], False)
            ref_frame = reference_train_frame(dev, cfg, abstract0[0], glob0[0], queries[0], target[0], weights)
        except Exception as exc:          # the baseline must never take the bench line down
            ref_frame = {'error': repr(exc)[:300]}
    extras = sampler_and_loss_timing(dev) if rank == 0 else None
    flop = 3.0 * frames * per_frame * 47.9e6                      # ~3x the forward, 9.9 TFLOP per sample
    peak, _ = measured_peak()
    return {'sampler_and_loss_heads': extras, 'reference_autograd_decoder_frame': ref_frame, 'ms_per_step': ms, 'samples_per_step': world, 'queries_per_sample': frames * per_frame,
            'points_per_sample': cfg['n_points'], 'query_grads_per_s': world * frames * per_frame / (ms / 1e3),
            'loss': loss, 'steps': steps, 'ms_each_step': each[0], 'ms_mean_step': means[0], 'statistic': 'median of the timed steps', 'diagnostics': diag[0], 'precision': 'bf16x3 (fp32-grade) forward and backward',
            'bf16': {'ms_per_step': ms_bf16, 'ms_each_step': each[1], 'ms_mean_step': means[1], 'diagnostics': diag[1], 'query_grads_per_s': world * frames * per_frame / (ms_bf16 / 1e3), 'loss': loss_bf16,
                     'precision': 'single-pass bf16 operands, fp32 accumulation, fp32 master weights (BASELINE config 5 "bf16")',
                     'roofline': {'bound': 'tensor', 'achieved': flop / (ms_bf16 / 1e3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                                  'frac': flop / (ms_bf16 / 1e3) / 1e12 / peak}},
            'roofline': {'bound': 'tensor', 'achieved': flop / (ms / 1e3) / 1e12, 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': flop / (ms / 1e3) / 1e12 / peak,
                         'note': 'algorithmic flops of the step (3 x forward, SURVEY 8d) over the whole step time'},
            'what': 'CARLA config 5 shape: encoder + 4 decoder frames forward/backward, fused loss heads, grad all-reduce, '
                    'AdamW; sampler excluded'}


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
        # the encoder's critical path: the serial arg-max chain of farthest point sampling (torch_cluster.fps semantics,
        # modules.py:133) over the level pyramid, timed on its own -- a latency bound, not a bandwidth / tensor one
        fps_chain = None
        try:
            sizes, nl = [], cfg['n_points']
            for _ in range(cfg['pcl_args']['down_blocks']):
                nn = -(-nl // cfg['pcl_args']['transition_factor'])
                sizes.append((nl, nn))
                nl = nn
            clouds = [pcl[:a, :3].contiguous() for a, _ in sizes]
            for c_, (a, b) in zip(clouds, sizes):
                ops.fps(c_, b, 0)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(3):
                for c_, (a, b) in zip(clouds, sizes):
                    ops.fps(c_, b, 0)
            f1.record()
            torch.cuda.synchronize()
            fps_ms = f0.elapsed_time(f1) / 3
            picks = sum(b for _, b in sizes)
            fps_chain = {'bound': 'latency (serial arg-max chain)', 'picks': picks, 'ms': fps_ms,
                         'us_per_pick': fps_ms * 1e3 / picks, 'share_of_encoder': fps_ms / enc_ms,
                         'levels': [list(x) for x in sizes]}
        except Exception as exc:               # noqa: BLE001
            fps_chain = {'error': '%s: %s' % (type(exc).__name__, exc)}
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

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()                  # nvidia-smi needs a few hundred ms to start: launch it before the warm-up
        for _ in range(args.warmup):
            step_device()
        barrier()
        sampler.mark_begin()
        launches0 = lib.o4d_launch_count()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            step_device()
        t1.record()
        barrier()
        sampler.mark_end()
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
            families = family_profile(lib, lambda: step_device(comm=False))     # rank 0 only: no collective here
            top = max(families, key=lambda k: families[k]['ms'])
            peak, which = measured_peak()
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
        # ---- BASELINE configs[3]: strong scaling of one 2 M-query frame (every rank takes part)
        strong = None
        if not args.no_strong_scaling:
            try:
                strong = strong_scaling(dev, world, rank, enc, dec, pcl, cfg)
            except Exception as exc:           # noqa: BLE001 -- side measurement
                strong = {'error': '%s: %s' % (type(exc).__name__, exc)}
        # ---- BASELINE configs[2]: CARLA-shape throughput (rank 0)
        carla = None
        if rank == 0 and not args.no_carla:
            try:
                carla = carla_config3(dev, lib, batch, max(1, min(args.steps, 3)))
            except Exception as exc:           # noqa: BLE001
                carla = {'error': '%s: %s' % (type(exc).__name__, exc)}
        second_dev = None
        if rank == 0 and world > 1:
            try:
                second_dev = second_device_check(dev)
            except Exception as exc:           # noqa: BLE001
                second_dev = {'error': '%s: %s' % (type(exc).__name__, exc)}
        if world > 1:
            dist.barrier()

    train = None
    if not args.no_train_step:
        torch.cuda.empty_cache()
        train = train_step_bench(dev, world, rank, max(1, min(args.steps, 5)))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, secs, threads, kind = cpu_reference_rate(args.cpu_sample)
        cpu_base = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': kind,
                    'sample': '%d of %d grid queries (evenly strided, mini-batches of 4096), decoder only, %s, torch '
                              'CPU fp32, %.1f s' % (args.cpu_sample, nq, 'unmodified reference module (oracle/_ref)'
                                                    if kind == 'reference' else 'oracle port', secs)}
    torch_gpu, enc_base = None, None
    if rank == 0 and world == 1 and not args.no_torch_gpu_baseline:
        try:                                   # reported baselines: never allowed to take the bench line down
            torch_gpu = reference_decoder_gpu(dev, cfg, abstract, glob, q_host, batch, value, out_dev)
        except Exception as exc:               # noqa: BLE001
            torch_gpu = {'error': '%s: %s' % (type(exc).__name__, exc)}
        try:
            torch.cuda.empty_cache()
            enc_base = reference_encoder_baselines(dev, cfg, pcl.cpu(), cpu=not args.no_cpu_baseline)
        except Exception as exc:               # noqa: BLE001
            enc_base = {'error': '%s: %s' % (type(exc).__name__, exc)}
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
                'encoder': {'pts_per_s': cfg['n_points'] / (enc_ms / 1e3), 'ms': enc_ms, 'n_points': cfg['n_points'],
                            'critical_path': fps_chain, 'reference': enc_base},
                'carla_config3': carla, 'strong_scaling': strong, 'second_device_in_process': second_dev,
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
    ap.add_argument('--no-strong-scaling', action='store_true', help='skip the 2 M-query strong-scaling frames')
    ap.add_argument('--no-carla', action='store_true', help='skip the CARLA config-3 throughput key')
    ap.add_argument('--no-torch-gpu-baseline', action='store_true',
                    help='skip the PyTorch-eager timing of the same decoder on the same GPU')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_o4d(args)


if __name__ == '__main__':
    main()
