"""The north star's denominator: the reference's PyTorch eager formulation of the decoder on the SAME B200
(fp32, TF32 off as in the reference), next to libo4d.so on the same queries (SURVEY.md 8d "Reference timing
beside it" (1)).

The reference's nn.Modules cannot travel to the GPU box (/root/reference is absent there), so this times the
oracle port -- the same eager op sequence (full distance matrix + sort for every kNN, materialised (N, K, D)
gathers, one cuBLAS SGEMM per nn.Linear, chunked like eval/inference.py's mini-batches) -- with its two
CPU-only helpers swapped for device-agnostic ones.  Measurement tool only: nothing in the product imports it.

    python tools/time_torch_gpu.py [--batches 2] [--batch 32768] > gpurun_out/torch_gpu.json
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'occlusions-4d_b200'))


def device_knn(orc):
    """knn_indices of the oracle without its numpy sqrt / CPU allocations (timing only; the tie rule of the
    rooted distances is irrelevant here)."""
    def knn_indices(query_xyz, ref_xyz, k, sqrt=False, chunk=2048):
        idx, dist = [], []
        for s in range(0, query_xyz.shape[0], chunk):
            d2 = orc._pair_sqdist(query_xyz[s:s + chunk], ref_xyz)
            if sqrt:
                d2 = d2.sqrt()
            val, order = torch.sort(d2, dim=1, stable=True)
            idx.append(order[:, :k])
            dist.append(val[:, :k])
        return torch.cat(idx), torch.cat(dist)
    return knn_indices


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', type=int, default=2)
    ap.add_argument('--batch', type=int, default=32768)
    ap.add_argument('--device', default='cuda')
    args = ap.parse_args()
    from oracle import o4d_oracle as orc
    from tests import configs
    import bench
    dev = torch.device(args.device)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    orc.knn_indices = device_knn(orc)
    cfg = configs.C2_GREATER
    _, dec = configs.build_modules(cfg)
    sd = {k: v.detach().float().to(dev) for k, v in dec.state_dict().items()}
    abstract, glob = bench.golden_scene()
    abstract, glob = abstract.to(dev), glob.to(dev)
    q_all = configs.synthetic_queries(cfg)
    n = args.batches * args.batch
    q = q_all[torch.linspace(0, q_all.shape[0] - 1, n).long()].to(dev)

    def sync():
        if dev.type == 'cuda':
            torch.cuda.synchronize()

    def run_torch():
        with torch.no_grad():
            return orc.decoder_forward(sd, cfg['implicit_args'], q, abstract, glob, chunk=args.batch)[0]

    ref_out = run_torch()                                    # warm-up (cuBLAS handles, allocator)
    sync()
    t0 = time.perf_counter()
    run_torch()
    sync()
    t_torch = time.perf_counter() - t0
    line = {'what': 'decoder forward, GREATER config 2, %d queries in mini-batches of %d' % (n, args.batch),
            'torch_eager_fp32': {'queries_per_s': n / t_torch, 'seconds': t_torch, 'tf32': False,
                                 'peak_mem_gib': torch.cuda.max_memory_allocated() / 2 ** 30 if dev.type == 'cuda' else None}}
    if dev.type == 'cuda':
        dec = dec.to(dev).eval()
        with torch.no_grad():
            out = dec(q, abstract, glob, None)[0]            # warm-up + parity
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(0, n, args.batch):
                dec(q[s:s + args.batch], abstract, glob, None)
            e1.record()
            torch.cuda.synchronize()
        t_o4d = e0.elapsed_time(e1) / 1e3
        err = float((out.reshape(ref_out.shape) - ref_out).abs().max() / ref_out.abs().max())
        line['o4d'] = {'queries_per_s': n / t_o4d, 'seconds': t_o4d}
        line['speedup_vs_torch_eager_same_gpu'] = t_torch / t_o4d
        line['max_rel_err_vs_torch_eager'] = err
        line['gpu'] = torch.cuda.get_device_name(0)
    print(json.dumps(line))


if __name__ == '__main__':
    main()
