"""Released-checkpoint accuracy of the decoder against the fp64 arbitration run, per precision variant
(tests/golden/_ckpt fixtures).  Usage (on a B200): python tools/ckpt_error.py
Variants: fp32 CUDA cores; bf16x3 (default); fused-attention logits contraction with one of the two correction
products dropped (experiment: is a 2-pass split inside the 1e-3 bar on REAL weights?)."""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'occlusions-4d_b200')]
import torch
import o4d
from o4d import _lib

lib = _lib.lib()
h = ctypes.CDLL(_lib.LIB_PATH)
res = {}
for which in ('greater', 'carla'):
    ck = torch.load(os.path.join(ROOT, 'tests', 'golden', '_ckpt', which + '_nets.pt'), map_location='cpu', weights_only=True)
    dec = o4d.LocalPclResnetFC(**ck['implicit_args'])
    dec.load_state_dict(ck['implicit_net'], strict=True)
    dec = dec.cuda().eval()
    q, a, g = ck['query'].cuda(), ck['abstract'].cuda(), ck['glob'].cuda()
    qbig = q.repeat(8, 1)
    row = {'ref_fp32_vs_fp64': float((ck['out'].double() - ck['out64']).abs().max() / ck['out64'].abs().max())}
    for name, prec, mode in (('fp32_cuda_cores', 0, 1), ('bf16x3_all_three_products', 1, 1), ('default_logits_W_bf16', 1, 2),
                             ('drop_hiddenlo_x_Whi', 1, 3), ('bf16_single_pass', 2, 1)):
        dec.o4d_precision = prec
        h.o4d_debug_set_fused_passes(mode)
        with torch.no_grad():
            out, pen = dec(q, a, g, None)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dec(qbig, a, g, None)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        e = float((out.cpu().double() - ck['out64']).abs().max() / ck['out64'].abs().max())
        ep = float((pen.cpu()[:, :16].double() - ck['penult64']).abs().max() / ck['penult64'].abs().max())
        row[name] = {'out_vs_fp64': e, 'penult_vs_fp64': ep, 'ms_65536_queries': dt * 1e3}
    h.o4d_debug_set_fused_passes(0)       # back to the library default
    res[which] = row
print(json.dumps(res, indent=1))
