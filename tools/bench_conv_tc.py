"""Micro-benchmark of conv_tc_kernel on the decoder's layer shapes (B=32, 32x32 maps), per cluster size.
CUDA-event timing, L2 flushed between launches. Usage: python tools/bench_conv_tc.py [--clusters 1,2,4,8]"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scflow_b200 as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--clusters', default='1,2,4,8')
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--debug', default='0')
args = ap.parse_args()
os.environ['SCFLOW_TC_DEBUG'] = args.debug

LAYERS = [  # name, cin, cout, kernel
    ('gru_zr_1x5', 384, 256, (1, 5)), ('gru_q_5x1', 384, 128, (5, 1)), ('heads_3x3', 128, 512, (3, 3)),
    ('corr1_3x3', 256, 192, (3, 3)), ('corr0_1x1', 328, 256, (1, 1)), ('out_3x3', 256, 126, (3, 3)),
    ('flow1_3x3', 128, 64, (3, 3)), ('fhp_3x3', 256, 2, (3, 3)),
]
dev = 'cuda'
b = args.batch
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
print(f'{"layer":12s} ' + ' '.join(f'cl={c:>2s}: us (TF/s alg)' for c in args.clusters.split(',')))
for name, cin, cout, k in LAYERS:
    x = torch.randn(b, cin, 32, 32, generator=g).to(dev)
    w = (torch.randn(cout, cin, *k, generator=g) / math.sqrt(cin * k[0] * k[1])).to(dev)
    xs = S.ops.split_nchw(x)
    pw = S.ops.pack_conv_weight_tc([w])
    out = torch.zeros(2, b, 32, 32, (cout + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
    ref = None
    cols = []
    for cl in args.clusters.split(','):
        os.environ['SCFLOW_TC_CLUSTER'] = cl
        fn = lambda: S.ops.conv2d_tc([(xs, 0, cin)], pw, None, cout, k, act='relu', out_hl=out)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        cur = S.ops.unsplit(out).clone()
        if ref is None:
            ref = cur
        ok = torch.equal(cur, ref)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for s, e in evs:
            flush.zero_()
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        us = 1e3 * sum(s.elapsed_time(e) for s, e in evs) / args.reps
        tf = 2.0 * b * 1024 * cout * cin * k[0] * k[1] / (us * 1e-6) / 1e12
        cols.append(f'{us:8.1f} ({tf:6.1f}){"" if ok else " MISMATCH"}')
    print(f'{name:12s} ' + '  '.join(cols))
