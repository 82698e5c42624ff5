"""Micro-benchmark of the folded 7x1 stem convolution (x-im2col'd RGB image, vertical stride 2) for different channel
paddings of the folded tensor.  Usage: python tools/bench_stem.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S

dev = 'cuda'
n = int(os.environ.get('N', '64'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
for kc, sy in [(24, 2), (32, 2), (64, 2), (24, 1)]:
    t = torch.randn(2, n, 256, 128, kc, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(64, kc, 7, 1, generator=g) / math.sqrt(7 * kc)).to(dev)
    pw = S.ops.pack_conv_weight_tc([w])
    ho = 128 if sy == 2 else 256
    out = torch.empty(n, ho, 128, 64, device=dev)
    fn = lambda: S.ops.conv2d_tc([(t, 0, kc)], pw, None, 64, (7, 1), act='none', out_f32=out, stride_xy=(1, sy))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for s, e in evs:
        flush.zero_(); s.record(); fn(); e.record()
    torch.cuda.synchronize()
    us = 1e3 * sum(s.elapsed_time(e) for s, e in evs) / 5
    print(f'kc={kc:3d} sy={sy}: {us:8.1f} us')
