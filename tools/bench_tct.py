"""conv_tc_kernel (pixels as MMA rows) against conv_tct_kernel (weights as MMA rows, SCFLOW_TC_T=1) per layer shape, with
the transposed kernel's timing-experiment switches (SCFLOW_TCT_DBG).  CUDA-event timing, L2 flushed between launches."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scflow_b200 as S  # noqa: E402

LAYERS = [  # name, batch, size, cin, cout, kernel, stride
    ('enc64_3x3', 64, 128, 64, 64, (3, 3), 1), ('enc96_s2', 64, 128, 64, 96, (3, 3), 2), ('enc96_3x3', 64, 64, 96, 96, (3, 3), 1),
    ('enc128_s2', 64, 64, 96, 128, (3, 3), 2), ('enc128_3x3', 64, 32, 128, 128, (3, 3), 1), ('enc_1x1', 64, 32, 128, 128, (1, 1), 1),
    ('out_3x3', 32, 32, 256, 126, (3, 3), 1), ('flow2_3x3', 32, 32, 128, 64, (3, 3), 1),
]
VARIANTS = [('rows', {'SCFLOW_TC_T': '0'}), ('T', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '0', 'SCFLOW_TCT_BK': '64'}),
            ('T bk32', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '0', 'SCFLOW_TCT_BK': '32', 'SCFLOW_TCT_HALO': '0'}),
            ('T halo', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '0', 'SCFLOW_TCT_BK': '32', 'SCFLOW_TCT_HALO': '1'}),
            ('halo noepi', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '1', 'SCFLOW_TCT_BK': '32', 'SCFLOW_TCT_HALO': '1'}),
            ('bk32 noepi', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '1', 'SCFLOW_TCT_BK': '32', 'SCFLOW_TCT_HALO': '0'}),
            ('T noepi', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '1', 'SCFLOW_TCT_BK': '64'}), ('T nomma', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '3'}),
            ('T noload', {'SCFLOW_TC_T': '2', 'SCFLOW_TCT_DBG': '13'})]
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator().manual_seed(0)
reps = 10
print(f'{"layer":12s} ' + ' '.join(f'{n:>11s}' for n, _ in VARIANTS) + '   (us)')
for name, b, hw, cin, cout, k, stride in LAYERS:
    x = torch.randn(b, cin, hw, hw, generator=g).to(dev)
    w = (torch.randn(cout, cin, *k, generator=g) / math.sqrt(cin * k[0] * k[1])).to(dev)
    xs = S.ops.split_nchw(x)
    pw = S.ops.pack_conv_weight_tc([w])
    ho = hw // stride
    out = torch.zeros(2, b, ho, ho, (cout + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
    of = torch.zeros(b, ho, ho, cout, device=dev)
    ref = None
    cols = []
    for vn, env in VARIANTS:
        os.environ.update(env)
        fn = lambda: S.ops.conv2d_tc([(xs, 0, cin)], pw, None, cout, k, act='relu', out_hl=out, out_f32=of, stride=stride)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tag = ''
        if vn in ('rows', 'T', 'T bk32', 'T halo'):
            cur = (S.ops.unsplit(out).clone(), of.clone())
            if ref is None:
                ref = cur
            else:
                err = max((cur[0] - ref[0]).abs().max().item(), (cur[1] - ref[1]).abs().max().item())
                tag = '' if err < 1e-4 else f' ERR {err:.2e}'
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for s, e in evs:
            flush.zero_()
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        us = 1e3 * sum(s.elapsed_time(e) for s, e in evs) / reps
        cols.append(f'{us:11.1f}{tag}')
    print(f'{name:12s} ' + ' '.join(cols))
