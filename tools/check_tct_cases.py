"""Transposed kernel (SCFLOW_TC_T=1) against the pixels-as-rows kernel on the decoder's exact call shapes."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scflow_b200 as S  # noqa: E402

dev = 'cuda'
g = torch.Generator().manual_seed(1)
# name, segs [(buffer channels, coff, nch)], cout, kernel, stride, out stride, out coff, f32?
CASES = [
    ('flow0', [(16, 0, 16)], 128, (7, 1), 1, 128, 0, False),
    ('flow1', [(128, 0, 128)], 64, (3, 3), 1, 256, 192, False),
    ('out0', [(256, 0, 256)], 126, (3, 3), 1, 128, 0, False),
    ('me1', [(64, 0, 64)], 32, (3, 3), 1, 32, 0, False),
    ('ph0', [(128, 0, 128), (64, 0, 64), (32, 0, 32)], 128, (3, 3), 2, 128, 0, True),
    ('ph1', [(128, 0, 128)], 128, (3, 3), 2, 128, 0, True),
]
for b in (2, 32):
    for name, segs, cout, k, stride, ostride, ocoff, f32 in CASES:
        hw = 32 if name != 'ph1' else 16
        xs = [S.ops.split_nchw(torch.randn(b, st, hw, hw, generator=g).to(dev)) for st, _, _ in segs]
        cin = sum(n for _, _, n in segs)
        w = (torch.randn(cout, cin, *k, generator=g) / math.sqrt(cin * k[0] * k[1])).to(dev)
        bias = (0.1 * torch.randn(cout, generator=g)).to(dev)
        pw = S.ops.pack_conv_weight_tc([w])
        ho = (hw - 1) // stride + 1
        res = []
        for t in ('0', '2'):
            os.environ['SCFLOW_TC_T'] = t
            out = torch.full((2, b, ho, ho, ostride), 7.0, device=dev, dtype=torch.bfloat16)
            of = torch.full((b, ho, ho, ostride), 7.0, device=dev) if f32 else None
            S.ops.conv2d_tc([(x, c, n) for x, (_, c, n) in zip(xs, segs)], pw, bias, cout, k, act='relu', out_hl=None if f32 else out,
                            out_hl_coff=ocoff, out_f32=of, stride=stride, out_pad_writable=(name == 'out0'))
            torch.cuda.synchronize()
            res.append((of if f32 else out).float()[..., :ocoff + cout].clone())
        d = (res[0] - res[1]).abs()
        print(f'B={b:2d} {name:6s} max diff {d.max().item():.3e}  mismatching elements {(d > 1e-3).sum().item()}')
