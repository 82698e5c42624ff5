"""Two representative conv_tct_kernel launches for `ncu --set full -k regex:conv_tct`: the motion-encoder output convolution
(B=32, 32x32, 256 -> 126, 3x3, split-bf16 output, halo form) and an encoder layer (B=64, 64x64, 96 -> 96, 3x3, fp32 output
with InstanceNorm partial sums).

    ncu --set full --clock-control none --import-source on -k regex:conv_tct -s 4 -c 2 -o gpurun_out/tct python tools/profile_tct.py
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import scflow_b200 as S  # noqa: E402

dev = 'cuda'
g = torch.Generator().manual_seed(0)


def layer(b, hw, cin, cout, stats):
    x = torch.randn(b, cin, hw, hw, generator=g).to(dev)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)).to(dev)
    xs = S.ops.split_nchw(x)
    pw = S.ops.pack_conv_weight_tc([w])
    if stats:
        of = torch.zeros(b, hw, hw, cout, device=dev)
        n_tiles, _ = S.ops.conv2d_tc_tiles(b, hw, hw)
        st = torch.zeros(n_tiles * 4 * 2 * cout, device=dev)
        return lambda: S.ops.conv2d_tc([(xs, 0, cin)], pw, None, cout, 3, act='none', out_f32=of, stats=st)
    out = torch.zeros(2, b, hw, hw, 128, device=dev, dtype=torch.bfloat16)
    return lambda: S.ops.conv2d_tc([(xs, 0, cin)], pw, None, cout, 3, act='relu', out_hl=out, out_pad_writable=True)


a = layer(32, 32, 256, 126, False)
e = layer(64, 64, 96, 96, True)
for _ in range(3):          # launches 0-5: warm-up; ncu -s 4 -c 2 captures the third pair
    a()
    e()
torch.cuda.synchronize()
print('done')
