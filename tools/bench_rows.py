"""Times the 64 -> 64 3x3 encoder convolution (N images of 128x128) on the rolling-rows kernel and on the generic tile.
   python tools/bench_rows.py [N]      env SCFLOW_ROWS_DBG: 1 no MMAs, 2 no stores, 4 no activation loads"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scflow_b200 as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = (128, 128)
gen = torch.Generator().manual_seed(0)
x = torch.randn(n, 64, *hw, generator=gen).cuda()
w = (torch.randn(64, 64, 3, 3, generator=gen) / math.sqrt(576)).cuda()
xs = S.ops.split_nchw(x)
pw = S.ops.pack_conv_weight_tc([w])
bias = torch.zeros(64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
n_tiles, _ = S.ops.conv2d_tc_tiles(n, *hw)
flops = 2.0 * n * hw[0] * hw[1] * 64 * 64 * 9


def run(mode, variant):
    os.environ['SCFLOW_TC_ROWS'] = mode
    out_f32 = torch.empty(n, *hw, 64, device='cuda')
    out_hl = torch.empty(2, n, *hw, 64, device='cuda', dtype=torch.bfloat16)
    st = torch.zeros(n_tiles * 4 * 2 * 64, device='cuda')
    kw = dict(in_raw=dict(out_f32=out_f32, stats=st), bn_relu=dict(act='relu', out_hl=out_hl),
              bn_res=dict(act='relu', out_f32=out_f32, out_hl=out_hl, aux0=x.permute(0, 2, 3, 1).contiguous()),
              bn_res_hl=dict(act='relu', out_hl=out_hl, aux0_hl=xs))[variant]
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S.ops.conv2d_tc([(xs, 0, 64)], pw, bias, 64, 3, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts[1:])[len(ts[1:]) // 2]
    print(f'rows={mode} {variant:8s} N={n}: {t:7.1f} us  {flops / t * 1e-6:6.1f} TFLOP/s algorithmic')


for variant in ('in_raw', 'bn_relu', 'bn_res', 'bn_res_hl'):
    for mode in ('1', '0'):
        if variant == 'bn_res_hl' and mode == '0':
            continue                      # the split-bf16 residual exists on the rolling-rows kernel only
        run(mode, variant)
