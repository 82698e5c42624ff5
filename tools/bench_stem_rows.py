"""Times the x-folded stem convolution (7x1, vertical stride 2, 32 -> 64) on the rolling-rows kernel and on the generic tile."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scflow_b200 as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
gen = torch.Generator().manual_seed(0)
xs = S.ops.split_nchw(torch.randn(n, 32, 256, 128, generator=gen).cuda())
pw = S.ops.pack_conv_weight_tc([(torch.randn(64, 32, 7, 1, generator=gen) / math.sqrt(224)).cuda()])
bias = torch.zeros(64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
n_tiles, _ = S.ops.conv2d_tc_tiles(n, 128, 128)
for variant in ('in_raw', 'bn_relu'):
    for mode in ('1', '0'):
        os.environ['SCFLOW_TC_ROWS'] = mode
        out_f32 = torch.empty(n, 128, 128, 64, device='cuda')
        out_hl = torch.empty(2, n, 128, 128, 64, device='cuda', dtype=torch.bfloat16)
        st = torch.zeros(n_tiles * 4 * 2 * 64, device='cuda')
        kw = dict(in_raw=dict(out_f32=out_f32, stats=st), bn_relu=dict(act='relu', out_f32=out_f32, out_hl=out_hl))[variant]
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            S.ops.conv2d_tc([(xs, 0, 32)], pw, bias, 64, (7, 1), stride_xy=(1, 2), **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f'stem rows={mode} {variant:8s} N={n}: {sorted(ts[1:])[2]:7.1f} us')

if os.environ.get('TRACE'):
    os.environ['SCFLOW_TC_ROWS'] = '1'
    tb = torch.zeros(148 * 8, dtype=torch.int64, device='cuda')
    os.environ['SCFLOW_ROWS_DBG_TIMES'] = hex(tb.data_ptr())
    out_f32 = torch.empty(n, 128, 128, 64, device='cuda')
    st = torch.zeros(n_tiles * 4 * 2 * 64, device='cuda')
    flush.zero_()
    S.ops.conv2d_tc([(xs, 0, 32)], pw, bias, 64, (7, 1), stride_xy=(1, 2), out_f32=out_f32, stats=st)
    torch.cuda.synchronize()
    del os.environ['SCFLOW_ROWS_DBG_TIMES']
    t = tb.view(148, 8).double().mean(0) / 1.965e3            # cycles -> us at 1965 MHz
    names = ['schedule arithmetic', 'wait TMEM slots', 'wait operand tile', 'descriptors + MMA issue', 'commits', 'MMA thread total']
    for k, nme in enumerate(names):
        print(f'  MMA thread: {nme:26s} {float(t[k]):7.1f} us')
