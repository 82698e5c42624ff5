"""MMA issue-rate experiment: main-loop time of a 3x3 convolution 128 -> COUT (B=32, 32x32) with ALL TMA loads skipped
(SCFLOW_TC_DBG_EPI=24: results are garbage, timing only), i.e. the pure tcgen05.mma rate for different N."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
dev = 'cuda'
g = torch.Generator().manual_seed(0)
b = 32
x = torch.randn(b, 128, 32, 32, generator=g).to(dev)
xs = S.ops.split_nchw(x)
for cout in (16, 64, 128, 256):
    w = (torch.randn(cout, 128, 3, 3, generator=g) / math.sqrt(128 * 9)).to(dev)
    pw = S.ops.pack_conv_weight_tc([w])
    out = torch.zeros(2, b, 32, 32, cout, device=dev, dtype=torch.bfloat16)
    fn = lambda: S.ops.conv2d_tc([(xs, 0, 128)], pw, None, cout, (3, 3), act='relu', out_hl=out)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    times = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
    os.environ['SCFLOW_TC_DBG_TIMES'] = hex(times.data_ptr())
    fn(); torch.cuda.synchronize()
    del os.environ['SCFLOW_TC_DBG_TIMES']
    t = times.cpu().double(); t = t[t[:, 2] > 0]; rel = (t - t[:, 0].min()) / 1e3
    ml = float((rel[:, 3] - rel[:, 2]).mean())       # first data -> last MMA issued
    ml2 = float((rel[:, 4] - rel[:, 2]).mean())      # first data -> accumulator ready
    stack = os.environ.get('SCFLOW_TC_STACKN', '0') != '0' and cout <= 128
    nmma = 72 * (2 if stack else 3)
    print(f'cout={cout:3d} stackn={int(stack)}: issue loop {ml:.2f} us, until accumulator ready {ml2:.2f} us -> {ml2 * 1e3 / nmma:.1f} ns per MMA ({nmma} MMAs)')
