"""Per-CTA timeline of a small-N convolution (3x3 128->64, B=32, 32x32: the flow_net.1 / delta_flow_encoder.1 shape)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
dev = 'cuda'
g = torch.Generator().manual_seed(0)
b = 32
x = torch.randn(b, 128, 32, 32, generator=g).to(dev)
w = (torch.randn(64, 128, 3, 3, generator=g) / math.sqrt(128 * 9)).to(dev)
xs = S.ops.split_nchw(x)
pw = S.ops.pack_conv_weight_tc([w])
out = torch.zeros(2, b, 32, 32, 64, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
fn = lambda: S.ops.conv2d_tc([(xs, 0, 128)], pw, None, 64, (3, 3), act='relu', out_hl=out)
for _ in range(3):
    fn()
torch.cuda.synchronize()
ref = S.ops.unsplit(out).clone()
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
for s, e in evs:
    flush.zero_(); s.record(); fn(); e.record()
torch.cuda.synchronize()
us = 1e3 * sum(s.elapsed_time(e) for s, e in evs) / 10
times = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
os.environ['SCFLOW_TC_DBG_TIMES'] = hex(times.data_ptr())
flush.zero_(); fn(); torch.cuda.synchronize()
t = times.cpu().double(); t = t[t[:, 2] > 0]; rel = (t - t[:, 0].min()) / 1e3
print(f'3x3 128->64: {us:.1f} us; first tile: wait-first-data {float((rel[:,2]-rel[:,1]).mean()):.2f} mainloop {float((rel[:,4]-rel[:,2]).mean()):.2f} '
      f'epilogue {float((rel[:,5]-rel[:,4]).mean()):.2f} lifetime {float((rel[:,6]-rel[:,0]).mean()):.2f}; checksum {float(ref.double().abs().sum()):.6f}')
