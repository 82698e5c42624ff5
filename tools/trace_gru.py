"""Per-CTA timeline (globaltimer stamps, SCFLOW_TC_DBG_TIMES) and CUDA-event timing of the two GRU convolutions exactly
as scf_decoder_forward launches them (B=32, 32x32 maps): z|r (N=256, K=5*256, context term added in the epilogue) and q
(N=128, stacked-N).  Usage: python tools/trace_gru.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
from scflow_b200 import _lib

dev = 'cuda'
b = int(os.environ.get('B', '32'))
g = torch.Generator().manual_seed(0)
h = torch.tanh(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
mot = torch.randn(b, 32, 32, 128, generator=g).to(dev)
z = torch.rand(b, 32, 32, 128, generator=g).to(dev)
hs = S.ops.split_nchw(h.permute(0, 3, 1, 2).contiguous())
ms = S.ops.split_nchw(mot.permute(0, 3, 1, 2).contiguous())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, cout, epi, act):
    w = (torch.randn(cout, 256, 1, 5, generator=g) / math.sqrt(1280)).to(dev)
    pw = S.ops.pack_conv_weight_tc([w])
    pre = torch.randn(b, 32, 32, cout, generator=g).to(dev)
    zo = torch.empty(b, 32, 32, 128, device=dev)
    rhs = torch.zeros(2, b, 32, 32, 128, device=dev, dtype=torch.bfloat16)
    hn = torch.zeros(2, b, 32, 32, 128, device=dev, dtype=torch.bfloat16)
    if epi == _lib.EPI_GRU_ZR:
        fn = lambda: S.ops.conv2d_tc([(hs, 0, 128), (ms, 0, 128)], pw, None, cout, (1, 5), act=act, out_f32=zo, epi=epi, aux0=h,
                                     out2_hl=rhs, pre=pre)
    else:
        fn = lambda: S.ops.conv2d_tc([(hs, 0, 128), (ms, 0, 128)], pw, None, cout, (1, 5), act=act, out_f32=zo, out_hl=hn, epi=epi,
                                     aux0=h, aux1=z, pre=pre)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for s, e in evs:
        flush.zero_(); s.record(); fn(); e.record()
    torch.cuda.synchronize()
    us = 1e3 * sum(s.elapsed_time(e) for s, e in evs) / 10
    times = torch.zeros(1024, 8, dtype=torch.int64, device=dev)
    os.environ['SCFLOW_TC_DBG_TIMES'] = hex(times.data_ptr())
    flush.zero_()
    fn()
    torch.cuda.synchronize()
    del os.environ['SCFLOW_TC_DBG_TIMES']
    t = times.cpu().double()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    span = float((t[:, 6].max() - t0) / 1e3)
    ncta = t.shape[0]
    t = t[t[:, 2] > 0]          # CTA-pair mode: only the leader CTA issues MMAs and stamps the main loop
    rel = (t - t0) / 1e3
    tf = 2.0 * b * 1024 * cout * 1280 / (us * 1e-6) / 1e12
    print(f'== {name}: {us:.1f} us by CUDA events ({tf:.0f} TFLOP/s algorithmic); span by stamps {span:.1f} us; {ncta} CTAs ({t.shape[0]} issuing)')
    print(f'   first tile of each CTA: prologue {float((rel[:,1]-rel[:,0]).mean()):.2f}  wait-first-data {float((rel[:,2]-rel[:,1]).mean()):.2f}  '
          f'mainloop {float((rel[:,4]-rel[:,2]).mean()):.2f}  epilogue {float((rel[:,5]-rel[:,4]).mean()):.2f} (max {float((rel[:,5]-rel[:,4]).max()):.2f})  '
          f'CTA lifetime mean {float((rel[:,6]-rel[:,0]).mean()):.2f} max {float((rel[:,6]-rel[:,0]).max()):.2f}')


run('gru z|r 1x5 N=256 K=1280', 256, _lib.EPI_GRU_ZR, 'sigmoid')
run('gru q   1x5 N=128 K=1280', 128, _lib.EPI_GRU_Q, 'tanh')
