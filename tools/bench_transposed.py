"""Orientation experiment for small-Cout layers: a 1x1 convolution 256 -> COUT over 32768 pixels computed (a) the usual way
(pixels = MMA rows, M = 128 pixels, N = COUT output channels) and (b) transposed through the same kernel (weights = MMA rows,
M = COUT padded to 128, N = 256 pixels per tile: the pixel tensor is passed as the 'weight' operand).  Same FLOPs."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S
dev = 'cuda'
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K, NPIX = int(os.environ.get("K", "2048")), 32768


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for s, e in evs:
        flush.zero_(); s.record(); fn(); e.record()
    torch.cuda.synchronize()
    return 1e3 * sum(s.elapsed_time(e) for s, e in evs) / 10


for cout in (64, 128):
    x = torch.randn(32, K, 32, 32, generator=g).to(dev)                 # 32768 pixels
    w = (torch.randn(cout, K, 1, 1, generator=g) / math.sqrt(K)).to(dev)
    xs = S.ops.split_nchw(x)                                            # [2, 32, 32, 32, K]
    pw = S.ops.pack_conv_weight_tc([w])                                 # [2, 1, cout_pad, K]
    out_a = torch.empty(32, 32, 32, cout, device=dev)
    ta = timeit(lambda: S.ops.conv2d_tc([(xs, 0, K)], pw, None, cout, (1, 1), out_f32=out_a))
    # transposed: "activation" = weights as 128 pixels (zero rows beyond cout), "weight" = the pixel tensor [2][1][32768][K]
    wt = torch.zeros(2, 1, 1, 128, K, device=dev, dtype=torch.bfloat16)
    wt[:, 0, 0, :cout] = pw[:, 0, :cout]
    px = xs.reshape(2, 1, NPIX, K)
    out_b = torch.empty(1, 1, 128, NPIX, device=dev)
    tb = timeit(lambda: S.ops.conv2d_tc([(wt, 0, K)], px, None, NPIX, (1, 1), out_f32=out_b))
    ref = out_a.reshape(NPIX, cout)
    got = out_b.reshape(128, NPIX)[:cout].t()
    err = float((ref - got).abs().max())
    fl = 2.0 * NPIX * cout * K
    print(f'cout={cout}: pixels-as-rows {ta:.1f} us ({fl / ta / 1e6:.0f} TFLOP/s alg) | weights-as-rows {tb:.1f} us ({fl / tb / 1e6:.0f} TFLOP/s alg) | max diff {err:.2e}')
