"""Times the pose head's FC tail: tcgen05 split-K layers (scf_linear_tc) against the CUDA-core kernel (scf_linear)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scflow_b200 as S  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
gen = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(b, 2048, generator=gen)).cuda()
w0, b0 = (torch.randn(1024, 2048, generator=gen) / 45).cuda(), torch.zeros(1024).cuda()
w1, b1 = (torch.randn(256, 1024, generator=gen) / 32).cuda(), torch.zeros(256).cuda()
p0w = S.ops.pack_conv_weight_tc([w0.reshape(1024, 2048, 1, 1)], cin_pad=2048)
p1w = S.ops.pack_conv_weight_tc([w1.reshape(256, 1024, 1, 1)], cin_pad=1024)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timeit(fn, name, warm_l2):
    ts = []
    for i in range(8):
        if not warm_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f'{name:34s} {"warm" if warm_l2 else "cold"} L2: {sorted(ts[2:])[len(ts[2:]) // 2]:6.1f} us')


for warm in (True, False):
    part0 = S.ops.linear_tc(x, p0w, kr=256)
    timeit(lambda: S.ops.linear_tc(x, p0w, kr=256), 'fc0 tcgen05 KR=256 (64 blocks)', warm)
    timeit(lambda: S.ops.linear_tc(x, p0w, kr=128), 'fc0 tcgen05 KR=128 (128 blocks)', warm)
    timeit(lambda: S.ops.linear_tc(part0, p1w, kr=128, x_bias=b0, x_relu=True), 'fc1 tcgen05 KR=128 (16 blocks)', warm)
    timeit(lambda: S.ops.linear_tc(part0, p1w, kr=64, x_bias=b0, x_relu=True), 'fc1 tcgen05 KR=64 (32 blocks)', warm)
    y0 = S.ops.linear(x, w0, b0, 'relu')
    timeit(lambda: S.ops.linear(x, w0, b0, 'relu'), 'fc0 CUDA cores', warm)
    timeit(lambda: S.ops.linear(y0, w1, b1, 'relu'), 'fc1 CUDA cores', warm)
