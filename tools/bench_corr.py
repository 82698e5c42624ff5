"""Times the correlation build (+ pyramid) alone: fused kernel vs GEMM + three pooling launches (CUDA events, L2 flushed)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scflow_b200 as S
from oracle import scflow_oracle as O

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
f = O.make_features(1, b, 32, 32)
fr, fe = f['feat_render'].cuda(), f['feat_real'].cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
import torch.cuda
lvls = None
def kernel_only():
    """the fused kernel alone through the decoder's presplit entry is not exposed; time corr_build minus the layout launches"""
for mode in (('1', '0') if not os.environ.get('ONLY_FUSED') else ('1',)):
    os.environ['SCFLOW_CORR_FUSED'] = mode
    for _ in range(3):
        S.ops.corr_build(fr, fe, 4, precision=1)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for s, e in evs:
        flush.zero_()
        s.record(); S.ops.corr_build(fr, fe, 4, precision=1); e.record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / 10
    mb = b * (2.10 + 5.57)
    print(f'corr_build fused={mode} B={b}: {ms * 1e3:.1f} us incl. the two layout-change launches; pyramid bytes {mb:.0f} MB -> {mb / ms / 1e3:.2f} TB/s')
