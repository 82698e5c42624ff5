"""Times the pyramid lookup (B=32, 256x256 crops -> 32x32 queries per sample) for each kernel variant (SCFLOW_LOOKUP_SMEM)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    import scflow_b200 as S
    b = 32
    g = torch.Generator().manual_seed(1)
    pyr = [torch.randn(b * 1024, 1, 32 >> l, 32 >> l, generator=g).cuda() for l in range(4)]
    flow = (torch.randn(b, 32, 32, 2, generator=g) * 3.).cuda()
    out = torch.zeros(2, b, 32, 32, 328, device='cuda', dtype=torch.bfloat16)
    import ctypes as C
    from scflow_b200 import _lib
    lib = _lib.load()
    arr = (C.c_void_p * 4)(*[t.data_ptr() for t in pyr])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    ts = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.scf_corr_lookup_split(arr, 4, 4, flow.data_ptr(), None, out.data_ptr(), out[0].numel(), 328, b, 32, 32, _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f'SCFLOW_LOOKUP_SMEM={sys.argv[1]}: {sorted(ts[2:])[3]:.1f} us')
else:
    for v in ('2', '1', '0'):
        subprocess.check_call([sys.executable, __file__, v], env=dict(os.environ, SCFLOW_LOOKUP_SMEM=v))
