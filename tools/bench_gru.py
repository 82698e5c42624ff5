"""Times the one-kernel SepConvGRU pass alone (CUDA events, L2 flushed between launches) and prints its tensor-pipe fraction."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scflow_b200 as S

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(1)
h = torch.tanh(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
cxt = torch.relu(torch.randn(b, 32, 32, 128, generator=g)).to(dev)
mot = torch.randn(b, 32, 32, 128, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for kernel in ((1, 5), (5, 1)):
    ws = [(torch.randn(128, 384, *kernel, generator=g) * 0.02).to(dev) for _ in range(3)]
    bs = [torch.zeros(128, device=dev) for _ in range(3)]
    op = S.ops.GruPassFused(*ws, *bs)
    op.precompute(cxt)
    hs = S.ops.split_nchw(h.permute(0, 3, 1, 2).contiguous())
    ms_ = S.ops.split_nchw(mot.permute(0, 3, 1, 2).contiguous())
    from scflow_b200 import _lib
    import ctypes as C
    out = torch.empty_like(h)
    out_hl = torch.empty(2, b, 32, 32, 128, device=dev, dtype=torch.bfloat16)
    d = _lib.GruPassDesc()
    d.h_hl, d.h_plane, d.h_f32 = hs.data_ptr(), hs[0].numel(), h.data_ptr()
    d.m_hl, d.m_plane = ms_.data_ptr(), ms_[0].numel()
    d.w_zr, d.w_q = op.w_zr.data_ptr(), op.w_q.data_ptr()
    d.pre_zr, d.pre_q = op.pre_zr.data_ptr(), op.pre_q.data_ptr()
    d.out_f32, d.out_hl, d.out_plane = out.data_ptr(), out_hl.data_ptr(), out_hl[0].numel()
    d.B, d.H, d.W, d.vertical = b, 32, 32, op.vertical
    lib = _lib.load()
    launch = lambda: _lib.check(lib.scf_gru_pass_fused(C.byref(d), _lib.stream_ptr()))
    for _ in range(3):
        launch()
    if os.environ.get('TRACE'):
        tb = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
        os.environ['SCFLOW_GRU_DBG_TIMES'] = hex(tb.data_ptr())
        flush.zero_(); launch(); torch.cuda.synchronize()
        del os.environ['SCFLOW_GRU_DBG_TIMES']
        t = tb.view(148, 16)[:min(148, b * 4)].double()
        t0 = t[:, 0:1]
        names = ['start', 'setup done', 'MMA: zr committed', 'EPI: zr_full seen', 'MMA: r.h ready seen', 'MMA: 2b (r.h) issued', 'MMA: 2a (motion) issued', '-', '-', '-', 'EPI: q_full seen', 'EPI: tile done']
        rel = (t - t0) / 1e3
        for i, n in enumerate(names):
            print(f'  {n:24s} mean {float(rel[:, i].mean()):7.2f} us  min {float(rel[:, i].min()):7.2f}  max {float(rel[:, i].max()):7.2f}')
        print(f'  CTA start spread {float((t0.max() - t0.min()) / 1e3):.2f} us')
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in evs:
        flush.zero_()
        s.record(); launch(); e.record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in evs) / reps
    flops = 2.0 * b * 1024 * 384 * 1280          # z, r, q over [h | motion], 5 taps
    print(f'gru_pass_kernel {kernel} B={b}: {ms * 1e3:.1f} us, {flops / ms / 1e9:.0f} TFLOP/s algorithmic, dbg={os.environ.get("SCFLOW_GRU_DBG", "0")}')
