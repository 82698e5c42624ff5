"""Per-CTA timeline of conv_tc_kernel via globaltimer stamps (timing experiment; SCFLOW_TC_DBG_TIMES)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import scflow_b200 as S

dev = 'cuda'
g = torch.Generator().manual_seed(0)
for name, cin, cout, k, b, hw in [('gru_zr_1x5', 384, 256, (1, 5), 32, 32), ('flow1_3x3', 128, 64, (3, 3), 32, 32),
                                  ('enc64_3x3', 64, 64, (3, 3), 8, 128), ('mhp_1x1', 256, 1, (1, 1), 32, 32)]:
    x = torch.randn(b, cin, hw, hw, generator=g).to(dev)
    w = (torch.randn(cout, cin, *k, generator=g) / math.sqrt(cin * k[0] * k[1])).to(dev)
    xs = S.ops.split_nchw(x)
    pw = S.ops.pack_conv_weight_tc([w])
    out = torch.zeros(2, b, hw, hw, (cout + 7) // 8 * 8, device=dev, dtype=torch.bfloat16)
    ncta = (b * hw * hw // 128) * ((cout + 255) // 256)
    times = torch.zeros(ncta, 8, dtype=torch.int64, device=dev)
    os.environ['SCFLOW_TC_CLUSTER'] = os.environ.get('SCFLOW_TC_CLUSTER', '1')
    fn = lambda: S.ops.conv2d_tc([(xs, 0, cin)], pw, None, cout, k, act='relu', out_hl=out)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    os.environ['SCFLOW_TC_DBG_TIMES'] = hex(times.data_ptr())
    fn()
    torch.cuda.synchronize()
    del os.environ['SCFLOW_TC_DBG_TIMES']
    t = times.cpu().double()
    t = t[t[:, 0] > 0]          # persistent kernel: only min(tiles, resident CTAs) rows are written (first tile of each CTA)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3   # us
    print(f'== {name}: kernel span {float((t[:, 6].max() - t0) / 1e3):.1f} us')
    names = ['start', 'prologue done', 'first data', 'last mma issued', 'accum ready', 'epilogue done', 'exit']
    order = torch.argsort(rel[:, 0])
    for label, idx in (('earliest CTA', order[0]), ('median CTA', order[len(order) // 2]), ('last CTA', order[-1])):
        r = rel[idx]
        print(f'  {label:28s} ' + '  '.join(f'{n}={float(r[i]):7.2f}' for i, n in enumerate(names)))
    dur = rel[:, 6] - rel[:, 0]
    print(f'  CTA lifetime us: mean {float(dur.mean()):.2f} min {float(dur.min()):.2f} max {float(dur.max()):.2f};  '
          f'prologue {float((rel[:,1]-rel[:,0]).mean()):.2f}  wait-first-data {float((rel[:,2]-rel[:,1]).mean()):.2f}  '
          f'mainloop {float((rel[:,4]-rel[:,2]).mean()):.2f}  epilogue {float((rel[:,5]-rel[:,4]).mean()):.2f}  teardown {float((rel[:,6]-rel[:,5]).mean()):.2f}')
    print(f'  ({t.shape[0]} CTAs; stamps 1-5 = first tile, exit = whole CTA) start-time quantiles us: ' + ' '.join(f'{float(q):.1f}' for q in torch.quantile(rel[:, 0], torch.tensor([0., .25, .5, .57, .6, .75, 1.]).double())))
