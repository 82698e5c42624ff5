"""Bit-compares the staged (shared-memory) and the direct lookup kernels on seeded inputs (run once per variant, the
kernel choice is read from SCFLOW_LOOKUP_SMEM at first use).  Usage: python tools/check_lookup_variants.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import torch
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    f = O.make_features(3, 4)
    pyr = S.ops.corr_build(f['feat_render'].cuda(), f['feat_real'].cuda(), 4, 1)
    g = torch.Generator().manual_seed(5)
    flow = (torch.randn(4, 32, 32, 2, generator=g) * 6.).cuda()
    flow[0, 0, 0] = 1e6; flow[0, 0, 1] = -1e6          # far outside the map
    out = S.ops.corr_lookup_nhwc(pyr, flow, 4)
    torch.save(out.cpu(), sys.argv[1])
else:
    outs = []
    for v in ('0', '1', '2'):
        path = f'/tmp/lookup_{v}.pt'
        subprocess.check_call([sys.executable, __file__, path], env=dict(os.environ, SCFLOW_LOOKUP_SMEM=v))
        import torch
        outs.append(torch.load(path))
    same = torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    print('direct == staged == lean lookup bit for bit:', same, tuple(outs[0].shape), float(outs[0].abs().sum()))
    sys.exit(0 if same else 1)
