import math, os, sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
import scflow_b200 as S
b, hw, cin, cout = 2, (128, 128), 64, 64
gen = torch.Generator().manual_seed(1)
x = torch.randn(b, cin, *hw, generator=gen)
w = torch.randn(cout, cin, 3, 3, generator=gen) / math.sqrt(cin * 9)
bias = 0.1 * torch.randn(cout, generator=gen)
ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
xs = S.ops.split_nchw(x.cuda()); pw = S.ops.pack_conv_weight_tc([w.cuda()])
n_tiles, _ = S.ops.conv2d_tc_tiles(b, *hw)
for mode in ('1', '0'):
    os.environ['SCFLOW_TC_ROWS'] = mode
    out = torch.zeros(b, *hw, cout, device='cuda')
    st = torch.zeros(n_tiles * 4 * 2 * cout, device='cuda')
    S.ops.conv2d_tc([(xs, 0, cin)], pw, bias.cuda(), cout, 3, out_f32=out, stats=st)
    torch.cuda.synchronize()
    rows = st.view(-1, 2, cout).double().sum(0).cpu()
    o = out.double().cpu()
    print(mode, 'stats-vs-own-output sum', float((rows[0] - o.sum((0, 1, 2))).abs().max()), 'sq', float((rows[1] - o.pow(2).sum((0, 1, 2))).abs().max()),
          '| own-output-vs-ref sq', float((o.pow(2).sum((0, 1, 2)) - ref.pow(2).sum((0, 2, 3))).abs().max()),
          'signed mean', float((o.pow(2).sum((0, 1, 2)) - ref.pow(2).sum((0, 2, 3))).mean()))
