"""GPU: tcgen05 split-bf16 convolution (scf_conv2d_tc) against fp64 CPU convolutions."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def S():
    import scflow_b200
    return scflow_b200


def _ref(x, w, bias, k, act):
    pad = (k[0] // 2, k[1] // 2)
    y = F.conv2d(x.double(), w.double(), None if bias is None else bias.double(), padding=pad)
    return {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh, 'none': lambda t: t}[act](y).float()


CASES = [
    # cin, cout, kernel, (H, W), B, act
    (64, 128, (1, 1), (32, 32), 1, 'none'),        # single chunk, single tap
    (256, 128, (1, 1), (32, 32), 2, 'none'),       # K loop over 4 chunks (pipeline wrap with 3 stages)
    (128, 64, (3, 3), (32, 32), 2, 'relu'),        # taps + zero padding via TMA OOB
    (256, 192, (3, 3), (32, 32), 2, 'relu'),       # BN=192
    (256, 126, (3, 3), (32, 32), 2, 'relu'),       # cout not multiple of 16
    (384, 256, (1, 5), (32, 32), 2, 'sigmoid'),    # GRU shape, BN=256 (2 stages)
    (384, 128, (5, 1), (32, 32), 2, 'tanh'),
    (128, 512, (3, 3), (32, 32), 2, 'relu'),       # two N tiles
    (256, 2, (3, 3), (32, 32), 2, 'none'),         # BN=16
    (64, 32, (3, 3), (32, 32), 2, 'relu'),         # BN=32
    (328, 256, (1, 1), (32, 32), 2, 'relu'),       # ragged K (channels multiple of 8 only): TMA zero fill on C
    (128, 64, (3, 3), (60, 80), 1, 'relu'),        # 16x8 tiles, ragged rows (60 = 7.5 tiles)
    (128, 64, (3, 3), (20, 24), 3, 'relu'),        # odd sizes
]


@pytest.mark.parametrize('cin,cout,k,hw,b,act', CASES)
def test_conv_tc_matches_fp64(S, cin, cout, k, hw, b, act):
    gen = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(b, cin, *hw, generator=gen)
    w = torch.randn(cout, cin, *k, generator=gen) / math.sqrt(cin * k[0] * k[1])
    bias = 0.1 * torch.randn(cout, generator=gen)
    ref = _ref(x, w, bias, k, act)
    xs = S.ops.split_nchw(x.cuda())
    assert float(((S.ops.unsplit(xs).cpu() - x).abs() / x.abs().clamp_min(1e-3)).max()) < 2.0 ** -16    # hi+lo carries 16 bits
    pw = S.ops.pack_conv_weight_tc([w.cuda()])
    out_f32 = torch.full((b, hw[0], hw[1], cout), float('nan'), device='cuda')
    out_hl = torch.zeros(2, b, hw[0], hw[1], (cout + 7) // 8 * 8, device='cuda', dtype=torch.bfloat16)
    S.ops.conv2d_tc([(xs, 0, cin)], pw, bias.cuda(), cout, k, act=act, out_f32=out_f32, out_hl=out_hl)
    torch.cuda.synchronize()
    got = out_f32.permute(0, 3, 1, 2).cpu()
    err = float((got - ref).abs().max())
    err_hl = float((S.ops.unsplit(out_hl)[:, :cout].cpu() - ref).abs().max())
    print(f'conv_tc cin={cin} cout={cout} k={k} hw={hw}: max err f32 {err:.3e}, split {err_hl:.3e}, |ref|max {float(ref.abs().max()):.2f}')
    assert err < 5e-5, f'max err {err:.3e}'                       # bf16x3: ~2^-16 relative per product, fp32 accumulate
    assert err_hl < 1e-4


@pytest.mark.parametrize('cin,cout,hw,b,k', [(224, 128, (32, 32), 4, 3), (128, 128, (16, 16), 5, 3), (128, 128, (8, 8), 11, 3),
                                            (64, 96, (64, 64), 2, 3), (64, 96, (64, 64), 2, 1), (96, 128, (30, 40), 2, 3)])
def test_conv_tc_stride2(S, cin, cout, hw, b, k):
    """Stride-2 convolutions (pose head, encoder down-sampling): TMA element strides; small maps pack several samples per tile."""
    gen = torch.Generator().manual_seed(cin + cout + hw[0])
    x = torch.randn(b, cin, *hw, generator=gen)
    w = torch.randn(cout, cin, k, k, generator=gen) / math.sqrt(cin * k * k)
    ref = F.conv2d(x.double(), w.double(), None, stride=2, padding=k // 2).float()
    xs = S.ops.split_nchw(x.cuda())
    ho, wo = ref.shape[-2:]
    out = torch.full((b, ho, wo, cout), float('nan'), device='cuda')
    S.ops.conv2d_tc([(xs, 0, cin)], S.ops.pack_conv_weight_tc([w.cuda()]), None, cout, k, out_f32=out, stride=2)
    err = float((out.permute(0, 3, 1, 2).cpu() - ref).abs().max())
    assert err < 5e-5, f'max err {err:.3e}'


def test_conv_tc_segments_slices_and_gru_epilogues(S):
    from scflow_b200 import _lib
    gen = torch.Generator().manual_seed(3)
    b, hh, ww = 2, 32, 32
    h = torch.tanh(torch.randn(b, 128, hh, ww, generator=gen))
    cxt = torch.relu(torch.randn(b, 128, hh, ww, generator=gen))
    mot = torch.randn(b, 128, hh, ww, generator=gen)
    wz, wr, wq = (torch.randn(128, 384, 1, 5, generator=gen) / math.sqrt(1920) for _ in range(3))
    bz, br, bq = (0.1 * torch.randn(128, generator=gen) for _ in range(3))
    hx = torch.cat([h, cxt, mot], 1).double()
    z = torch.sigmoid(F.conv2d(hx, wz.double(), bz.double(), padding=(0, 2)))
    r = torch.sigmoid(F.conv2d(hx, wr.double(), br.double(), padding=(0, 2)))
    q = torch.tanh(F.conv2d(torch.cat([r * h.double(), cxt.double(), mot.double()], 1), wq.double(), bq.double(), padding=(0, 2)))
    hn = ((1 - z) * h.double() + z * q).float()
    # x lives in one wide split buffer [cxt | motion] to exercise channel offsets
    hs, h32 = S.ops.split_nchw(h.cuda(), want_f32=True)
    xs = torch.zeros(2, b, hh, ww, 256, device='cuda', dtype=torch.bfloat16)
    S.ops.split_nchw(cxt.cuda(), out=xs, coff=0)
    S.ops.split_nchw(mot.cuda(), out=xs, coff=128)
    zbuf = torch.empty(b, hh, ww, 128, device='cuda')
    rh = torch.zeros(2, b, hh, ww, 128, device='cuda', dtype=torch.bfloat16)
    pzr = S.ops.pack_conv_weight_tc([wz.cuda(), wr.cuda()])
    S.ops.conv2d_tc([(hs, 0, 128), (xs, 0, 128), (xs, 128, 128)], pzr, torch.cat([bz, br]).cuda(), 256, (1, 5), act='sigmoid',
                    out_f32=zbuf, epi=_lib.EPI_GRU_ZR, aux0=h32, out2_hl=rh)
    assert float((zbuf.permute(0, 3, 1, 2).cpu() - z.float()).abs().max()) < 2e-5
    assert float((S.ops.unsplit(rh).cpu() - (r * h.double()).float()).abs().max()) < 2e-5
    hn32 = torch.empty(b, hh, ww, 128, device='cuda')
    hns = torch.zeros(2, b, hh, ww, 128, device='cuda', dtype=torch.bfloat16)
    S.ops.conv2d_tc([(rh, 0, 128), (xs, 0, 128), (xs, 128, 128)], S.ops.pack_conv_weight_tc([wq.cuda()]), bq.cuda(), 128, (1, 5),
                    act='tanh', out_f32=hn32, out_hl=hns, epi=_lib.EPI_GRU_Q, aux0=h32, aux1=zbuf)
    assert float((hn32.permute(0, 3, 1, 2).cpu() - hn).abs().max()) < 3e-5
    assert float((S.ops.unsplit(hns).cpu() - hn).abs().max()) < 3e-5


T_CASES = [
    # cin, cout, kernel, (H, W), B, stride, residual, stats
    (64, 64, (3, 3), (64, 64), 1, 1, True, True),       # encoder residual block: raw fp32 + InstanceNorm partial sums
    (64, 96, (3, 3), (64, 64), 1, 2, False, True),      # down-sampling
    (128, 128, (3, 3), (30, 40), 1, 1, True, True),     # ragged rows and columns
    (96, 128, (1, 1), (32, 32), 3, 2, False, False),    # 16-wide output: stays on the pixels-as-rows kernel
    (16, 128, (7, 1), (32, 32), 2, 1, False, False),    # x-folded thin input (cin_pad 16, one ragged chunk per tap)
    (256, 126, (3, 3), (32, 32), 2, 1, False, False),   # cout % 8 != 0 (motion encoder output)
]


@pytest.mark.parametrize('transposed', ['0', '2'])
@pytest.mark.parametrize('cin,cout,k,hw,b,stride,residual,stats', T_CASES)
def test_conv_tc_small_n_variants(S, monkeypatch, transposed, cin, cout, k, hw, b, stride, residual, stats):
    """Layers of <= 128 output channels through both tilings (SCFLOW_TC_T: weights or pixels as the MMA's rows): same
    results, including the residual input, the InstanceNorm partial sums and untouched neighbouring channels."""
    monkeypatch.setenv('SCFLOW_TC_T', transposed)
    gen = torch.Generator().manual_seed(cin + 3 * cout + hw[1])
    x = torch.randn(b, cin, *hw, generator=gen)
    w = torch.randn(cout, cin, *k, generator=gen) / math.sqrt(cin * k[0] * k[1])
    bias = 0.1 * torch.randn(cout, generator=gen)
    pad = (k[0] // 2, k[1] // 2)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=stride, padding=pad)
    ho, wo = ref.shape[-2:]
    res = torch.randn(b, ho, wo, cout, generator=gen) if residual else None
    if residual:
        ref = ref + res.permute(0, 3, 1, 2).double()
    ref = torch.relu(ref).float()
    xs = S.ops.split_nchw(x.cuda())
    pw = S.ops.pack_conv_weight_tc([w.cuda()])
    cpad = (cout + 7) // 8 * 8
    out_f32 = torch.full((b, ho, wo, cpad + 8), 7.0, device='cuda')
    out_hl = torch.full((2, b, ho, wo, cpad + 16), 7.0, device='cuda', dtype=torch.bfloat16)
    n_tiles, _ = S.ops.conv2d_tc_tiles(b, ho, wo)
    st = torch.zeros(n_tiles * 4 * 2 * cout, device='cuda') if stats else None
    S.ops.conv2d_tc([(xs, 0, cin)], pw, bias.cuda(), cout, k, act='relu', out_f32=out_f32, out_hl=out_hl, out_hl_coff=8,
                    stride=stride, aux0=None if res is None else res.cuda(), stats=st, out_pad_writable=cout % 8 != 0)
    torch.cuda.synchronize()
    got = out_f32[..., :cout].permute(0, 3, 1, 2).cpu()
    got_hl = S.ops.unsplit(out_hl)[:, 8:8 + cout].cpu()
    assert float((got - ref).abs().max()) < 5e-5
    assert float((got_hl - ref).abs().max()) < 1e-4
    # channels outside [coff, coff + round_up(cout, 8)) keep their contents
    assert bool((out_f32[..., cpad:] == 7.0).all()) and bool((out_hl[..., :8] == 7.0).all()) and bool((out_hl[..., 8 + cpad:] == 7.0).all())
    if stats:
        rows = st.view(-1, 2, cout).double().sum(0).cpu()
        assert float((rows[0] - ref.double().sum((0, 2, 3))).abs().max()) < 1e-2
        assert float((rows[1] - ref.double().pow(2).sum((0, 2, 3))).abs().max()) < 2e-2


ROWS_CASES = [
    # (H, W), B, residual, stats, act, out_hl channel offset
    ((128, 128), 2, False, True, 'none', 16),     # encoder stage 1 (InstanceNorm): raw fp32 + partial sums, full-width strips
    ((128, 128), 2, True, False, 'relu', 16),     # encoder stage 1 (folded BatchNorm): residual + ReLU, fp32 + split outputs
    ((96, 160), 1, True, True, 'relu', 8),        # two column strips, the second ragged (32 of 128 pixels); 16 B store path
    ((80, 96), 2, False, False, 'relu', 0),       # one ragged strip; CTA ranges cut inside images
    ((7, 128), 24, False, True, 'none', 0),       # images shorter than the TMEM ring: many segment starts / ends per CTA
]


@pytest.mark.parametrize('hw,b,residual,stats,act,hl_coff', ROWS_CASES)
def test_conv_rows_kernel_matches_fp64_and_generic_tile(S, monkeypatch, hw, b, residual, stats, act, hl_coff):
    """scf_conv_rows.cu (3x3, 64 -> 64, rolling-rows accumulation in TMEM) against an fp64 convolution and against the generic
    tile (SCFLOW_TC_ROWS=0): outputs, InstanceNorm partial sums, untouched neighbouring channels."""
    cin = cout = 64
    gen = torch.Generator().manual_seed(hw[0] * 1000 + hw[1] + b)
    x = torch.randn(b, cin, *hw, generator=gen)
    w = torch.randn(cout, cin, 3, 3, generator=gen) / math.sqrt(cin * 9)
    bias = 0.1 * torch.randn(cout, generator=gen)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1)
    res = torch.randn(b, *hw, cout, generator=gen) if residual else None
    if residual:
        ref = ref + res.permute(0, 3, 1, 2).double()
    if act == 'relu':
        ref = torch.relu(ref)
    ref = ref.float()
    xs = S.ops.split_nchw(x.cuda())
    pw = S.ops.pack_conv_weight_tc([w.cuda()])
    n_tiles, _ = S.ops.conv2d_tc_tiles(b, *hw)
    outs = {}
    for mode in ('1', 'ring8', '0'):       # rolling rows with the alias-slot TMEM scheme (default) / plain ring of eight / generic tile
        monkeypatch.setenv('SCFLOW_TC_ROWS', '0' if mode == '0' else '1')
        monkeypatch.setenv('SCFLOW_ROWS_ALIAS', '0' if mode == 'ring8' else '1')
        out_f32 = torch.full((b, *hw, cout + 8), 7.0, device='cuda')
        out_hl = torch.full((2, b, *hw, cout + 32), 7.0, device='cuda', dtype=torch.bfloat16)
        st = torch.zeros(n_tiles * 4 * 2 * cout, device='cuda') if stats else None
        S.ops.conv2d_tc([(xs, 0, cin)], pw, bias.cuda(), cout, 3, act=act, out_f32=out_f32, out_hl=out_hl, out_hl_coff=hl_coff,
                        aux0=None if res is None else res.cuda(), stats=st)
        torch.cuda.synchronize()
        got = out_f32[..., :cout].permute(0, 3, 1, 2).cpu()
        got_hl = S.ops.unsplit(out_hl)[:, hl_coff:hl_coff + cout].cpu()
        err, err_hl = float((got - ref).abs().max()), float((got_hl - ref).abs().max())
        print(f'conv rows={mode} hw={hw} b={b}: max err f32 {err:.3e}, split {err_hl:.3e}')
        assert err < 5e-5 and err_hl < 1e-4
        assert bool((out_f32[..., cout:] == 7.0).all()) and bool((out_hl[..., :hl_coff] == 7.0).all()) and \
            bool((out_hl[..., hl_coff + cout:] == 7.0).all())
        if stats:
            # the partial sums describe the kernel's OWN fp32 output (what InstanceNorm normalises) ...
            rows = st.view(-1, 2, cout).double().sum(0).cpu()
            own = out_f32[..., :cout].double().cpu()
            assert float((rows[0] - own.sum((0, 1, 2))).abs().max()) < 1e-3
            assert float((rows[1] - own.pow(2).sum((0, 1, 2))).abs().max()) < 1e-3
            # ... and agree with the fp64 reference to the split-bf16 products' relative accuracy
            sq = ref.double().pow(2).sum((0, 2, 3))
            assert float((rows[0] - ref.double().sum((0, 2, 3))).abs().max()) < 2e-2
            assert float(((rows[1] - sq).abs() / sq).max()) < 2e-5
        outs[mode] = got
    # same products, different summation order inside the fp32 accumulator
    assert float((outs['1'] - outs['0']).abs().max()) < 2e-5 and float((outs['ring8'] - outs['0']).abs().max()) < 2e-5


@pytest.mark.parametrize('hw,b,stats,act', [((256, 128), 2, True, 'none'), ((96, 160), 2, False, 'relu'), ((30, 128), 11, True, 'none')])
def test_conv_stem_rows_kernel_matches_fp64_and_generic_tile(S, monkeypatch, hw, b, stats, act):
    """The encoders' stem after the x-fold (7x1 kernel, vertical stride 2, 32 -> 64 channels) on the rolling-rows kernel
    (conv_stem_rows_kernel) against an fp64 convolution and the generic tile (SCFLOW_TC_ROWS=0)."""
    cin, cout = 32, 64
    gen = torch.Generator().manual_seed(hw[0] + 7 * hw[1] + b)
    x = torch.randn(b, cin, *hw, generator=gen)
    w = torch.randn(cout, cin, 7, 1, generator=gen) / math.sqrt(cin * 7)
    bias = 0.1 * torch.randn(cout, generator=gen)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=(2, 1), padding=(3, 0))
    if act == 'relu':
        ref = torch.relu(ref)
    ref = ref.float()
    ho, wo = ref.shape[-2:]
    xs = S.ops.split_nchw(x.cuda())
    pw = S.ops.pack_conv_weight_tc([w.cuda()])
    n_tiles, _ = S.ops.conv2d_tc_tiles(b, ho, wo)
    outs = {}
    for mode in ('1', 'ring8', '0'):       # alias-slot TMEM scheme (default) / plain ring of eight / generic tile
        monkeypatch.setenv('SCFLOW_TC_ROWS', '0' if mode == '0' else '1')
        monkeypatch.setenv('SCFLOW_ROWS_ALIAS', '0' if mode == 'ring8' else '1')
        out_f32 = torch.full((b, ho, wo, cout), 7.0, device='cuda')
        out_hl = torch.full((2, b, ho, wo, cout), 7.0, device='cuda', dtype=torch.bfloat16)
        st = torch.zeros(n_tiles * 4 * 2 * cout, device='cuda') if stats else None
        S.ops.conv2d_tc([(xs, 0, cin)], pw, bias.cuda(), cout, (7, 1), act=act, out_f32=out_f32, out_hl=None if stats else out_hl,
                        stride_xy=(1, 2), stats=st)
        torch.cuda.synchronize()
        got = out_f32.permute(0, 3, 1, 2).cpu()
        err = float((got - ref).abs().max())
        print(f'stem rows={mode} hw={hw} b={b}: max err f32 {err:.3e}')
        assert err < 5e-5
        if not stats:
            assert float((S.ops.unsplit(out_hl).cpu() - ref).abs().max()) < 1e-4
        else:
            rows = st.view(-1, 2, cout).double().sum(0).cpu()
            own = out_f32.double().cpu()
            assert float((rows[0] - own.sum((0, 1, 2))).abs().max()) < 1e-3
            assert float((rows[1] - own.pow(2).sum((0, 1, 2))).abs().max()) < 1e-3
        outs[mode] = got
    assert float((outs['1'] - outs['0']).abs().max()) < 2e-5 and float((outs['ring8'] - outs['0']).abs().max()) < 2e-5


def test_conv_rows_kernel_split_residual(S):
    """The residual of a 64-channel BasicBlock taken from the block input's split-bf16 planes (scf_tc_conv_desc.aux0_hl) equals the
    fp32 residual to the planes' 2^-17 relative accuracy; layers outside the rolling-rows kernel refuse the option."""
    gen = torch.Generator().manual_seed(77)
    b, hw = 2, (128, 128)
    x = torch.randn(b, 64, *hw, generator=gen)
    w = torch.randn(64, 64, 3, 3, generator=gen) / 24.0
    bias = 0.1 * torch.randn(64, generator=gen)
    res = torch.randn(b, 64, *hw, generator=gen)
    ref = torch.relu(F.conv2d(x.double(), w.double(), bias.double(), padding=1) + res.double()).float()
    xs, rs = S.ops.split_nchw(x.cuda()), S.ops.split_nchw(res.cuda())
    pw = S.ops.pack_conv_weight_tc([w.cuda()])
    out_hl = torch.zeros(2, b, *hw, 64, device='cuda', dtype=torch.bfloat16)
    S.ops.conv2d_tc([(xs, 0, 64)], pw, bias.cuda(), 64, 3, act='relu', out_hl=out_hl, aux0_hl=rs)
    torch.cuda.synchronize()
    err = float((S.ops.unsplit(out_hl).cpu() - ref).abs().max())
    print(f'rows kernel, split residual: max err {err:.3e}')
    assert err < 1e-4
    xs2 = S.ops.split_nchw(x[:, :, :32, :32].contiguous().cuda())           # a 32 x 32 map stays on the generic tile
    with pytest.raises(Exception, match='aux0_hl'):
        S.ops.conv2d_tc([(xs2, 0, 64)], pw, bias.cuda(), 64, 3, act='relu', out_hl=torch.zeros(2, b, 32, 32, 64, device='cuda', dtype=torch.bfloat16),
                        aux0_hl=S.ops.split_nchw(res[:, :, :32, :32].contiguous().cuda()))
