"""GPU: the fused kernels north_star names - one-kernel SepConvGRU pass, lookup + first motion-encoder convolution, pyramid
pooling in the build - each against a plain PyTorch fp32 restatement of the reference lines it replaces."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _gru_pass_ref(h, cxt, motion, wz, wr, wq, bz, br, bq):
    """raft_decoder.py:245-253 for one pass, NCHW fp64 on the CPU."""
    pad = (wz.shape[2] // 2, wz.shape[3] // 2)
    hx = torch.cat([h, cxt, motion], 1)
    z = torch.sigmoid(F.conv2d(hx, wz, bz, padding=pad))
    r = torch.sigmoid(F.conv2d(hx, wr, br, padding=pad))
    q = torch.tanh(F.conv2d(torch.cat([r * h, cxt, motion], 1), wq, bq, padding=pad))
    return (1 - z) * h + z * q


@pytest.mark.parametrize('kernel', [(1, 5), (5, 1)])
@pytest.mark.parametrize('b', [1, 3, 40])
def test_gru_pass_fused_matches_torch(kernel, b):
    import scflow_b200 as S
    g = torch.Generator().manual_seed(11 + b)
    hh = ww = 32
    h = torch.tanh(torch.randn(b, 128, hh, ww, generator=g))
    cxt = torch.relu(torch.randn(b, 128, hh, ww, generator=g))
    mot = torch.relu(torch.randn(b, 128, hh, ww, generator=g))
    ws = [torch.randn(128, 384, *kernel, generator=g) * 0.03 for _ in range(3)]
    bs = [torch.randn(128, generator=g) * 0.1 for _ in range(3)]
    ref = _gru_pass_ref(*(t.double() for t in (h, cxt, mot, *ws, *bs))).float()
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    op = S.ops.GruPassFused(*(t.cuda() for t in (*ws, *bs)))
    op.precompute(nhwc(cxt))
    out, out_hl = op(nhwc(h), nhwc(mot))
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).cpu()
    err = float((got - ref).abs().max())
    print(f'fused GRU pass {kernel} B={b}: max |err| {err:.2e} (|h| <= 1)')
    assert err < 5e-5            # split-bf16 products (2^-16 each) over K = 1920 with |w| ~ 0.03, |x| ~ 1
    # the split-bf16 copy (the next convolution's operand) carries the same values to ~2^-17
    assert float((S.ops.unsplit(out_hl).cpu() - got).abs().max()) < 2e-5


def test_gru_pass_fused_rejects_other_shapes():
    import scflow_b200 as S
    from scflow_b200 import ScfError
    g = torch.Generator().manual_seed(0)
    ws = [torch.randn(128, 384, 1, 5, generator=g).cuda() * 0.03 for _ in range(3)]
    bs = [torch.zeros(128).cuda() for _ in range(3)]
    op = S.ops.GruPassFused(*ws, *bs)
    x = torch.zeros(1, 16, 80, 128, device='cuda')            # 480x640 crops: 60x80 maps keep the two-kernel form
    op.precompute(x)
    with pytest.raises(ScfError, match='32 x 32 map'):
        op(x, x)


@pytest.mark.parametrize('size', [(256, 256), (480, 640)])
def test_reproject_emits_the_next_coarse_flow(size):
    """K3: the re-projection launch also writes flow8 = 1/8 * interpolate(flow, 1/8, bilinear, align_corners=True)
    (scflow_decoder.py:196-197) - against torch on the dense flow the same launch wrote."""
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    h, w = size
    scene = O.make_scene(31, 2, h, w)
    c = {k: v.cuda() for k, v in scene.items()}
    pts4 = S.ops.unproject(c['depth'], c['internel_k'], c['ref_rotation'], c['ref_translation'])
    g = torch.Generator().manual_seed(3)
    rot = (scene['ref_rotation'] + 0.01 * torch.randn(2, 3, 3, generator=g)).cuda()
    trs = (scene['ref_translation'] + torch.randn(2, 3, generator=g)).cuda()
    flow, flow8 = S.ops.reproject_down(pts4, c['internel_k'], rot, trs)
    assert torch.equal(flow, S.ops.reproject(pts4, c['internel_k'], rot, trs))
    want = F.interpolate(flow.cpu(), scale_factor=1 / 8, mode='bilinear', align_corners=True) / 8
    got = flow8.permute(0, 3, 1, 2).cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
    # and it is what the separate resize kernel produced (same taps, same blend)
    sep = S.ops.resize_bilinear_nchw(flow, h // 8, w // 8, scale=1 / 8)
    assert float((got - sep.cpu()).abs().max()) < 2e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('b,h8', [(2, 32), (5, 8), (1, 16)])
def test_corr_pyramid_one_kernel_matches_build_plus_pools(b, h8, monkeypatch):
    """K2: volume + pooled pyramid from one kernel vs (a) the fp32 oracle and (b) the GEMM + three floor-pool launches -
    same accumulation order and the reference's summation order in the pools, so (b) must agree bit for bit."""
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    f = O.make_features(21, b, h8, 32, channels=256)
    ref = O.correlation_pyramid(f['feat_render'], f['feat_real'], 4)
    monkeypatch.setenv('SCFLOW_CORR_FUSED', '1')
    fused = S.ops.corr_build(f['feat_render'].cuda(), f['feat_real'].cuda(), 4, precision=1)
    monkeypatch.setenv('SCFLOW_CORR_FUSED', '0')
    split = S.ops.corr_build(f['feat_render'].cuda(), f['feat_real'].cuda(), 4, precision=1)
    for l, (r, a, c) in enumerate(zip(ref, fused, split)):
        assert a.shape == r.shape == c.shape
        assert float((a.cpu() - r).abs().max()) < 2e-4, f'level {l} vs oracle'
        assert torch.equal(a, c), f'level {l}: fused and GEMM + pool forms differ (max {float((a - c).abs().max()):.2e})'


@pytest.mark.parametrize('b,mag', [(1, 3.0), (3, 12.0), (37, 40.0)])
def test_lookup_fused_with_first_motion_encoder_conv(b, mag):
    """N1: lookup + corr_net[0] (1x1, 324 -> 256, bias, ReLU) in one kernel vs (a) the oracle's lookup + fp64 matmul and (b) the
    stand-alone lookup kernel (whose neighbour indices are tested bit-exact) followed by the same matmul."""
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    f = O.make_features(41, b, 32, 32)
    pyr = O.correlation_pyramid(f['feat_render'], f['feat_real'], 4)
    g = torch.Generator().manual_seed(7)
    flow = mag * torch.randn(b, 2, 32, 32, generator=g)
    flow[0, :, :2, :] = torch.round(flow[0, :, :2, :])            # integral centres: the round-trip floor cases
    flow[-1, :, -1, :] = 0.
    w = torch.randn(256, 324, 1, 1, generator=g) * 0.05
    bias = torch.randn(256, generator=g) * 0.1
    corr = O.corr_lookup(pyr, flow, 4)                                                  # [b, 324, 32, 32]
    ref = torch.relu(torch.einsum('oc,bchw->bohw', w[:, :, 0, 0].double(), corr.double()) + bias.double().view(1, -1, 1, 1)).float()
    flow8 = flow.permute(0, 2, 3, 1).contiguous().cuda()
    levels = [p.cuda() for p in pyr]
    out = S.ops.lookup_conv(levels, flow8, w.cuda(), bias.cuda())
    got = S.ops.unsplit(out).cpu()
    err = float((got - ref).abs().max())
    print(f'lookup + conv1x1 fused, B={b}: max |err| {err:.2e} (|ref| max {float(ref.abs().max()):.2f})')
    assert err < 2e-4
    own = S.ops.corr_lookup_nhwc(levels, flow8, 4)[..., :324].permute(0, 3, 1, 2).cpu()
    ref2 = torch.relu(torch.einsum('oc,bchw->bohw', w[:, :, 0, 0].double(), own.double()) + bias.double().view(1, -1, 1, 1)).float()
    assert float((got - ref2).abs().max()) < 1e-4


@pytest.mark.parametrize('b', [32, 5, 1])
def test_linear_tc_split_k_chain_matches_fp64(b):
    """The pose head's FC tail on tcgen05 (scf_linear_tc; pose_head.py:203-210): fc0 writes raw split-K partial maps, fc1 sums them,
    adds fc0's bias and ReLU while forming its operand; the sums of fc1's maps + bias + ReLU equal the fp64 layers."""
    import scflow_b200 as S
    gen = torch.Generator().manual_seed(40 + b)
    x = torch.relu(torch.randn(b, 2048, generator=gen))
    w0, b0 = torch.randn(1024, 2048, generator=gen) / 2048 ** 0.5, 0.1 * torch.randn(1024, generator=gen)
    w1, b1 = torch.randn(256, 1024, generator=gen) / 1024 ** 0.5, 0.1 * torch.randn(256, generator=gen)
    y0 = torch.relu(x.double() @ w0.double().t() + b0.double())
    y1 = torch.relu(y0 @ w1.double().t() + b1.double())
    p0 = S.ops.linear_tc(x.cuda(), w0.cuda(), kr=256)
    assert p0.shape == (8, b, 1024)
    got0 = torch.relu(p0.sum(0).cpu().double() + b0.double())
    p1 = S.ops.linear_tc(p0, w1.cuda(), kr=128, x_bias=b0.cuda(), x_relu=True)
    assert p1.shape == (8, b, 256)
    got1 = torch.relu(p1.sum(0).cpu().double() + b1.double())
    e0, e1 = float((got0 - y0).abs().max()), float((got1 - y1).abs().max())
    print(f'linear_tc B={b}: fc0 max err {e0:.3e}, fc1 max err {e1:.3e} (|y1| max {float(y1.abs().max()):.2f})')
    assert e0 < 2e-5 and e1 < 2e-5
    # deterministic: a second launch reproduces the partial maps bit for bit
    assert torch.equal(p0, S.ops.linear_tc(x.cuda(), w0.cuda(), kr=256))
    # reduced form: the eight K-range blocks of a row tile reduce through distributed shared memory; one launch = one layer
    r0 = S.ops.linear_tc(x.cuda(), w0.cuda(), kr=256, reduce=True, bias=b0.cuda(), relu=True)
    r1 = S.ops.linear_tc(r0, w1.cuda(), kr=128, reduce=True, bias=b1.cuda(), relu=True)
    f0, f1 = float((r0.cpu().double() - y0).abs().max()), float((r1.cpu().double() - y1).abs().max())
    print(f'linear_tc reduced B={b}: fc0 max err {f0:.3e}, fc1 max err {f1:.3e}')
    assert r0.shape == (b, 1024) and r1.shape == (b, 256) and f0 < 2e-5 and f1 < 2e-5
    assert torch.equal(r0, S.ops.linear_tc(x.cuda(), w0.cuda(), kr=256, reduce=True, bias=b0.cuda(), relu=True))
