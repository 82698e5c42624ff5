"""GPU: every exported kernel family against the CPU oracle on seeded inputs (through the C ABI via scflow_b200.ops)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import scflow_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def S():
    import scflow_b200
    return scflow_b200


def _rand(gen, *shape, scale=1.0):
    return torch.randn(*shape, generator=gen) * scale


@pytest.mark.parametrize('cin,cout,k,stride,hw,act', [
    (324, 256, (1, 1), 1, (32, 32), 'relu'),       # corr_net.0
    (256, 192, (3, 3), 1, (32, 32), 'relu'),       # corr_net.1
    (2, 128, (7, 7), 1, (32, 32), 'relu'),         # flow_net.0 (scalar loader)
    (256, 126, (3, 3), 1, (32, 32), 'relu'),       # out_net (cout not multiple of 4)
    (384, 128, (1, 5), 1, (32, 32), 'tanh'),       # gru 1x5
    (384, 128, (5, 1), 1, (20, 28), 'sigmoid'),    # gru 5x1, ragged tile
    (256, 2, (3, 3), 1, (32, 32), 'none'),         # flow predict
    (256, 1, (1, 1), 1, (32, 32), 'sigmoid'),      # mask predict
    (1, 64, (3, 3), 1, (32, 32), 'relu'),          # mask_encoder.0
    (224, 128, (3, 3), 2, (32, 32), 'none'),       # pose head stride 2
    (128, 128, (3, 3), 2, (15, 9), 'none'),        # odd sizes, stride 2
])
def test_conv2d_against_torch_cpu(S, cin, cout, k, stride, hw, act):
    gen = torch.Generator().manual_seed(cin * 131 + cout)
    b = 2
    x = _rand(gen, b, cin, *hw)
    w = _rand(gen, cout, cin, *k, scale=1.0 / math.sqrt(cin * k[0] * k[1]))
    bias = _rand(gen, cout, scale=0.1)
    pad = (k[0] // 2, k[1] // 2)
    ref = F.conv2d(x, w, bias, stride=stride, padding=pad)
    ref = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh, 'none': lambda t: t}[act](ref)
    xh = S.ops.nchw_to_nhwc(x.cuda())
    pw = S.ops.pack_conv_weight([w.cuda()])
    out = S.ops.conv2d_nhwc([(xh, 0, cin)], pw, bias.cuda(), cout, k, stride, pad, act=act)
    got = S.ops.nhwc_to_nchw(out).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-5          # fp32 accumulate, different summation order only


def test_conv2d_segments_and_slices(S):
    """Three-segment input (torch.cat removal) written into a channel slice of a wider buffer."""
    gen = torch.Generator().manual_seed(1)
    a, b_, c = _rand(gen, 2, 128, 16, 16), _rand(gen, 2, 64, 16, 16), _rand(gen, 2, 32, 16, 16)
    w = _rand(gen, 40, 224, 3, 3, scale=0.02)
    ref = F.conv2d(torch.cat([a, b_, c], 1), w, None, padding=1)
    wide = torch.zeros(2, 16, 16, 192, device='cuda')          # b lives at channel offset 64 of a wider buffer
    S.ops.nchw_to_nhwc(b_.cuda(), out=wide, coff=64)
    out = torch.full((2, 16, 16, 64), -7.0, device='cuda')
    S.ops.conv2d_nhwc([(S.ops.nchw_to_nhwc(a.cuda()), 0, 128), (wide, 64, 64), (S.ops.nchw_to_nhwc(c.cuda()), 0, 32)],
                      S.ops.pack_conv_weight([w.cuda()]), None, 40, 3, out=out, out_coff=8)
    got = out.cpu()
    assert float((got[..., 8:48].permute(0, 3, 1, 2) - ref).abs().max()) < 2e-5
    assert torch.all(got[..., :8] == -7.0) and torch.all(got[..., 48:] == -7.0)      # neighbours untouched


@pytest.mark.parametrize('precision', [0, 1])
@pytest.mark.parametrize('shape,C,levels', [((32, 32), 256, 4), ((8, 12), 64, 3), ((15, 20), 32, 3), ((60, 80), 256, 4)])
def test_corr_build(S, shape, C, levels, precision):
    if precision == 1 and (shape[0] * shape[1]) % 16:
        pytest.skip('tensor-core build needs H8*W8 % 16 == 0')
    f = O.make_features(3, 2 if shape[0] < 60 else 1, shape[0], shape[1], channels=C)
    ref = O.correlation_pyramid(f['feat_render'], f['feat_real'], levels)
    got = S.ops.corr_build(f['feat_render'].cuda(), f['feat_real'].cuda(), levels, precision=precision)
    assert len(got) == levels
    for l, (r, g) in enumerate(zip(ref, got)):
        assert g.shape == r.shape
        assert float((g.cpu() - r).abs().max()) < (5e-5 if precision == 0 else 2e-4), f'level {l}'
    mod = S.CorrelationPyramid(num_levels=levels, precision=precision)
    again = mod(f['feat_render'].cuda(), f['feat_real'].cuda())
    assert all(torch.equal(a, b) for a, b in zip(got, again))      # deterministic


@pytest.mark.parametrize('shape,levels,radius,mag', [((32, 32), 4, 4, 6.0), ((16, 24), 3, 3, 30.0), ((15, 20), 3, 4, 3.0)])
def test_corr_lookup_values_and_bit_exact_indices(S, shape, levels, radius, mag):
    h, w = shape
    f = O.make_features(11, 2, h, w, channels=32)
    pyr = O.correlation_pyramid(f['feat_render'], f['feat_real'], levels)
    gen = torch.Generator().manual_seed(17)
    flow = mag * torch.randn(2, 2, h, w, generator=gen)
    flow[0, :, :2, :] = torch.round(flow[0, :, :2, :])          # integral centres: the round-trip floor cases
    flow[1, :, -1, :] = 0.
    ref = O.corr_lookup(pyr, flow, radius)
    got = S.CorrLookup(radius=radius, align_corners=True)([p.cuda() for p in pyr], flow.cuda()).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-5
    # index work: bit exact against the oracle's replay of the reference's fp32 sequence
    flow8 = flow.permute(0, 2, 3, 1).contiguous().cuda()
    hl, wl = h, w
    for lvl in range(levels):
        x0, y0, _, _ = O.lookup_taps(flow, lvl, wl, hl, radius)
        gx0, gy0 = S.ops.corr_lookup_taps(flow8, lvl, radius)
        assert torch.equal(gx0.cpu(), x0[:, :, :, :, 0]), f'x0 level {lvl}'
        assert torch.equal(gy0.cpu(), y0[:, :, :, 0, :]), f'y0 level {lvl}'
        hl, wl = hl // 2, wl // 2


def test_corr_lookup_ramp_kat(S):
    h = w = 8
    vol = (100. * torch.arange(h).view(h, 1) + torch.arange(w).view(1, w)).float().view(1, 1, h, w).repeat(h * w, 1, 1, 1)
    out = S.CorrLookup(radius=1)([vol.cuda()], torch.zeros(1, 2, h, w, device='cuda')).cpu()
    assert out[0, :, 4, 3].tolist()[:4] == [302., 402., 502., 303.]
    # zeros padding off the border; the normalise/un-normalise round trip leaves ~6e-8 weights on some OOB-adjacent taps
    # (SURVEY Appendix A.4), so compare with the oracle rather than exact zeros
    ref = O.corr_lookup([vol], torch.zeros(1, 2, h, w), radius=1)
    assert float((out - ref).abs().max()) < 1e-5
    assert [round(v) for v in out[0, :, 0, 0].tolist()] == [0, 0, 0, 0, 0, 100, 0, 1, 101]


def test_motion_encoder_gru_heads_modules(S):
    from tests.util import build_decoder_from_oracle_weights
    dec, sd = build_decoder_from_oracle_weights(4, 1)
    gen = torch.Generator().manual_seed(23)
    corr, flow = _rand(gen, 2, 324, 32, 32), _rand(gen, 2, 2, 32, 32, scale=3.0)
    h, x = torch.tanh(_rand(gen, 2, 128, 32, 32)), _rand(gen, 2, 256, 32, 32)
    with torch.no_grad():
        assert float((dec.encoder(corr.cuda(), flow.cuda()).cpu() - O.motion_encoder(sd, corr, flow)).abs().max()) < 5e-5
        assert float((dec.gru(h.cuda(), x.cuda()).cpu() - O.sepconv_gru(sd, h, x)).abs().max()) < 2e-5
        assert float((dec.flow_pred(h.cuda()).cpu() - O.xhead(sd, 'flow_pred.', h, 'flow')).abs().max()) < 2e-5
        assert float((dec.mask_pred(h.cuda()).cpu() - O.xhead(sd, 'mask_pred.', h, 'mask')).abs().max()) < 2e-5


def test_pose_head_module_and_label0_quirk(S):
    from tests.util import build_decoder_from_oracle_weights, load_golden, assert_matches_digest
    dec, sd = build_decoder_from_oracle_weights(3, 1)
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(3, 224, 32, 32, generator=gen)
    with torch.no_grad():
        r1, t1 = dec.pose_pred(x.cuda(), torch.tensor([5, 7, 9]).cuda())
        r2, t2 = dec.pose_pred(x.cuda(), torch.tensor([5, 5, 5]).cuda())
        r3, t3 = dec.pose_pred(x.cuda(), torch.tensor([6, 5, 5]).cuda())
    assert torch.equal(r1, r2) and torch.equal(t1, t2)
    assert not torch.equal(r1, r3)
    g = load_golden('pose_head')
    assert_matches_digest(g, 'rot', r1, atol=2e-5)
    assert_matches_digest(g, 'trans', t1, atol=2e-5)
    mr, mt = O.pose_head(sd, x, torch.tensor([5, 7, 9]))
    assert float((r1.cpu() - mr).abs().max()) < 2e-5 and float((t1.cpu() - mt).abs().max()) < 2e-5


def test_geometry_kernels(S):
    from tests.util import load_golden, assert_matches_digest
    g = load_golden('geometry')
    scene = O.make_scene(21, 3, 64, 64)
    gen = torch.Generator().manual_seed(2)
    d_rot = torch.tensor([1., 0., 0., 0., 1., 0.]) + 0.1 * torch.randn(3, 6, generator=gen)
    d_trs = 0.1 * torch.randn(3, 3, generator=gen)
    c = {k: v.cuda() for k, v in scene.items()}
    r, t = S.get_pose_from_delta_pose(d_rot.cuda(), d_trs.cuda(), c['ref_rotation'], c['ref_translation'],
                                      depth_transform='exp', detach_depth_for_xy=True)
    assert_matches_digest(g, 'rot', r, atol=1e-6)
    assert_matches_digest(g, 'trans', t, atol=2e-4)
    pts4 = S.unproject_dense(c['depth'], c['internel_k'], c['ref_rotation'], c['ref_translation'])
    ref_pts = O.unproject_dense(scene['depth'], scene['internel_k'], scene['ref_rotation'], scene['ref_translation'])
    fg = scene['depth'] > 0
    err = (pts4[..., :3].cpu().permute(0, 3, 1, 2) - ref_pts).abs().amax(1)
    assert float(err[fg].max()) < 2e-2                           # mm; fp32 inverse differences (Appendix A.10)
    assert torch.equal(pts4[..., 3].cpu() > 0, fg)
    flow = S.get_flow_from_delta_pose_dense(r, t, c['internel_k'], pts4, invalid_num=0.)
    assert_matches_digest(g, 'flow', flow, atol=3e-3)
    ident = S.get_flow_from_delta_pose_dense(c['ref_rotation'], c['ref_translation'], c['internel_k'], pts4, invalid_num=7.).cpu()
    assert float(ident[:, 0][fg].abs().max()) < 3e-3 and torch.all(ident[:, 1][~fg] == 7.)
    # reference-compatible list API
    p2, p3 = S.cal_3d_2d_corr(c['depth'][0], c['internel_k'][0], c['ref_rotation'][0], c['ref_translation'][0])
    assert p2.shape[0] == int(fg[0].sum()) and p3.shape == (p2.shape[0], 3)
    f_list = S.get_flow_from_delta_pose_and_points(r[:1], t[:1], c['internel_k'][:1], [p2], [p3], 64, 64, invalid_num=0.)
    assert torch.equal(f_list, flow[:1])


def test_resize_bilinear_align_corners(S):
    gen = torch.Generator().manual_seed(5)
    x = _rand(gen, 2, 2, 64, 48)
    down = S.ops.resize_bilinear_nchw(x.cuda(), 8, 6, scale=0.125).cpu()
    ref = 0.125 * F.interpolate(x, size=(8, 6), mode='bilinear', align_corners=True)
    assert float((down - ref).abs().max()) < 1e-6
    y = _rand(gen, 2, 2, 8, 6)
    up = S.ops.resize_bilinear_nchw(y.cuda(), 64, 48, scale=8.0).cpu()
    ref_up = 8.0 * F.interpolate(y, size=(64, 48), mode='bilinear', align_corners=True)
    assert float((up - ref_up).abs().max()) < 1e-5


def test_argument_errors_are_reported(S):
    from scflow_b200 import ScfError
    with pytest.raises(ScfError, match='bad shape'):
        S.ops.corr_build(torch.zeros(1, 8, 4, 4, device='cuda'), torch.zeros(1, 8, 4, 4, device='cuda'), num_levels=9)
    with pytest.raises(ValueError):
        S.ops.corr_lookup_nhwc([torch.zeros(5, device='cuda')], torch.zeros(1, 4, 4, 2, device='cuda'), 4)
