"""CPU: the oracle restatement reproduces the committed fixtures generated from the reference itself
(oracle/make_golden.py).  This is the pin that travels: /root/reference is not needed here."""
import numpy as np
import pytest
import torch

from oracle import scflow_oracle as O
from tests.util import assert_matches_digest, load_golden

NAMES = ['flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation']
TOLS = dict(flow_from_pose=2e-3, flow_from_pred=2e-4, rotation=1e-5, translation=2e-3, mask=1e-5, delta_rotation=1e-5,
            delta_translation=1e-5)


def run_decoder_case(name):
    g = load_golden(name)
    seed, b, h, w, iters = (int(g['meta/' + k]) for k in ('seed', 'batch', 'h', 'w', 'iters'))
    ident = bool(int(g['meta/identity_head']))
    scene = O.make_scene(seed, b, h, w)
    f = O.make_features(seed, b, h // 8, w // 8)
    sd = O.make_decoder_weights(seed)
    trace = {}
    with torch.no_grad():
        outs = O.decoder_forward(sd, f['feat_render'], f['feat_real'], f['h_feat'], f['cxt_feat'], scene['ref_rotation'],
                                 scene['ref_translation'], scene['depth'], scene['internel_k'], scene['label'],
                                 torch.zeros(b, 2, h, w), 0., iters=iters, identity_pose_head=ident, trace=trace)
        pyr = O.correlation_pyramid(f['feat_render'], f['feat_real'])
    for nm, lst in zip(NAMES, outs):
        for i, t in enumerate(lst):
            assert_matches_digest(g, f'{nm}/{i}', t, atol=TOLS[nm])
    for l, t in enumerate(pyr):
        assert_matches_digest(g, f'pyramid/{l}', t, atol=1e-5)
    for key in ('flow8', 'corr', 'motion', 'h', 'd_flow', 'mask8'):
        for i, t in enumerate(trace[key]):
            assert_matches_digest(g, f'trace/{key}/{i}', t, atol=1e-4)


def test_decoder_256_b2_it4():
    run_decoder_case('decoder_256_b2_it4')


def test_decoder_480x640_identity_head():
    run_decoder_case('decoder_480x640_b1_it2')


def test_get_pose_config1():
    g = load_golden('get_pose_256_b1_it4')
    seed, b, iters = int(g['meta/seed']), int(g['meta/batch']), int(g['meta/iters'])
    scene = O.make_scene(seed, b)
    sd = O.make_model_weights(seed)
    with torch.no_grad():
        outs = O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                          scene['depth'], scene['internel_k'], scene['label'], iters=iters)
        enc = {k[len('render_encoder.'):]: v for k, v in sd.items() if k.startswith('render_encoder.')}
        assert_matches_digest(g, 'feat_real', O.raft_encoder(enc, scene['real_images'], 'IN'), atol=1e-4)
    for nm, lst in zip(NAMES, outs):
        for i, t in enumerate(lst):
            assert_matches_digest(g, f'{nm}/{i}', t, atol=max(TOLS[nm], 5e-4) if 'flow' in nm or nm == 'translation' else TOLS[nm])


def test_lookup_ramp_kat():
    """x-major window order: ramp volume 100*y + x, query (4,3), r=1 -> [302, 402, 502, 303, ...]."""
    g = load_golden('lookup_ramp')
    h = w = 8
    vol = (100. * torch.arange(h).view(h, 1) + torch.arange(w).view(1, w)).float().view(1, 1, h, w).repeat(h * w, 1, 1, 1)
    out = O.corr_lookup([vol], torch.zeros(1, 2, h, w), radius=1)
    assert out[0, :, 4, 3].tolist()[:4] == [302., 402., 502., 303.]
    assert_matches_digest(g, 'ramp/r1', out, atol=0.0)
    gen = torch.Generator().manual_seed(5)
    flow2 = 3.0 * torch.randn(1, 2, h, w, generator=gen)
    pyr = [vol, torch.nn.functional.avg_pool2d(vol, 2, 2)]
    assert_matches_digest(g, 'ramp/frac_r2', O.corr_lookup(pyr, flow2, radius=2), atol=1e-5)
    assert_matches_digest(g, 'ramp/frac_r2', O.corr_lookup_explicit(pyr, flow2, radius=2), atol=1e-4)


def test_lookup_explicit_matches_grid_sample_and_taps_are_integral():
    f = O.make_features(4, 1, 16, 16, channels=32)
    pyr = O.correlation_pyramid(f['feat_render'], f['feat_real'], 3)
    gen = torch.Generator().manual_seed(9)
    flow = 4.0 * torch.randn(1, 2, 16, 16, generator=gen)
    flow[0, :, 0, 0] = 0.           # exactly integral centre: exercises the normalise/un-normalise round trip
    a = O.corr_lookup(pyr, flow, 4)
    b = O.corr_lookup_explicit(pyr, flow, 4)
    assert float((a - b).abs().max()) < 2e-5
    x0, y0, fx, fy = O.lookup_taps(flow, 1, 8, 8, 4)
    assert x0.dtype == torch.int32 and float(fx.min()) >= 0. and float(fx.max()) < 1.


def test_pose_head_label0_quirk_and_golden():
    g = load_golden('pose_head')
    sd = O.make_decoder_weights(3)
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(3, 224, 32, 32, generator=gen)
    with torch.no_grad():
        r1, t1 = O.pose_head(sd, x, torch.tensor([5, 7, 9]))
        r2, t2 = O.pose_head(sd, x, torch.tensor([5, 5, 5]))
    assert torch.equal(r1, r2) and torch.equal(t1, t2)
    assert_matches_digest(g, 'rot', r1, atol=1e-5)
    assert_matches_digest(g, 'trans', t1, atol=1e-5)


def test_geometry_golden_and_identity():
    g = load_golden('geometry')
    scene = O.make_scene(21, 3, 64, 64)
    gen = torch.Generator().manual_seed(2)
    d_rot = torch.tensor([1., 0., 0., 0., 1., 0.]) + 0.1 * torch.randn(3, 6, generator=gen)
    d_trs = 0.1 * torch.randn(3, 3, generator=gen)
    r, t = O.update_pose(d_rot, d_trs, scene['ref_rotation'], scene['ref_translation'])
    assert_matches_digest(g, 'rot', r, atol=1e-6)
    assert_matches_digest(g, 'trans', t, atol=1e-4)
    pts = O.unproject_dense(scene['depth'], scene['internel_k'], scene['ref_rotation'], scene['ref_translation'])
    flow = O.reproject_dense(pts, scene['depth'], scene['internel_k'], r, t, 0.)
    assert_matches_digest(g, 'flow', flow, atol=2e-3)
    ident = O.reproject_dense(pts, scene['depth'], scene['internel_k'], scene['ref_rotation'], scene['ref_translation'], 7.)
    fg = scene['depth'] > 0
    assert float(ident[:, 0][fg].abs().max()) < 2e-3           # KAT (iv): identity delta pose -> zero flow on fg
    assert torch.all(ident[:, 0][~fg] == 7.)                   # and invalid_num on bg
    rot = O.ortho6d_to_matrix(d_rot)
    eye = torch.bmm(rot, rot.transpose(1, 2))
    assert float((eye - torch.eye(3)).abs().max()) < 1e-5 and float((torch.det(rot) - 1).abs().max()) < 1e-5


def test_pyramid_matches_einsum():
    """KAT (ii): pyr[0] == einsum('nchw,ncij->nhwij') / sqrt(C)."""
    f = O.make_features(1, 2, 8, 8, channels=64)
    pyr = O.correlation_pyramid(f['feat_render'], f['feat_real'], 3)
    ref = torch.einsum('nchw,ncij->nhwij', f['feat_render'], f['feat_real']) / 8.0
    assert float((pyr[0].view(2, 8, 8, 8, 8) - ref).abs().max()) < 1e-5
    assert pyr[1].shape == (128, 1, 4, 4) and pyr[2].shape == (128, 1, 2, 2)
    # floor pooling on odd sizes (60x80 -> 30x40 -> 15x20 -> 7x10)
    odd = O.correlation_pyramid(torch.randn(1, 4, 15, 20), torch.randn(1, 4, 15, 20), 2)
    assert odd[1].shape == (300, 1, 7, 10)
