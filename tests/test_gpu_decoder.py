"""GPU: the whole decoder loop / refiner against the golden fixtures (generated from the reference itself), the
live oracle, and size-independent properties at BASELINE config-2 size (B=32, 8 iterations)."""
import pytest
import torch

from oracle import scflow_oracle as O
from tests.util import (assert_matches_digest, build_decoder_from_oracle_weights, load_golden, scflow_model_cfg)

pytestmark = pytest.mark.gpu

NAMES = ['flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation']
# tolerances vs the fp32 CPU reference (px, px, -, mm, -, -, -). North star: flow EPE within 1e-3.
TOL_FP32 = dict(flow_from_pose=3e-3, flow_from_pred=1e-3, rotation=2e-6, translation=3e-3, mask=2e-5, delta_rotation=5e-6,
                delta_translation=5e-6)
# per-element tolerances of the tcgen05 split-bf16 path (each product carries ~2^-16 relative error; the loop feeds
# errors back through pose -> flow -> lookup). The mean-EPE bar (1e-3 px) is asserted separately below.
TOL_BF16X3 = dict(flow_from_pose=2e-2, flow_from_pred=2e-2, rotation=2e-5, translation=5e-2, mask=1e-3, delta_rotation=1e-4,
                  delta_translation=1e-4)
TOLS = {0: TOL_FP32, 1: TOL_BF16X3}


def _inputs(seed, b, h, w):
    scene = O.make_scene(seed, b, h, w)
    f = O.make_features(seed, b, h // 8, w // 8)
    return scene, f


def _call(dec, scene, f, b, h, w):
    c = lambda t: t.cuda()
    with torch.no_grad():
        return dec(c(f['feat_render']), c(f['feat_real']), c(f['h_feat']), c(f['cxt_feat']), c(scene['ref_rotation']),
                   c(scene['ref_translation']), c(scene['depth']), c(scene['internel_k']), label=c(scene['label']),
                   init_flow=torch.zeros(b, 2, h, w, device='cuda'), invalid_flow_num=0.)


def _epe(a, b):
    return float((a - b).pow(2).sum(1).sqrt().mean())


@pytest.mark.parametrize('precision', [0, 1])
@pytest.mark.parametrize('name', ['decoder_256_b2_it4', 'decoder_256_b3_it8'])
def test_decoder_matches_reference_golden(name, precision):
    g = load_golden(name)
    seed, b, h, w, iters = (int(g['meta/' + k]) for k in ('seed', 'batch', 'h', 'w', 'iters'))
    dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=precision)
    scene, f = _inputs(seed, b, h, w)
    outs = _call(dec, scene, f, b, h, w)
    assert len(outs) == 7 and all(len(o) == iters for o in outs)
    worst = {}
    for nm, lst in zip(NAMES, outs):
        for i, t in enumerate(lst):
            worst[nm] = max(worst.get(nm, 0.), assert_matches_digest(g, f'{nm}/{i}', t, atol=TOLS[precision][nm]))
    print(f'{name} precision={precision} worst abs err:', {k: f'{v:.2e}' for k, v in worst.items()})


@pytest.mark.parametrize('precision', [0, 1])
def test_decoder_480x640_identity_head_golden(precision):
    g = load_golden('decoder_480x640_b1_it2')
    seed, b, h, w, iters = (int(g['meta/' + k]) for k in ('seed', 'batch', 'h', 'w', 'iters'))
    dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=precision)
    dec.identity_pose_head = True
    scene, f = _inputs(seed, b, h, w)
    outs = _call(dec, scene, f, b, h, w)
    for nm, lst in zip(NAMES, outs):
        for i, t in enumerate(lst):
            assert_matches_digest(g, f'{nm}/{i}', t, atol=TOLS[precision][nm])
    # the stock head cannot run here, exactly like the reference
    dec.identity_pose_head = False
    from scflow_b200 import ScfError
    with pytest.raises(ScfError, match='32x32'):
        _call(dec, scene, f, b, h, w)


@pytest.mark.parametrize('precision', [0, 1])
def test_decoder_flow_epe_vs_live_oracle(precision):
    """The north-star number: mean end-point error of every iteration's flows against the fp32 CPU reference path."""
    seed, b, iters = 7, 2, 8
    dec, sd = build_decoder_from_oracle_weights(seed, iters, precision=precision)
    scene, f = _inputs(seed, b, 256, 256)
    outs = _call(dec, scene, f, b, 256, 256)
    with torch.no_grad():
        ref = O.decoder_forward(sd, f['feat_render'], f['feat_real'], f['h_feat'], f['cxt_feat'], scene['ref_rotation'],
                                scene['ref_translation'], scene['depth'], scene['internel_k'], scene['label'],
                                torch.zeros(b, 2, 256, 256), 0., iters=iters)
    print(f'precision={precision}: pose-flow EPE/iter', [f'{_epe(outs[0][i].cpu(), ref[0][i]):.2e}' for i in range(iters)],
          'pred-flow EPE/iter', [f'{_epe(outs[1][i].cpu(), ref[1][i]):.2e}' for i in range(iters)],
          'max|dR|', f'{max(float((outs[2][i].cpu() - ref[2][i]).abs().max()) for i in range(iters)):.2e}',
          'max|dt| mm', f'{max(float((outs[3][i].cpu() - ref[3][i]).abs().max()) for i in range(iters)):.2e}')
    for i in range(iters):
        assert _epe(outs[0][i].cpu(), ref[0][i]) < 1e-3, f'pose-flow EPE at iteration {i}'
        assert _epe(outs[1][i].cpu(), ref[1][i]) < 1e-3, f'pred-flow EPE at iteration {i}'
        assert float((outs[2][i].cpu() - ref[2][i]).abs().max()) < (1e-5 if precision == 0 else 2e-5)
        assert float((outs[3][i].cpu() - ref[3][i]).abs().max()) < (5e-3 if precision == 0 else 5e-2)       # mm


@pytest.mark.parametrize('precision', [0, 1])
def test_cuda_graph_replay_is_bit_identical(precision):
    seed, b, iters = 9, 2, 3
    dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=precision)
    scene, f = _inputs(seed, b, 256, 256)
    eager = [[t.clone() for t in lst] for lst in _call(dec, scene, f, b, 256, 256)]
    dec.use_cuda_graph = True
    for _ in range(2):      # capture, then replay
        graphed = _call(dec, scene, f, b, 256, 256)
        for a, g_ in zip(eager, graphed):
            for x, y in zip(a, g_):
                assert torch.equal(x, y)


def test_refiner_get_pose_config1_golden():
    """BASELINE config 1: images -> encoders (stock cuDNN, fp32) -> decoder, vs the reference run on CPU."""
    import scflow_b200 as S
    g = load_golden('get_pose_256_b1_it4')
    seed, b, iters = int(g['meta/seed']), int(g['meta/batch']), int(g['meta/iters'])
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False         # the oracle is fp32; keep the (unreplaced) encoders fp32 too
    try:
        model = S.build_refiner(scflow_model_cfg(iters=iters))
        model.load_state_dict(O.make_model_weights(seed), strict=False)
        model = model.cuda().eval()
        scene = O.make_scene(seed, b)
        c = {k: v.cuda() for k, v in scene.items()}
        with torch.no_grad():
            outs = model.get_pose(c['render_images'], c['real_images'], c['ref_rotation'], c['ref_translation'], c['depth'],
                                  c['internel_k'], c['label'])
            res = model.forward_single_pass(dict(rendered_images=c['render_images'], real_images=c['real_images'],
                                                 ref_rotations=c['ref_rotation'], ref_translations=c['ref_translation'],
                                                 rendered_depths=c['depth'], internel_k=c['internel_k'], labels=c['label'],
                                                 per_img_patch_num=[b]))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    tol = dict(TOL_FP32, flow_from_pose=2e-2, flow_from_pred=1e-2, translation=2e-2, rotation=2e-5, delta_rotation=5e-5,
               delta_translation=5e-5, mask=2e-4)   # encoder (cuDNN vs MKL-DNN) noise feeds the loop here
    for nm, lst in zip(NAMES, outs):
        for i, t in enumerate(lst):
            assert_matches_digest(g, f'{nm}/{i}', t, atol=tol[nm])
    assert torch.equal(res['rotations'][0], outs[2][-1]) and res['scores'][0].shape == (b,)


@pytest.mark.parametrize('precision', [0, 1])
def test_full_size_properties_b32(precision):
    """BASELINE config 2 size (B=32, 8 iterations): properties that need no oracle run."""
    seed, b, iters = 13, 32, 8
    dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=precision)
    scene, f = _inputs(seed, b, 256, 256)
    scene['label'][:] = 4
    outs = _call(dec, scene, f, b, 256, 256)
    again = _call(dec, scene, f, b, 256, 256)
    for a, b_ in zip(outs, again):                                   # deterministic
        assert all(torch.equal(x, y) for x, y in zip(a, b_))
    for lst in outs:
        assert all(torch.isfinite(t).all() for t in lst)
    rot = outs[2][-1]
    eye = torch.bmm(rot, rot.transpose(1, 2))
    assert float((eye - torch.eye(3, device='cuda')).abs().max()) < 1e-4      # rotations stay orthonormal
    bg = (scene['depth'] <= 0).cuda()
    assert float(outs[0][-1][:, 0][bg].abs().max()) == 0.                     # invalid_flow_num=0 on background
    assert 0. <= float(outs[4][-1].min()) and float(outs[4][-1].max()) <= 1.  # sigmoid mask
    # samples are independent: a sub-batch (same label[0]) gives the same poses
    sub = {k: (v[8:12].clone() if isinstance(v, torch.Tensor) else v) for k, v in scene.items()}
    fsub = {k: v[8:12].clone() for k, v in f.items()}
    o2 = _call(dec, sub, fsub, 4, 256, 256)
    assert float((o2[2][-1] - outs[2][-1][8:12]).abs().max()) < 1e-5
    assert float((o2[3][-1] - outs[3][-1][8:12]).abs().max()) < 1e-2
    if precision == 1:      # the two arithmetic paths agree at full size too
        base, _ = build_decoder_from_oracle_weights(seed, iters, precision=0)
        o0 = _call(base, scene, f, b, 256, 256)
        assert _epe(outs[1][-1], o0[1][-1]) < 1e-3 and _epe(outs[0][-1], o0[0][-1]) < 1e-3


def test_mask_options_run():
    """mask_corr / mask_flow decoder options (off in the shipped config) against the oracle-equivalent formulas."""
    import scflow_b200 as S
    seed, b, iters = 5, 1, 2
    cfg = scflow_model_cfg(iters)['decoder']
    cfg.update(mask_corr=True, mask_flow=True)
    dec = S.build_decoder(cfg)
    sd = O.make_decoder_weights(seed)
    dec.load_state_dict(sd, strict=True)
    dec = dec.cuda().eval()
    scene, f = _inputs(seed, b, 256, 256)
    outs = _call(dec, scene, f, b, 256, 256)
    assert all(torch.isfinite(t).all() for lst in outs for t in lst)
    base, _ = build_decoder_from_oracle_weights(seed, iters)
    ref = _call(base, scene, f, b, 256, 256)
    assert torch.equal(outs[2][0], ref[2][0])            # first iteration: mask is all ones -> identical
    assert not torch.equal(outs[2][1], ref[2][1])        # second iteration: the predicted mask gates corr / flow


@pytest.mark.parametrize('norm', ['IN', 'BN'])
def test_native_encoder_matches_oracle(norm):
    """RAFTEncoder through scf_encoder_forward (tcgen05 split-bf16) vs the fp32 CPU oracle."""
    import scflow_b200 as S
    enc = S.build_encoder(dict(type='RAFTEncoder', in_channels=3, out_channels=256, net_type='Basic', norm_cfg=dict(type=norm)))
    sd = O.make_encoder_weights(5, norm)
    enc.load_state_dict(sd, strict=False)
    enc = enc.cuda().eval()
    scene = O.make_scene(5, 3)
    x = scene['real_images']
    with torch.no_grad():
        got = enc(x.cuda()).cpu()
        ref = O.raft_encoder(sd, x, norm)
        enc.use_native = False
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        stock = enc(x.cuda()).cpu()
        torch.backends.cudnn.allow_tf32 = prev
    assert got.shape == ref.shape == (3, 256, 32, 32)
    scale = float(ref.abs().max())
    err, err_stock = float((got - ref).abs().max()), float((stock - ref).abs().max())
    print(f'encoder {norm}: max|native - oracle| {err:.2e}, max|cuDNN fp32 - oracle| {err_stock:.2e}, |ref|max {scale:.2f}')
    assert err < 2e-4 * max(scale, 1.0)
    # odd batch / non-square input
    x2 = torch.rand(1, 3, 64, 96)
    with torch.no_grad():
        enc.use_native = True
        g2 = enc(x2.cuda()).cpu()
        r2 = O.raft_encoder(sd, x2, norm)
    assert float((g2 - r2).abs().max()) < 2e-4 * max(float(r2.abs().max()), 1.0)
    # full frame (BASELINE config 3): three 128-pixel column strips per row, the last one ragged, in the rolling-rows kernels
    x3 = torch.rand(1, 3, 480, 640)
    with torch.no_grad():
        g3 = enc(x3.cuda()).cpu()
        r3 = O.raft_encoder(sd, x3, norm)
    assert g3.shape == r3.shape == (1, 256, 60, 80)
    assert float((g3 - r3).abs().max()) < 2e-4 * max(float(r3.abs().max()), 1.0)


@pytest.mark.parametrize('norm', ['IN', 'BN'])
def test_encoder_stem_variants_are_bit_identical(norm, monkeypatch):
    """The stem's three forms - generic tile, rolling rows over the im2col'd image, rolling rows x-folding the fp32 image in the
    kernel - use the same split-bf16 operands; the two rolling forms also share the summation order, so they agree bit for bit."""
    import scflow_b200 as S
    enc = S.build_encoder(dict(type='RAFTEncoder', in_channels=3, out_channels=256, net_type='Basic', norm_cfg=dict(type=norm)))
    enc.load_state_dict(O.make_encoder_weights(7, norm), strict=False)
    enc = enc.cuda().eval()
    x = O.make_scene(7, 4)['real_images'].cuda()
    outs = {}
    for name, env in (('fold', {}), ('im2col', {'SCFLOW_STEM_FOLD': '0'}), ('generic', {'SCFLOW_TC_ROWS_STEM': '0'})):
        for k in ('SCFLOW_STEM_FOLD', 'SCFLOW_TC_ROWS_STEM'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with torch.no_grad():
            outs[name] = enc(x).clone()
    assert torch.equal(outs['fold'], outs['im2col'])
    assert float((outs['fold'] - outs['generic']).abs().max()) < 1e-4 * float(outs['generic'].abs().max())


@pytest.mark.parametrize('graph', [False, True])
def test_native_feature_path_matches_generic_path_and_oracle(graph):
    """get_pose with the encoders writing the loop's inputs straight into the decoder workspace (no NCHW round trip) against
    the generic module-by-module path and against the CPU oracle (flow EPE < 1e-3 px, the stated tolerance).  EVERY call is
    checked: the first (eager), the second (eager + graph capture) and the third (graph replay) - a replay-only comparison
    would hide a capture call that consumed the encoder-written hidden state twice."""
    import scflow_b200 as S
    seed, b, iters = 4, 2, 3
    model = S.build_refiner(scflow_model_cfg(iters=iters, precision=1, use_cuda_graph=graph))
    sd = O.make_model_weights(seed)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    scene = O.make_scene(seed, b)
    c = {k: v.cuda() for k, v in scene.items()}
    args = (c['render_images'], c['real_images'], c['ref_rotation'], c['ref_translation'], c['depth'], c['internel_k'], c['label'])
    with torch.no_grad():
        assert model.native_feature_path
        calls = []
        for _ in range(3):
            calls.append([[t.clone() for t in lst] for lst in model.get_pose(*args)])
        assert bool(model._graphs) == graph, 'the whole step must be captured once the shape repeats (and only then)'
        model.native_feature_path = False
        slow = [[t.clone() for t in lst] for lst in model.get_pose(*args)]
        ref = O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                         scene['depth'], scene['internel_k'], scene['label'], iters=iters)
    for n, fast in enumerate(calls):
        for k in (0, 1):
            for i in range(iters):
                d = float((fast[k][i] - slow[k][i]).pow(2).sum(1).sqrt().mean())
                e = float((fast[k][i].cpu() - ref[k][i]).pow(2).sum(1).sqrt().mean())
                assert d < 1e-4, f'call {n}: native vs generic path: list {k} iter {i} EPE {d:.3e}'
                assert e < 1e-3, f'call {n}: native path vs oracle: list {k} iter {i} EPE {e:.3e}'
        assert float((fast[2][-1] - slow[2][-1]).abs().max()) < 1e-5
        for k in range(7):                       # eager, capture and replay calls agree bit for bit
            assert all(torch.equal(a, b_) for a, b_ in zip(fast[k], calls[0][k])), f'call {n} differs from call 0 in list {k}'


def test_graph_survives_shape_changes():
    """Patch counts change from batch to batch at test time: A, A, A (captured, replayed), B, A, A, B must each give the
    result of an eager model, whatever the graph cache holds."""
    import scflow_b200 as S
    seed, iters = 6, 2
    sd = O.make_model_weights(seed)
    models = {}
    for graph in (False, True):
        m = S.build_refiner(scflow_model_cfg(iters=iters, precision=1, use_cuda_graph=graph))
        m.load_state_dict(sd, strict=False)
        models[graph] = m.cuda().eval()
    scenes = {b: {k: v.cuda() for k, v in O.make_scene(seed + b, b).items()} for b in (2, 3)}

    def run(m, b):
        c = scenes[b]
        with torch.no_grad():
            outs = m.get_pose(c['render_images'], c['real_images'], c['ref_rotation'], c['ref_translation'], c['depth'],
                              c['internel_k'], c['label'])
        return [outs[k][-1].clone() for k in (0, 1, 2, 3, 4)]
    want = {b: run(models[False], b) for b in (2, 3)}
    for n, b in enumerate((2, 2, 2, 3, 2, 2, 3, 3, 3)):
        got = run(models[True], b)
        for g, w_ in zip(got, want[b]):
            assert torch.equal(g, w_), f'call {n} (batch {b}) differs from the eager model'


def test_decoder_graph_every_call_matches_eager():
    """SCFlowDecoder's own graph (generic API): eager first call, capture on the repeat, replay afterwards - all identical."""
    seed, b, iters = 8, 2, 2
    scene, f = O.make_scene(seed, b), O.make_features(seed, b)
    outs = {}
    for graph in (False, True):
        dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=1, use_cuda_graph=graph)
        res = []
        with torch.no_grad():
            for _ in range(3):
                o = dec(f['feat_render'].cuda(), f['feat_real'].cuda(), f['h_feat'].cuda(), f['cxt_feat'].cuda(),
                        scene['ref_rotation'].cuda(), scene['ref_translation'].cuda(), scene['depth'].cuda(), scene['internel_k'].cuda(),
                        label=scene['label'].cuda(), init_flow=torch.zeros(b, 2, 256, 256, device='cuda'), invalid_flow_num=0.)
                res.append([o[k][-1].clone() for k in range(7)])
        outs[graph] = res
    for n in range(3):
        for a, b_ in zip(outs[True][n], outs[False][0]):
            assert torch.equal(a, b_), f'graph call {n} differs from the eager decoder'
