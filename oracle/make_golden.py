"""Generate tests/golden/*.npz by running the UNMODIFIED reference (via oracle/ref_shim.py) on seeded inputs,
and check the restatement in oracle/scflow_oracle.py against it.  Runs only in the build container
(needs /root/reference); the fixtures it writes are committed and travel to the GPU box.

    python oracle/make_golden.py            # regenerate + validate

Fixture format (kept small): every tensor is stored as a *digest* = values at a seeded random subset of
<= 4096 flat indices, plus float64 sum / abs-sum and the shape.  Inputs are never stored: they are
re-derived from the seeds by oracle.scflow_oracle.make_* (pure CPU torch RNG => identical everywhere).
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import scflow_oracle as O          # noqa: E402
from oracle import ref_shim                    # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MAX_PTS = 4096


def digest_indices(name: str, numel: int) -> np.ndarray:
    if numel <= MAX_PTS:
        return np.arange(numel, dtype=np.int64)
    rng = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
    return np.sort(rng.choice(numel, MAX_PTS, replace=False)).astype(np.int64)


def digest(name: str, t: torch.Tensor) -> dict:
    a = t.detach().cpu().double().numpy().reshape(-1)
    idx = digest_indices(name, a.size)
    return {name + '/vals': a[idx].astype(np.float32 if t.dtype != torch.float64 else np.float64),
            name + '/sum': np.float64(a.sum()), name + '/asum': np.float64(np.abs(a).sum()),
            name + '/shape': np.asarray(t.shape, dtype=np.int64)}


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())


def report(tag, ref, mine, tol):
    d = maxdiff(ref, mine)
    flag = 'OK ' if d <= tol else 'BAD'
    print(f'  [{flag}] {tag:34s} max|ref-oracle| = {d:.3e} (tol {tol:.0e})')
    if d > tol:
        raise SystemExit(f'oracle restatement disagrees with the reference on {tag}')


class _IdentityHead(torch.nn.Module):
    """config-3 stand-in for the pose head (stock head cannot run at 480x640)."""

    def forward(self, x, label):
        b = x.shape[0]
        return torch.tensor([1., 0., 0., 0., 1., 0.]).repeat(b, 1), torch.zeros(b, 3)


def build_ref_decoder(R, sd, iters, identity_head=False):
    dec = R.SCFlowDecoder(**ref_shim.decoder_cfg(iters=iters)).eval()
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    if identity_head:
        dec.pose_pred = _IdentityHead()
    return dec


def case_decoder(R, name, seed, batch, h, w, iters, identity_head=False):
    print(f'case {name}: B={batch} {h}x{w} iters={iters}')
    scene = O.make_scene(seed, batch, h, w)
    feats = O.make_features(seed, batch, h // 8, w // 8)
    sd = O.make_decoder_weights(seed)
    init_flow = torch.zeros(batch, 2, h, w)
    dec = build_ref_decoder(R, sd, iters, identity_head)
    with torch.no_grad():
        ref = dec(feats['feat_render'], feats['feat_real'], feats['h_feat'], feats['cxt_feat'],
                  scene['ref_rotation'], scene['ref_translation'], scene['depth'], scene['internel_k'],
                  label=scene['label'], init_flow=init_flow, invalid_flow_num=0.)
        trace = {}
        mine = O.decoder_forward(sd, feats['feat_render'], feats['feat_real'], feats['h_feat'], feats['cxt_feat'],
                                 scene['ref_rotation'], scene['ref_translation'], scene['depth'], scene['internel_k'],
                                 scene['label'], init_flow, 0., iters=iters, identity_pose_head=identity_head,
                                 trace=trace)
    names = ['flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation']
    tols = [2e-3, 2e-4, 1e-5, 2e-3, 1e-5, 1e-5, 1e-5]
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(batch), 'meta/h': np.int64(h), 'meta/w': np.int64(w),
           'meta/iters': np.int64(iters), 'meta/identity_head': np.int64(identity_head)}
    for nm, tol, rl, ml in zip(names, tols, ref, mine):
        for i, (r_, m_) in enumerate(zip(rl, ml)):
            report(f'{nm}[{i}]', r_, m_, tol)
            out.update(digest(f'{nm}/{i}', r_))
    # stage-level intermediates of the ORACLE (already tied to the reference through the outputs above, and
    # individually below for iteration 0) - these pin the per-kernel GPU tests.
    with torch.no_grad():
        pyr_ref = dec.corr_block(feats['feat_render'], feats['feat_real'])
        pyr = O.correlation_pyramid(feats['feat_render'], feats['feat_real'])
        for l, (a, b_) in enumerate(zip(pyr_ref, pyr)):
            report(f'pyramid[{l}]', a, b_, 1e-5)
            out.update(digest(f'pyramid/{l}', a))
        flow8 = trace['flow8'][min(1, iters - 1)]
        c_ref = dec.corr_lookup(pyr_ref, flow8.clone())
        report('lookup(grid_sample)', c_ref, O.corr_lookup(pyr, flow8), 1e-6)
        report('lookup(explicit gather)', c_ref, O.corr_lookup_explicit(pyr, flow8), 2e-5)
        m_ref = dec.encoder(c_ref, flow8)
        report('motion_encoder', m_ref, O.motion_encoder(sd, c_ref, flow8), 1e-5)
        x = torch.cat([feats['cxt_feat'], m_ref], 1)
        h_ref = dec.gru(feats['h_feat'], x)
        report('sepconv_gru', h_ref, O.sepconv_gru(sd, feats['h_feat'], x), 1e-5)
        report('flow_head', dec.flow_pred(h_ref), O.xhead(sd, 'flow_pred.', h_ref, 'flow'), 1e-5)
        report('mask_head', dec.mask_pred(h_ref), O.xhead(sd, 'mask_pred.', h_ref, 'mask'), 1e-5)
    for key in ('flow8', 'corr', 'motion', 'h', 'd_flow', 'mask8'):
        for i, t in enumerate(trace[key]):
            out.update(digest(f'trace/{key}/{i}', t))
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)


def case_lookup_ramp(R):
    """KAT (i): ramp volume corr[q, y, x] = 100*y + x  =>  channel order is x-major."""
    print('case lookup_ramp')
    h = w = 8
    vol = (100. * torch.arange(h).view(h, 1) + torch.arange(w).view(1, w)).float()
    vol = vol.view(1, 1, h, w).repeat(h * w, 1, 1, 1)
    flow = torch.zeros(1, 2, h, w)
    ref = R.CorrLookup(radius=1, align_corners=True)([vol], flow.clone())
    mine = O.corr_lookup([vol], flow, radius=1)
    report('ramp lookup', ref, mine, 0.0)
    got = ref[0, :, 4, 3].tolist()          # query (y=4, x=3)
    assert got[:4] == [302., 402., 502., 303.], got
    # fractional flow + out-of-bounds taps
    g = torch.Generator().manual_seed(5)
    flow2 = 3.0 * torch.randn(1, 2, h, w, generator=g)
    pyr = [vol, torch.nn.functional.avg_pool2d(vol, 2, 2)]
    ref2 = R.CorrLookup(radius=2, align_corners=True)(pyr, flow2.clone())
    report('ramp lookup frac r=2', ref2, O.corr_lookup(pyr, flow2, radius=2), 1e-6)
    report('ramp lookup frac explicit', ref2, O.corr_lookup_explicit(pyr, flow2, radius=2), 1e-4)
    out = {}
    out.update(digest('ramp/r1', ref))
    out.update(digest('ramp/frac_r2', ref2))
    np.savez_compressed(os.path.join(GOLDEN, 'lookup_ramp.npz'), **out)


def case_pose_head(R):
    """KAT (iii): mixed labels [5,7,9] behave like [5,5,5]; plus head outputs on seeded input."""
    print('case pose_head')
    sd = O.make_decoder_weights(3)
    head = R.MultiClassPoseHead(num_class=21, in_channels=224, net_type='Basic', rotation_mode='ortho6d',
                                norm_cfg=dict(type='GN', num_groups=32, requires_grad=True), act_cfg=dict(type='ReLU')).eval()
    head.load_state_dict({k[len('pose_pred.'):]: v for k, v in sd.items() if k.startswith('pose_pred.')}, strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 224, 32, 32, generator=g)
    with torch.no_grad():
        r1, t1 = head(x, torch.tensor([5, 7, 9]))
        r2, t2 = head(x, torch.tensor([5, 5, 5]))
        mr, mt = O.pose_head(sd, x, torch.tensor([5, 7, 9]))
    assert torch.equal(r1, r2) and torch.equal(t1, t2)
    report('pose_head rot', r1, mr, 1e-5)
    report('pose_head trans', t1, mt, 1e-5)
    out = {}
    out.update(digest('rot', r1))
    out.update(digest('trans', t1))
    np.savez_compressed(os.path.join(GOLDEN, 'pose_head.npz'), **out)


def case_geometry(R):
    """pose update + un-projection + re-projection against models/utils/pose.py, incl. KAT (iv)."""
    print('case geometry')
    scene = O.make_scene(21, 3, 64, 64)
    g = torch.Generator().manual_seed(2)
    d_rot = torch.tensor([1., 0., 0., 0., 1., 0.]) + 0.1 * torch.randn(3, 6, generator=g)
    d_trs = 0.1 * torch.randn(3, 3, generator=g)
    r_ref, t_ref = R.pose.get_pose_from_delta_pose(d_rot, d_trs, scene['ref_rotation'], scene['ref_translation'],
                                                   depth_transform='exp', detach_depth_for_xy=True)
    r_m, t_m = O.update_pose(d_rot, d_trs, scene['ref_rotation'], scene['ref_translation'])
    report('update_pose R', r_ref, r_m, 1e-6)
    report('update_pose t', t_ref, t_m, 1e-4)
    p2, p3 = [], []
    for i in range(3):
        a, b_ = R.pose.cal_3d_2d_corr(scene['depth'][i], scene['internel_k'][i], scene['ref_rotation'][i], scene['ref_translation'][i])
        p2.append(a)
        p3.append(b_)
    f_ref = R.pose.get_flow_from_delta_pose_and_points(r_ref, t_ref, scene['internel_k'], p2, p3, 64, 64, invalid_num=0.)
    pts = O.unproject_dense(scene['depth'], scene['internel_k'], scene['ref_rotation'], scene['ref_translation'])
    f_m = O.reproject_dense(pts, scene['depth'], scene['internel_k'], r_m, t_m, 0.)
    report('pose flow', f_ref, f_m, 2e-3)
    f_id = O.reproject_dense(pts, scene['depth'], scene['internel_k'], scene['ref_rotation'], scene['ref_translation'], 0.)
    f_id_ref = R.pose.get_flow_from_delta_pose_and_points(scene['ref_rotation'], scene['ref_translation'], scene['internel_k'], p2, p3, 64, 64, invalid_num=0.)
    assert float(f_id_ref.abs().max()) < 2e-3, 'identity delta pose must give ~zero flow'
    report('identity pose flow', f_id_ref, f_id, 2e-3)
    out = {}
    for k, v in (('rot', r_ref), ('trans', t_ref), ('flow', f_ref)):
        out.update(digest(k, v))
    np.savez_compressed(os.path.join(GOLDEN, 'geometry.npz'), **out)


def case_get_pose(R, name, seed, batch, iters):
    """BASELINE config 1: images -> encoders -> decoder (scflow_refiner.py:88-142) on the reference modules."""
    print(f'case {name}: B={batch} iters={iters}')
    scene = O.make_scene(seed, batch)
    sd = O.make_model_weights(seed)
    enc = R.RAFTEncoder(**ref_shim.encoder_cfg('IN')).eval()
    ctx = R.RAFTEncoder(**ref_shim.encoder_cfg('BN')).eval()
    enc.load_state_dict({k[len('render_encoder.'):]: v for k, v in sd.items() if k.startswith('render_encoder.')}, strict=True)
    m, u = ctx.load_state_dict({k[len('context.'):]: v for k, v in sd.items() if k.startswith('context.')}, strict=False)
    assert not u and all(x.endswith('num_batches_tracked') for x in m), (m, u)
    dec = build_ref_decoder(R, {k[len('decoder.'):]: v for k, v in sd.items() if k.startswith('decoder.')}, iters)
    with torch.no_grad():
        f_real = enc(scene['real_images'])
        f_render = enc(scene['render_images'])
        c = ctx(scene['render_images'])
        h_feat, cxt = torch.split(c, [128, 128], dim=1)
        h_feat, cxt = torch.tanh(h_feat), torch.relu(cxt)
        ref = dec(f_render, f_real, h_feat, cxt, scene['ref_rotation'], scene['ref_translation'], scene['depth'],
                  scene['internel_k'], label=scene['label'], init_flow=torch.zeros(batch, 2, 256, 256), invalid_flow_num=0.)
        mine = O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                          scene['depth'], scene['internel_k'], scene['label'], iters=iters)
        enc_sd = {k[len('render_encoder.'):]: v for k, v in sd.items() if k.startswith('render_encoder.')}
        ctx_sd = {k[len('context.'):]: v for k, v in sd.items() if k.startswith('context.')}
        report('encoder IN', f_real, O.raft_encoder(enc_sd, scene['real_images'], 'IN'), 1e-4)
        report('encoder BN', c, O.raft_encoder(ctx_sd, scene['render_images'], 'BN'), 1e-4)
    names = ['flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation']
    tols = [5e-3, 5e-4, 1e-5, 5e-3, 1e-5, 1e-5, 1e-5]
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(batch), 'meta/iters': np.int64(iters)}
    out.update(digest('feat_real', f_real))
    out.update(digest('feat_render', f_render))
    out.update(digest('context', c))
    for nm, tol, rl, ml in zip(names, tols, ref, mine):
        for i, (r_, m_) in enumerate(zip(rl, ml)):
            report(f'{nm}[{i}]', r_, m_, tol)
            out.update(digest(f'{nm}/{i}', r_))
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    os.makedirs(GOLDEN, exist_ok=True)
    R = ref_shim.load_reference()
    case_lookup_ramp(R)
    case_pose_head(R)
    case_geometry(R)
    case_decoder(R, 'decoder_256_b2_it4', seed=1, batch=2, h=256, w=256, iters=4)
    case_decoder(R, 'decoder_256_b3_it8', seed=2, batch=3, h=256, w=256, iters=8)
    case_decoder(R, 'decoder_480x640_b1_it2', seed=3, batch=1, h=480, w=640, iters=2, identity_head=True)
    case_get_pose(R, 'get_pose_256_b1_it4', seed=0, batch=1, iters=4)
    print('all golden fixtures written to', GOLDEN)


if __name__ == '__main__':
    main()
