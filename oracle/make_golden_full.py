"""Full-size golden fixtures (BASELINE config 2: B=32, 256x256, 8 iterations) from the UNMODIFIED reference run here
through oracle/ref_shim.py - the same recipe as oracle/make_golden.py, outputs only (no stage traces) to keep the files
small.  Runs only in the build container (needs /root/reference); the fixtures travel to the GPU box.

    python oracle/make_golden_full.py
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import scflow_oracle as O          # noqa: E402
from oracle import ref_shim                    # noqa: E402
from oracle.make_golden import GOLDEN, build_ref_decoder, digest, report   # noqa: E402

NAMES = ['flow_from_pose', 'flow_from_pred', 'rotation', 'translation', 'mask', 'delta_rotation', 'delta_translation']


def case_decoder_full(R, name, seed, batch, iters):
    print(f'case {name}: B={batch} iters={iters}')
    scene, feats = O.make_scene(seed, batch), O.make_features(seed, batch)
    sd = O.make_decoder_weights(seed)
    init_flow = torch.zeros(batch, 2, 256, 256)
    dec = build_ref_decoder(R, sd, iters)
    t0 = time.time()
    with torch.no_grad():
        ref = dec(feats['feat_render'], feats['feat_real'], feats['h_feat'], feats['cxt_feat'], scene['ref_rotation'],
                  scene['ref_translation'], scene['depth'], scene['internel_k'], label=scene['label'], init_flow=init_flow,
                  invalid_flow_num=0.)
        t1 = time.time()
        mine = O.decoder_forward(sd, feats['feat_render'], feats['feat_real'], feats['h_feat'], feats['cxt_feat'],
                                 scene['ref_rotation'], scene['ref_translation'], scene['depth'], scene['internel_k'],
                                 scene['label'], init_flow, 0., iters=iters)
    print(f'  reference {t1 - t0:.1f} s, oracle {time.time() - t1:.1f} s')
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(batch), 'meta/iters': np.int64(iters)}
    for nm, tol, rl, ml in zip(NAMES, [3e-3, 3e-4, 1e-5, 3e-3, 1e-5, 1e-5, 1e-5], ref, mine):
        for i, (r_, m_) in enumerate(zip(rl, ml)):
            report(f'{nm}[{i}]', r_, m_, tol)
            out.update(digest(f'{nm}/{i}', r_))
        # whole-tensor statistic of the LAST iteration: mean end-point magnitude per sample (flows) / the tensor itself (poses)
    out['epe_ref_vs_oracle'] = np.float64((ref[1][-1] - mine[1][-1]).pow(2).sum(1).sqrt().mean())
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)


def case_get_pose_full(R, name, seed, batch, iters):
    """Images -> 3 encoder passes -> loop on the reference's own modules (scflow_refiner.py:88-142)."""
    print(f'case {name}: B={batch} iters={iters}')
    scene = O.make_scene(seed, batch)
    sd = O.make_model_weights(seed)
    enc = R.RAFTEncoder(**ref_shim.encoder_cfg('IN')).eval()
    ctx = R.RAFTEncoder(**ref_shim.encoder_cfg('BN')).eval()
    enc.load_state_dict({k[len('render_encoder.'):]: v for k, v in sd.items() if k.startswith('render_encoder.')}, strict=True)
    m, u = ctx.load_state_dict({k[len('context.'):]: v for k, v in sd.items() if k.startswith('context.')}, strict=False)
    assert not u and all(x.endswith('num_batches_tracked') for x in m), (m, u)
    dec = build_ref_decoder(R, {k[len('decoder.'):]: v for k, v in sd.items() if k.startswith('decoder.')}, iters)
    t0 = time.time()
    with torch.no_grad():
        f_real, f_render, c = enc(scene['real_images']), enc(scene['render_images']), ctx(scene['render_images'])
        h_feat, cxt = torch.split(c, [128, 128], dim=1)
        ref = dec(f_render, f_real, torch.tanh(h_feat), torch.relu(cxt), scene['ref_rotation'], scene['ref_translation'],
                  scene['depth'], scene['internel_k'], label=scene['label'], init_flow=torch.zeros(batch, 2, 256, 256),
                  invalid_flow_num=0.)
        t1 = time.time()
        mine = O.get_pose(sd, scene['render_images'], scene['real_images'], scene['ref_rotation'], scene['ref_translation'],
                          scene['depth'], scene['internel_k'], scene['label'], iters=iters)
    print(f'  reference {t1 - t0:.1f} s, oracle {time.time() - t1:.1f} s')
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(batch), 'meta/iters': np.int64(iters)}
    for nm, tol, rl, ml in zip(NAMES, [5e-3, 5e-4, 1e-5, 5e-3, 1e-5, 1e-5, 1e-5], ref, mine):
        report(f'{nm}[-1]', rl[-1], ml[-1], tol)
        out.update(digest(f'{nm}/{iters - 1}', rl[-1]))
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    R = ref_shim.load_reference()
    case_decoder_full(R, 'decoder_256_b32_it8', seed=7, batch=32, iters=8)
    case_get_pose_full(R, 'get_pose_256_b32_it8', seed=5, batch=32, iters=8)
    print('full-size golden fixtures written to', GOLDEN)


if __name__ == '__main__':
    main()
