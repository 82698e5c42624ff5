"""Generate tests/golden/train_grad_b2_it3.npz: loss and GRADIENTS of the unmodified reference (SCFlowDecoder.forward with the
shipped detach_* configuration + the reference's own loss modules, via oracle/ref_shim.py) under autograd on seeded inputs, and
check the oracle's autograd graph (oracle/scflow_oracle.py with detach=True + oracle/loss_oracle.py) against it.  Build
container only (needs /root/reference).

    python oracle/make_golden_train.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import loss_oracle as L            # noqa: E402
from oracle import scflow_oracle as O          # noqa: E402
from oracle import ref_shim                    # noqa: E402
from oracle.make_golden import GOLDEN, build_ref_decoder, digest, report   # noqa: E402


def train_case(seed: int, batch: int, iters: int):
    """Seeded inputs of the gradient-parity case (shared with tests/test_train.py)."""
    c = L.make_loss_case(seed, batch, iters)
    c['scene']['label'][:] = torch.tensor(([12, 3] * batch)[:batch])          # class 12 is symmetric (nearest-neighbour matching)
    return c, O.make_features(seed, batch), O.make_decoder_weights(seed)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    R = ref_shim.load_reference()
    seed, b, iters = 5, 2, 3
    c, f, sd = train_case(seed, b, iters)
    sc = c['scene']
    max_flow = 400.
    # ---- the reference under autograd
    dec = build_ref_decoder(R, sd, iters).train()
    fr = {k: v.clone().requires_grad_(True) for k, v in f.items()}
    outs = dec(fr['feat_render'], fr['feat_real'], fr['h_feat'], fr['cxt_feat'], sc['ref_rotation'], sc['ref_translation'], sc['depth'],
               sc['internel_k'], label=sc['label'], init_flow=torch.zeros(b, 2, 256, 256), invalid_flow_num=0.)
    gt_flow = R.pose.get_flow_from_delta_pose_and_depth(sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'], sc['depth'],
                                                       sc['internel_k'], invalid_num=max_flow)
    gt_flow = R.flow.filter_flow_by_mask(gt_flow, c['gt_mask'], invalid_num=max_flow)
    S = R.sequence_loss
    flow_fn = S.SequenceLoss(dict(type='RAFTLoss', loss_weight=.1, max_flow=max_flow), gamma=0.8)
    mask_fn = S.SequenceLoss(dict(type='L1Loss', loss_weight=10.), gamma=0.8)
    sym_types = {f'cls_{k + 1}': 1 for k, s in enumerate(c['symmetric']) if s}
    pm = R.point_matching_loss.DisentanglePointMatchingLoss(symmetry_types=sym_types, mesh_diameter=c['diameters'],
                                                            use_perspective_shape=True, loss_type='l1', disentangle_z=True, loss_weight=10.)
    points_list = [c['meshes'][int(l)] for l in sc['label']]
    seq_pose = [pm(r, t, gt_r=c['gt_rot'], gt_t=c['gt_trs'], labels=sc['label'], points_list=points_list) for r, t in zip(outs[2], outs[3])]
    loss_flow, _ = flow_fn(outs[1], gt_flow=gt_flow, valid=c['rendered_mask'])
    occ = (torch.sum(gt_flow, dim=1) < max_flow).to(torch.float32)
    loss_mask, _ = mask_fn([m.squeeze(1) for m in outs[4]], gt_mask=occ, valid=c['rendered_mask'])
    loss_pose = sum(0.8 ** (iters - i - 1) * l for i, l in enumerate(seq_pose))
    loss = loss_pose + loss_flow + loss_mask
    loss.backward()
    ref_grads = {k: p.grad for k, p in dec.named_parameters()}
    # ---- the oracle under autograd
    sd_o = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    fo = {k: v.clone().requires_grad_(True) for k, v in f.items()}
    o = O.decoder_forward(sd_o, fo['feat_render'], fo['feat_real'], fo['h_feat'], fo['cxt_feat'], sc['ref_rotation'], sc['ref_translation'],
                          sc['depth'], sc['internel_k'], sc['label'], torch.zeros(b, 2, 256, 256), 0., iters=iters, detach=True)
    mine = L.refiner_loss(o[1], o[2], o[3], o[4], sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'], sc['depth'],
                          sc['internel_k'], c['rendered_mask'], c['gt_mask'], sc['label'], points_list, c['symmetric'], c['diameters'])
    mine['loss'].backward()
    report('loss', loss.detach(), mine['loss'].detach(), 1e-4)
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(b), 'meta/iters': np.int64(iters)}
    out.update(digest('loss', loss.detach().reshape(1)))
    worst = 0.
    for k, g in ref_grads.items():
        assert g is not None and sd_o[k].grad is not None, k
        scale = float(g.abs().max())
        rel = float((g - sd_o[k].grad).abs().max()) / max(scale, 1e-12)
        worst = max(worst, rel)
        if rel > 2e-3:
            raise SystemExit(f'oracle gradient of {k} disagrees with the reference: relative {rel:.3e}')
        out.update(digest('grad/' + k, g))
    for k in fr:
        rel = float((fr[k].grad - fo[k].grad).abs().max()) / max(float(fr[k].grad.abs().max()), 1e-12)
        worst = max(worst, rel)
        if rel > 2e-3:
            raise SystemExit(f'oracle gradient of input {k} disagrees with the reference: relative {rel:.3e}')
        out.update(digest('grad_in/' + k, fr[k].grad))
    print(f'  loss {float(loss):.6f}; worst relative gradient difference reference vs oracle {worst:.2e} over {len(ref_grads)} parameters')
    np.savez_compressed(os.path.join(GOLDEN, 'train_grad_b2_it3.npz'), **out)
    print('gradient fixture written')


if __name__ == '__main__':
    main()
