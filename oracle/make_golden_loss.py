"""Generate tests/golden/loss_*.npz: run the UNMODIFIED reference loss code (models/utils/pose.py, models/utils/flow.py,
models/loss/{sequence_loss,point_matching_loss}.py via oracle/ref_shim.py) on seeded inputs and check the restatement in
oracle/loss_oracle.py against it.  Build container only (needs /root/reference).

    python oracle/make_golden_loss.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import loss_oracle as L            # noqa: E402
from oracle import ref_shim                    # noqa: E402
from oracle.make_golden import GOLDEN, digest, report   # noqa: E402


def case_loss(R, name, seed, batch, iters, h, w):
    print(f'case {name}: B={batch} iters={iters} {h}x{w}')
    c = L.make_loss_case(seed, batch, iters, h, w)
    sc = c['scene']
    max_flow = 400.
    # ---- reference: scflow_refiner.py:204-258 with the shipped loss configs (configs/refine_models/scflow.py:75-104)
    gt_flow_ref = R.pose.get_flow_from_delta_pose_and_depth(sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'],
                                                           sc['depth'], sc['internel_k'], invalid_num=max_flow)
    gt_flow_mine = L.gt_flow_from_poses(sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'], sc['depth'], sc['internel_k'], max_flow)
    report('gt_flow', gt_flow_ref, gt_flow_mine, 2e-3)
    filt_ref = R.flow.filter_flow_by_mask(gt_flow_ref.clone(), c['gt_mask'], invalid_num=max_flow)
    filt_mine = L.filter_flow_by_mask(gt_flow_ref.clone(), c['gt_mask'], max_flow)
    report('filter_flow_by_mask', filt_ref, filt_mine, 0.)
    S = R.sequence_loss
    flow_fn = S.SequenceLoss(dict(type='RAFTLoss', loss_weight=.1, max_flow=max_flow), gamma=0.8)
    mask_fn = S.SequenceLoss(dict(type='L1Loss', loss_weight=10.), gamma=0.8)
    sym_types = {f'cls_{k + 1}': 1 for k, s in enumerate(c['symmetric']) if s}
    pm = R.point_matching_loss.DisentanglePointMatchingLoss(symmetry_types=sym_types, mesh_diameter=c['diameters'],
                                                            use_perspective_shape=True, loss_type='l1', disentangle_z=True,
                                                            loss_weight=10.)
    points_list = [c['meshes'][int(l)] for l in sc['label']]
    seq_pose_ref = [pm(r, t, gt_r=c['gt_rot'], gt_t=c['gt_trs'], labels=sc['label'], points_list=points_list)
                    for r, t in zip(c['seq_rot'], c['seq_trs'])]
    loss_flow_ref, seq_flow_ref = flow_fn(c['seq_flow'], gt_flow=filt_ref, valid=c['rendered_mask'])
    occ = (torch.sum(filt_ref, dim=1) < max_flow).to(torch.float32)
    loss_mask_ref, seq_mask_ref = mask_fn([m.squeeze(1) for m in c['seq_mask']], gt_mask=occ, valid=c['rendered_mask'])
    loss_pose_ref = sum(0.8 ** (iters - i - 1) * l for i, l in enumerate(seq_pose_ref))
    mine = L.refiner_loss(c['seq_flow'], c['seq_rot'], c['seq_trs'], c['seq_mask'], sc['ref_rotation'], sc['ref_translation'],
                          c['gt_rot'], c['gt_trs'], sc['depth'], sc['internel_k'], c['rendered_mask'], c['gt_mask'], sc['label'],
                          points_list, c['symmetric'], c['diameters'])
    report('seq pose loss', torch.stack(seq_pose_ref), mine['seq_pose'], 1e-5)
    report('seq flow loss', torch.stack(seq_flow_ref), mine['seq_flow'], 1e-6)
    report('seq mask loss', torch.stack(seq_mask_ref), mine['seq_mask'], 1e-6)
    total_ref = loss_pose_ref + loss_flow_ref + loss_mask_ref
    report('total loss', total_ref, mine['loss'], 1e-4)
    out = {'meta/seed': np.int64(seed), 'meta/batch': np.int64(batch), 'meta/iters': np.int64(iters), 'meta/h': np.int64(h), 'meta/w': np.int64(w)}
    for k, v in (('gt_flow', gt_flow_ref), ('gt_flow_filtered', filt_ref), ('seq_pose', torch.stack(seq_pose_ref)),
                 ('seq_flow', torch.stack(seq_flow_ref)), ('seq_mask', torch.stack(seq_mask_ref)), ('loss', total_ref.reshape(1)),
                 ('loss_pose', loss_pose_ref.reshape(1)), ('loss_flow', loss_flow_ref.reshape(1)), ('loss_mask', loss_mask_ref.reshape(1))):
        out.update(digest(k, v))
    np.savez_compressed(os.path.join(GOLDEN, name + '.npz'), **out)
    print(f'  total loss {float(total_ref):.6f} (pose {float(loss_pose_ref):.6f} flow {float(loss_flow_ref):.6f} mask {float(loss_mask_ref):.6f})')


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    R = ref_shim.load_reference()
    case_loss(R, 'loss_256_b4_it3', seed=11, batch=4, iters=3, h=256, w=256)
    case_loss(R, 'loss_96x128_b6_it2', seed=12, batch=6, iters=2, h=96, w=128)
    print('loss fixtures written')


if __name__ == '__main__':
    main()
