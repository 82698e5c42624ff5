"""Training-loss forward: oracle vs the golden fixtures produced from the reference's own loss
code (CPU), and the CUDA kernels (scf_filter_flow_by_mask, scf_refiner_loss, SCFlowRefiner.loss) vs the oracle (GPU)."""
import os

import pytest
import torch

from oracle import loss_oracle as L
from oracle import scflow_oracle as O
from tests.util import assert_matches_digest as check_digest, load_golden, scflow_model_cfg

CASES = ['loss_256_b4_it3', 'loss_96x128_b6_it2']


def _case(name):
    g = load_golden(name)
    m = {k: int(g['meta/' + k]) for k in ('seed', 'batch', 'iters', 'h', 'w')}
    return g, m, L.make_loss_case(m['seed'], m['batch'], m['iters'], m['h'], m['w'])


def _oracle_loss(c):
    sc = c['scene']
    points_list = [c['meshes'][int(l)] for l in sc['label']]
    return L.refiner_loss(c['seq_flow'], c['seq_rot'], c['seq_trs'], c['seq_mask'], sc['ref_rotation'], sc['ref_translation'],
                          c['gt_rot'], c['gt_trs'], sc['depth'], sc['internel_k'], c['rendered_mask'], c['gt_mask'], sc['label'],
                          points_list, c['symmetric'], c['diameters'])


@pytest.mark.parametrize('name', CASES)
def test_loss_oracle_matches_reference_fixture(name):
    g, m, c = _case(name)
    out = _oracle_loss(c)
    sc = c['scene']
    raw = L.gt_flow_from_poses(sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'], sc['depth'], sc['internel_k'], 400.)
    check_digest(g, 'gt_flow', raw, 2e-3)
    check_digest(g, 'gt_flow_filtered', out['gt_flow'], 2e-3)
    check_digest(g, 'seq_pose', out['seq_pose'], 1e-5)
    check_digest(g, 'seq_flow', out['seq_flow'], 1e-6)
    check_digest(g, 'seq_mask', out['seq_mask'], 1e-6)
    for k in ('loss', 'loss_pose', 'loss_flow', 'loss_mask'):
        check_digest(g, k, out[k].reshape(1), 1e-4)


def test_loss_registry_builds_reference_config():
    import scflow_b200 as S
    cfg = dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='DisentanglePointMatchingLoss', symmetry_types={'cls_13': 1},
               mesh_diameter=[100.] * 21, mesh_path='/nonexistent/models_eval', loss_type='l1', disentangle_z=True, loss_weight=10.0))
    f = S.build_loss(cfg)
    assert isinstance(f.loss_func, S.DisentanglePointMatchingLoss) and f.gamma == 0.8
    with pytest.raises(RuntimeError):
        f.loss_func.packed(torch.device('cpu'))        # no model points yet
    with pytest.raises(NotImplementedError):
        S.build_loss(dict(type='DisentanglePointMatchingLoss', symmetry_types={}, mesh_diameter=[1.], loss_type='l2'))


def test_ply_reader(tmp_path):
    import numpy as np
    import scflow_b200.loss as SL
    v = np.random.RandomState(0).rand(17, 3).astype(np.float32)
    p = tmp_path / 'a.ply'
    with open(p, 'w') as f:
        f.write('ply\nformat ascii 1.0\nelement vertex 17\nproperty float x\nproperty float y\nproperty float z\nelement face 0\n'
                'property list uchar int vertex_indices\nend_header\n')
        for r in v:
            f.write(' '.join(repr(float(x)) for x in r) + '\n')
    assert torch.allclose(SL.read_ply_vertices(str(p)), torch.from_numpy(v))
    p2 = tmp_path / 'b.ply'
    with open(p2, 'wb') as f:
        f.write(b'ply\nformat binary_little_endian 1.0\nelement vertex 17\nproperty float x\nproperty float y\nproperty float z\n'
                b'property uchar red\nend_header\n')
        for r in v:
            f.write(r.tobytes() + b'\x07')
    assert torch.equal(SL.read_ply_vertices(str(p2)), torch.from_numpy(v))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_gpu_filter_flow_bit_exact(name):
    import scflow_b200 as S
    _, m, c = _case(name)
    sc = c['scene']
    raw = L.gt_flow_from_poses(sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'], sc['depth'], sc['internel_k'], 400.)
    ref = L.filter_flow_by_mask(raw.clone(), c['gt_mask'], 400.)
    got = S.filter_flow_by_mask(raw.clone().cuda(), c['gt_mask'].cuda(), 400.).cpu()
    assert torch.equal(got, ref), f'{int((got != ref).sum())} elements differ'      # byte/decision work: bit-exact
    assert int((ref >= 400.).sum()) > 0 and int((ref < 400.).sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize('name', CASES)
def test_gpu_refiner_loss_matches_oracle(name):
    import scflow_b200 as S
    _, m, c = _case(name)
    ref = _oracle_loss(c)
    sc = c['scene']
    sym = {f'cls_{k + 1}': 1 for k, s in enumerate(c['symmetric']) if s}
    pose_f = S.build_loss(dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(
        type='DisentanglePointMatchingLoss', symmetry_types=sym, mesh_diameter=c['diameters'], loss_type='l1', disentangle_z=True,
        loss_weight=10.)))
    pose_f.loss_func.set_meshes(c['meshes'])
    flow_f = S.build_loss(dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='RAFTLoss', loss_weight=.1, max_flow=400.)))
    mask_f = S.build_loss(dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='L1Loss', loss_weight=10.)))
    cu = lambda t: t.cuda()
    out, iters = S.refiner_loss([cu(t) for t in c['seq_flow']], [cu(t) for t in c['seq_mask']], [cu(t) for t in c['seq_rot']],
                                [cu(t) for t in c['seq_trs']], cu(ref['gt_flow']), cu(c['rendered_mask']), cu(c['gt_rot']),
                                cu(c['gt_trs']), cu(sc['label']), pose_f, flow_f, mask_f)
    out = out.cpu()
    assert iters == m['iters']
    # floating point: relative 2e-5 (fp32 reductions in a different order; stated tolerance)
    want = torch.cat([torch.stack([ref['loss'], ref['loss_pose'], ref['loss_flow'], ref['loss_mask']]), ref['seq_pose'], ref['seq_flow'],
                      ref['seq_mask']])
    rel = ((out - want).abs() / want.abs().clamp_min(1e-6)).max()
    print(f'{name}: loss {float(out[0]):.6f} vs oracle {float(want[0]):.6f}, max rel err {float(rel):.2e}')
    assert float(rel) < 2e-5


@pytest.mark.gpu
def test_gpu_refiner_loss_method_end_to_end():
    """SCFlowRefiner.loss (get_pose + GT flow + filter + fused losses) against the oracle composition on the same inputs."""
    import scflow_b200 as S
    seed, b, iters = 5, 2, 2
    c = L.make_loss_case(seed, b, iters)
    sc = c['scene']
    sym = {f'cls_{k + 1}': 1 for k, s in enumerate(c['symmetric']) if s}
    cfg = scflow_model_cfg(iters=iters, precision=1)
    cfg.update(pose_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(
                   type='DisentanglePointMatchingLoss', symmetry_types=sym, mesh_diameter=c['diameters'], loss_type='l1',
                   disentangle_z=True, loss_weight=10.)),
               flow_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='RAFTLoss', loss_weight=.1, max_flow=400.)),
               mask_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='L1Loss', loss_weight=10.)))
    model = S.build_refiner(cfg)
    sd = O.make_model_weights(seed)
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval()
    model.loss_functions()[0].loss_func.set_meshes(c['meshes'])
    data = dict(gt_rotations=c['gt_rot'], gt_translations=c['gt_trs'], ref_rotations=sc['ref_rotation'], ref_translations=sc['ref_translation'],
                real_images=sc['real_images'], rendered_images=sc['render_images'], rendered_depths=sc['depth'],
                rendered_masks=c['rendered_mask'], gt_masks=c['gt_mask'], internel_k=sc['internel_k'], labels=sc['label'])
    data = {k: v.cuda() for k, v in data.items()}
    with pytest.raises(RuntimeError, match='autograd is enabled'):
        model.loss(data)                                   # eval() + autograd: refused (train() gives the differentiable graph)
    with torch.no_grad():
        loss, log_vars, seq_rot, seq_trs = model.loss(data)
        outs = O.get_pose(sd, sc['render_images'], sc['real_images'], sc['ref_rotation'], sc['ref_translation'], sc['depth'],
                          sc['internel_k'], sc['label'], iters=iters)
    points_list = [c['meshes'][int(l)] for l in sc['label']]
    ref = L.refiner_loss(outs[1], outs[2], outs[3], outs[4], sc['ref_rotation'], sc['ref_translation'], c['gt_rot'], c['gt_trs'],
                         sc['depth'], sc['internel_k'], c['rendered_mask'], c['gt_mask'], sc['label'], points_list, c['symmetric'],
                         c['diameters'])
    print('loss', float(loss), 'oracle', float(ref['loss']), log_vars)
    assert abs(float(loss) - float(ref['loss'])) / float(ref['loss']) < 1e-4
    assert abs(log_vars['loss_flow'] - float(ref['loss_flow'])) / float(ref['loss_flow']) < 1e-4
    assert set(log_vars) >= {'loss', 'loss_pose', 'loss_flow', 'loss_mask', 'seq_0_pose_loss', 'seq_1_mask_loss'}
