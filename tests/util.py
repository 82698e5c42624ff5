"""Shared helpers for the test-suite: golden digests, module construction from oracle weights."""
import os
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
MAX_PTS = 4096


def digest_indices(name: str, numel: int) -> np.ndarray:
    """Must match oracle/make_golden.py."""
    if numel <= MAX_PTS:
        return np.arange(numel, dtype=np.int64)
    rng = np.random.RandomState(zlib.crc32(name.encode()) & 0x7fffffff)
    return np.sort(rng.choice(numel, MAX_PTS, replace=False)).astype(np.int64)


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def assert_matches_digest(g, name: str, t: torch.Tensor, atol: float, rtol: float = 0.0, sum_rtol: float = None):
    """Compare tensor ``t`` against the stored digest ``name`` of golden file ``g``."""
    shape = tuple(int(s) for s in g[name + '/shape'])
    assert tuple(t.shape) == shape, f'{name}: shape {tuple(t.shape)} != golden {shape}'
    a = t.detach().cpu().double().numpy().reshape(-1)
    idx = digest_indices(name, a.size)
    ref = g[name + '/vals'].astype(np.float64)
    err = np.abs(a[idx] - ref)
    tol = atol + rtol * np.abs(ref)
    assert np.all(err <= tol), f'{name}: max err {err.max():.3e} (tol {atol:.1e}+{rtol:.1e}*|ref|), worst at {int(err.argmax())}'
    if sum_rtol is not None:
        asum = float(g[name + '/asum'])
        assert abs(np.abs(a).sum() - asum) <= sum_rtol * max(asum, 1e-30), f'{name}: abs-sum mismatch'
    return float(err.max())


def scflow_model_cfg(iters: int = 8, num_class: int = 21, precision: int = 0, use_cuda_graph: bool = False) -> dict:
    """The `model` dict of configs/refine_models/scflow.py:16-113 (reference), minus losses/renderer/init_cfg."""
    enc_init = [dict(type='Kaiming', layer=['Conv2d'], mode='fan_out', nonlinearity='relu'),
                dict(type='Constant', layer=['InstanceNorm2d'], val=1, bias=0)]
    return dict(
        type='SCFlowRefiner', cxt_channels=128, h_channels=128, seperate_encoder=False, max_flow=400.,
        filter_invalid_flow=True,
        encoder=dict(type='RAFTEncoder', in_channels=3, out_channels=256, net_type='Basic', norm_cfg=dict(type='IN'), init_cfg=enc_init),
        cxt_encoder=dict(type='RAFTEncoder', in_channels=3, out_channels=256, net_type='Basic', norm_cfg=dict(type='BN'), init_cfg=enc_init),
        decoder=dict(
            type='SCFlowDecoder', net_type='Basic', num_levels=4, radius=4, iters=iters, detach_flow=True, detach_mask=True,
            detach_pose=True, detach_depth_for_xy=True, mask_flow=False, mask_corr=False,
            pose_head_cfg=dict(type='MultiClassPoseHead', num_class=num_class, in_channels=224, net_type='Basic',
                               rotation_mode='ortho6d', norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                               act_cfg=dict(type='ReLU')),
            corr_lookup_cfg=dict(align_corners=True), gru_type='SeqConv', act_cfg=dict(type='ReLU'),
            precision=precision, use_cuda_graph=use_cuda_graph),
        flow_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='RAFTLoss', loss_weight=.1, max_flow=400.)),
        pose_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='DisentanglePointMatchingLoss', loss_weight=10.0)),
        mask_loss_cfg=dict(type='SequenceLoss', gamma=0.8, loss_func_cfg=dict(type='L1Loss', loss_weight=10.)),
        freeze_bn=False, freeze_encoder=False, train_cfg=dict(), test_cfg=dict(iters=iters),
        init_cfg=dict(type='Pretrained', checkpoint='work_dirs/raft_8x2_100k_flyingthings3d_400x720_convertered.pth'))


def build_decoder_from_oracle_weights(seed: int, iters: int, device='cuda', precision: int = 0, use_cuda_graph: bool = False):
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    cfg = scflow_model_cfg(iters, precision=precision, use_cuda_graph=use_cuda_graph)['decoder']
    dec = S.build_decoder(cfg)
    sd = O.make_decoder_weights(seed)
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return dec.to(device).eval(), sd


def to_dev(d: dict, device='cuda'):
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
