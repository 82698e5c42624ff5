"""Input formatting next to the path: oracle vs the fixture generated from the reference's own
BaseRefiner.format_data_test (oracle/make_golden_format.py), and the CUDA path vs the oracle (bit-exact)."""
import types

import numpy as np
import pytest
import torch

from oracle import format_oracle as FO
from tests.util import load_golden


def test_format_oracle_matches_reference_golden():
    g = load_golden('format_test_b5')
    out = FO.format_data_test(FO.make_data_batch(5), FO.fake_renderer())
    for k in FO.TENSOR_KEYS:
        assert np.array_equal(out[k].numpy(), g[k]), k
    assert out['per_img_patch_num'] == g['per_img_patch_num'].tolist()
    # structure of the formatted render: NCHW image, silhouette mask = depth > 0
    assert out['rendered_images'].shape == (5, 3, 48, 64) and out['rendered_depths'].shape == (5, 48, 64)
    assert torch.equal(out['rendered_masks'], (out['rendered_depths'] > 0).float())


def test_format_without_gt_and_ragged_patch_counts():
    out = FO.format_data_test(FO.make_data_batch(7, patch_nums=(1, 4, 2), with_gt=False), FO.fake_renderer())
    assert 'gt_rotations' not in out and 'gt_masks' not in out and 'real_depths' not in out
    assert out['per_img_patch_num'] == [1, 4, 2] and out['ori_k'].shape == (7, 3, 3)


def test_product_formatting_has_no_cpu_path():
    from scflow_b200 import formatting
    with pytest.raises(RuntimeError, match='renderer'):
        formatting.format_data_test(FO.make_data_batch(5), None)
    with pytest.raises(RuntimeError, match='CUDA'):
        formatting.format_data_test(FO.make_data_batch(5), FO.fake_renderer())      # CPU tensors: must refuse, not fall back
    mean, std = formatting.norm_constants(dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375]))
    assert mean == (torch.tensor([123.675, 116.28, 103.53]) / 255.).tolist() and len(std) == 3


def _to(obj, dev):
    if isinstance(obj, torch.Tensor):
        return obj.to(dev)
    if isinstance(obj, (list, tuple)):
        return [_to(o, dev) for o in obj]
    if isinstance(obj, dict):
        return {k: _to(v, dev) for k, v in obj.items()}
    return obj


def _cuda_renderer(**kw):
    cpu = FO.fake_renderer(**kw)

    def render(r, t, k, l):
        o = cpu(r.cpu(), t.cpu(), k.cpu(), l.cpu())
        return dict(images=o['images'].cuda(), fragments=types.SimpleNamespace(zbuf=o['fragments'].zbuf.cuda()))
    return render


def _shared_renderers(**kw):
    """(cpu_renderer, cuda_renderer) that hand out THE SAME pixels: the stand-in renderer is evaluated once (first call, on
    the CPU) and its output tensors are replayed to the other side, so the comparison is hermetic - it never depends on two
    evaluations of the host's sin / cos giving the same bits (they need not: vectorised vs scalar tails, thread splits)."""
    cpu = FO.fake_renderer(**kw)
    cache = {}

    def _render(r, t, k, l):
        if 'o' not in cache:
            cache['args'] = [a.detach().cpu().clone() for a in (r, t, k, l)]
            cache['o'] = cpu(*cache['args'])
        else:                       # both sides must ask for the same render
            for a, b in zip(cache['args'], (r, t, k, l)):
                assert torch.equal(a, b.detach().cpu()), 'renderer called with different arguments by the two sides'
        return cache['o']

    def cpu_renderer(r, t, k, l):
        o = _render(r, t, k, l)
        return dict(images=o['images'].clone(), fragments=types.SimpleNamespace(zbuf=o['fragments'].zbuf.clone()))

    def cuda_renderer(r, t, k, l):
        assert r.is_cuda and t.is_cuda and k.is_cuda and l.is_cuda
        o = _render(r, t, k, l)
        return dict(images=o['images'].cuda(), fragments=types.SimpleNamespace(zbuf=o['fragments'].zbuf.cuda()))
    return cpu_renderer, cuda_renderer


def _assert_bit_equal(got: torch.Tensor, ref: torch.Tensor, name: str):
    got = got.cpu()
    if torch.equal(got, ref):
        return
    if got.dtype == torch.float32:
        gi, ri = got.contiguous().view(torch.int32).to(torch.int64), ref.contiguous().view(torch.int32).to(torch.int64)
        bad = gi != ri
        raise AssertionError(f'{name}: {int(bad.sum())} of {bad.numel()} elements differ, max |ulp| {int((gi - ri).abs().max())}, '
                             f'max |diff| {float((got - ref).abs().max()):.3e}')
    raise AssertionError(f'{name}: {int((got != ref).sum())} of {got.numel()} elements differ')


@pytest.mark.gpu
@pytest.mark.parametrize('seed,patch_nums,faces', [(5, (2, 3), 2), (11, (1, 4, 2), 1), (3, (8,), 3)])
def test_format_data_test_matches_oracle_bit_exact(seed, patch_nums, faces):
    from scflow_b200 import formatting
    batch = FO.make_data_batch(seed, patch_nums=patch_nums)
    cpu_renderer, cuda_renderer = _shared_renderers(faces_per_pixel=faces)        # ONE render, identical pixels for both sides
    ref = FO.format_data_test(batch, cpu_renderer)
    got = formatting.format_data_test(dict(img=_to(batch['img'], 'cuda'), annots=_to(batch['annots'], 'cuda'), img_metas=batch['img_metas']),
                                      cuda_renderer)
    assert set(got.keys()) == set(ref.keys())
    for k in FO.TENSOR_KEYS:
        assert got[k].is_cuda and got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, k
        _assert_bit_equal(got[k], ref[k], k)
    assert got['per_img_patch_num'] == ref['per_img_patch_num']


@pytest.mark.gpu
def test_format_rendered_rgb_without_alpha_and_odd_sizes():
    import scflow_b200 as S
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 37, 53, 3, generator=g)
    zb = torch.rand(3, 37, 53, 1, generator=g) - 0.3
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
    out, depth, mask = S.ops.format_rendered(img.cuda(), zb.cuda(), mean, std)
    m = torch.tensor(mean).view(1, 3, 1, 1)
    s = torch.tensor(std).view(1, 3, 1, 1)
    assert torch.equal(out.cpu(), (img.permute(0, 3, 1, 2) - m) / s)
    assert torch.equal(depth.cpu(), zb[..., 0]) and torch.equal(mask.cpu(), (zb[..., 0] > 0).float())


@pytest.mark.gpu
def test_refiner_forward_formats_and_refines_a_dataset_batch():
    """SCFlowRefiner.forward(data_batch) = format_data_test -> forward_single_pass (base_refiner.py:338-343)."""
    import scflow_b200 as S
    from oracle import scflow_oracle as O
    from tests.util import scflow_model_cfg
    model = S.build_refiner(scflow_model_cfg(iters=2, precision=1))
    model.load_state_dict(O.make_model_weights(0), strict=False)
    model = model.cuda().eval()
    batch = FO.make_data_batch(9, patch_nums=(2, 1), height=256, width=256, with_gt=False)
    data_batch = dict(img=_to(batch['img'], 'cuda'), annots=_to(batch['annots'], 'cuda'), img_metas=batch['img_metas'])
    with pytest.raises(NotImplementedError, match='renderer'):
        model(data_batch)
    model.set_renderer(_cuda_renderer(height=256, width=256))
    with torch.no_grad():
        out = model(data_batch)
    assert [r.shape for r in out['rotations']] == [(2, 3, 3), (1, 3, 3)]
    assert all(torch.isfinite(t).all() for t in out['translations'])
    # 'adapt_intrinsic' (shipped pipelines): poses are returned as predicted; the cv2.solvePnP modes are refused loudly
    data_batch['img_metas'] = [dict(m, geometry_transform_mode='adapt_intrinsic') for m in batch['img_metas']]
    with torch.no_grad():
        again = model(data_batch)
    assert all(torch.equal(a, b) for a, b in zip(again['rotations'], out['rotations']))
    data_batch['img_metas'] = [dict(m, geometry_transform_mode='target_intrinsic') for m in batch['img_metas']]
    with pytest.raises(NotImplementedError, match='PnP'), torch.no_grad():
        model(data_batch)


# ---------------------------------------------------------------------------------------------------------------------
# format_data_train_sup
# ---------------------------------------------------------------------------------------------------------------------
def test_format_train_oracle_matches_reference_golden():
    g = load_golden('format_train_b5')
    out = FO.format_data_train_sup(FO.make_train_batch(5), FO.fake_renderer())
    assert set(out.keys()) == set(FO.TRAIN_KEYS)
    for k in FO.TRAIN_KEYS:
        assert np.array_equal(out[k].numpy(), g[k]), k


def test_product_train_formatting_refuses_augmentations_and_cpu():
    from scflow_b200 import formatting
    with pytest.raises(NotImplementedError, match='augment'):
        formatting.format_data_train_sup(FO.make_train_batch(5), FO.fake_renderer(), render_augmentation=object())
    with pytest.raises(RuntimeError, match='CUDA'):
        formatting.format_data_train_sup(FO.make_train_batch(5), FO.fake_renderer())


@pytest.mark.gpu
def test_format_data_train_sup_matches_oracle_bit_exact():
    from scflow_b200 import formatting
    batch = FO.make_train_batch(6, patch_nums=(3, 1, 2))
    cpu_renderer, cuda_renderer = _shared_renderers()
    ref = FO.format_data_train_sup(batch, cpu_renderer)
    got = formatting.format_data_train_sup(dict(img=_to(batch['img'], 'cuda'), annots=_to(batch['annots'], 'cuda'),
                                                img_metas=batch['img_metas']), cuda_renderer)
    assert set(got.keys()) == set(ref.keys())
    for k in FO.TRAIN_KEYS:
        assert got[k].is_cuda and got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype, k
        if k.startswith('init_'):          # torch.std_mean on the GPU vs the CPU: reduction order
            assert float((got[k].cpu() - ref[k]).abs()) < 1e-4 * max(1.0, float(ref[k].abs())), k
        else:
            _assert_bit_equal(got[k], ref[k], k)
