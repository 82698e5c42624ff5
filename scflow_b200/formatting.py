"""Input formatting around the refinement path: the host-side mirror of
``BaseRefiner.format_data_test`` (models/refiner/base_refiner.py:79-133 of the reference).

The renderer itself (pytorch3d, models/utils/renderer.py) is outside this package: any callable with the reference's
interface ``renderer(rotations, translations, internel_k, labels) -> {'images': [N,H,W,4], 'fragments': obj with .zbuf
[N,H,W,K]}`` can be plugged in.  Everything after the renderer call - alpha drop, NHWC -> NCHW, normalisation, depth and
silhouette mask - is one CUDA pass (``ops.format_rendered``); the rest is list concatenation (plumbing).
"""
from typing import Callable, Dict

import torch

from . import ops


def norm_constants(img_norm_cfg: Dict):
    """The fp32 mean / std the reference subtracts / divides by (base_refiner.py:103-106): Tensor(values) / 255."""
    mean = (torch.tensor(img_norm_cfg['mean'], dtype=torch.float32) / 255.).tolist()
    std = (torch.tensor(img_norm_cfg['std'], dtype=torch.float32) / 255.).tolist()
    return mean, std


def format_data_test(data_batch: Dict, renderer: Callable) -> Dict:
    """base_refiner.py:79-133: per-image patch lists -> flat batch, render the reference poses, format the render."""
    if renderer is None:
        raise RuntimeError('format_data_test needs a renderer (the reference builds models/utils/renderer.py:Renderer from '
                           'the config; plug any callable with the same interface in via SCFlowRefiner.set_renderer)')
    real_images, annots, meta_infos = data_batch['img'], data_batch['annots'], data_batch['img_metas']
    per_img_patch_num = [len(images) for images in real_images]
    real_images = torch.cat(list(real_images))
    ref_rotations = torch.cat(list(annots['ref_rotations']), dim=0)
    ref_translations = torch.cat(list(annots['ref_translations']), dim=0)
    labels = torch.cat(list(annots['labels']))
    internel_k = torch.cat(list(annots['k']))
    transform_matrixs = torch.cat(list(annots['transform_matrix']))
    ori_k = torch.cat([k[None].expand(n, 3, 3) for k, n in zip(annots['ori_k'], per_img_patch_num)])

    render_outputs = renderer(ref_rotations, ref_translations, internel_k, labels)
    mean, std = norm_constants(meta_infos[0]['img_norm_cfg'])
    rendered_images, rendered_depths, rendered_masks = ops.format_rendered(
        render_outputs['images'].float().contiguous(), render_outputs['fragments'].zbuf.float().contiguous(), mean, std)
    output = dict(real_images=real_images, rendered_images=rendered_images, labels=labels, ori_k=ori_k,
                  transform_matrix=transform_matrixs, internel_k=internel_k, ref_rotations=ref_rotations,
                  ref_translations=ref_translations, rendered_masks=rendered_masks, rendered_depths=rendered_depths,
                  per_img_patch_num=per_img_patch_num, meta_infos=meta_infos)
    if 'depths' in annots:
        output.update(real_depths=torch.cat(list(annots['depths']), dim=0))
    if 'gt_rotations' in annots:
        output.update(gt_rotations=torch.cat(list(annots['gt_rotations']), dim=0),
                      gt_translations=torch.cat(list(annots['gt_translations']), dim=0))
    if 'gt_masks' in annots:
        dev = ref_rotations.device
        masks = [m.to_tensor(dtype=torch.bool, device=dev) if hasattr(m, 'to_tensor') else m.to(device=dev, dtype=torch.bool)
                 for m in annots['gt_masks']]
        output.update(gt_masks=torch.cat(masks, dim=0))
    return output


def format_data_train_sup(data_batch: Dict, renderer: Callable, render_augmentation=None) -> Dict:
    """base_refiner.py:136-191: the supervised-training counterpart of ``format_data_test`` (adds the ground-truth poses and
    the statistics of the initial pose error; no ``ori_k`` / ``transform_matrix``).  ``render_augmentation`` (kornia in the
    reference, ``None`` in every shipped config) is not part of this package."""
    if renderer is None:
        raise RuntimeError('format_data_train_sup needs a renderer (see format_data_test)')
    if render_augmentation is not None:
        raise NotImplementedError('render augmentations (kornia AugmentationSequential, base_refiner.py:50-62) are outside this '
                                  'package; the shipped configs use none')
    real_images, annots, meta_infos = data_batch['img'], data_batch['annots'], data_batch['img_metas']
    stats = {}
    for name in ('add', 'rot', 'trans'):
        std, mean = torch.std_mean(annots[f'init_{name}_error'], unbiased=False)          # :142-144
        stats[f'init_{name}_error_mean'], stats[f'init_{name}_error_std'] = mean, std
    real_images = torch.cat(list(real_images))
    ref_rotations = torch.cat(list(annots['ref_rotations']), dim=0)
    ref_translations = torch.cat(list(annots['ref_translations']), dim=0)
    gt_rotations = torch.cat(list(annots['gt_rotations']), dim=0)
    gt_translations = torch.cat(list(annots['gt_translations']), dim=0)
    labels, internel_k = torch.cat(list(annots['labels'])), torch.cat(list(annots['k']))
    render_outputs = renderer(ref_rotations, ref_translations, internel_k, labels)
    mean, std = norm_constants(meta_infos[0]['img_norm_cfg'])
    rendered_images, rendered_depths, rendered_masks = ops.format_rendered(
        render_outputs['images'].float().contiguous(), render_outputs['fragments'].zbuf.float().contiguous(), mean, std)
    output = dict(ref_rotations=ref_rotations, ref_translations=ref_translations, gt_rotations=gt_rotations,
                  gt_translations=gt_translations, labels=labels, internel_k=internel_k, rendered_images=rendered_images,
                  real_images=real_images, rendered_masks=rendered_masks, rendered_depths=rendered_depths, **stats)
    if 'gt_masks' in annots:
        dev = gt_rotations.device
        masks = [m.to_tensor(dtype=torch.bool, device=dev) if hasattr(m, 'to_tensor') else m.to(device=dev, dtype=torch.bool)
                 for m in annots['gt_masks']]
        output['gt_masks'] = torch.cat(masks, dim=0)
    return output
