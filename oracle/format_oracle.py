"""CPU oracle of the input formatting next to the path: BaseRefiner.format_data_test
(models/refiner/base_refiner.py:79-133 of the reference), restated with plain torch ops.

TEST INFRASTRUCTURE ONLY (same rule as scflow_oracle.py).  Parity pin: checked against the reference's own
``BaseRefiner.format_data_test`` run through oracle/ref_shim.py with a deterministic stand-in renderer
(oracle/make_golden_format.py; fixture tests/golden/format_test_b5.npz).  All paths relative to /root/reference.
"""
import types
from typing import Dict

import torch


def fake_renderer(height: int = 48, width: int = 64, faces_per_pixel: int = 2):
    """Deterministic stand-in for models/utils/renderer.py:Renderer with the same call interface and output structure:
    images [N,H,W,4] in [0,1] (RGBA), fragments.zbuf [N,H,W,K] (> 0 on the object, -1 on the background as pytorch3d does).
    The content is a function of the poses / intrinsics / labels so that argument order mistakes show up."""
    def render(rotations, translations, internel_k, labels):
        n = rotations.shape[0]
        dev = rotations.device
        ys, xs = torch.meshgrid(torch.arange(height, device=dev, dtype=torch.float32),
                                torch.arange(width, device=dev, dtype=torch.float32), indexing='ij')
        cx = internel_k[:, 0, 2].view(n, 1, 1) * (width / 640.)
        cy = internel_k[:, 1, 2].view(n, 1, 1) * (height / 480.)
        rad = (6. + labels.to(torch.float32).view(n, 1, 1)) * (1. + rotations[:, 0, 0].abs().view(n, 1, 1))
        inside = ((xs[None] - cx) ** 2 + (ys[None] - cy) ** 2) < rad ** 2
        z = translations[:, 2].view(n, 1, 1) + 0.5 * xs[None] - 0.25 * ys[None]
        zbuf = torch.where(inside, z, torch.full_like(z, -1.))
        zbuf = torch.stack([zbuf] + [torch.where(inside, z + 10. * (k + 1), torch.full_like(z, -1.)) for k in range(faces_per_pixel - 1)], -1)
        rgb = torch.stack([torch.sin(0.11 * xs[None] + rotations[:, 0, 1].view(n, 1, 1)) * 0.5 + 0.5,
                           torch.cos(0.07 * ys[None] + rotations[:, 1, 2].view(n, 1, 1)) * 0.5 + 0.5,
                           ((xs[None] + ys[None]) % 17.) / 17. + 0. * cx], -1)
        images = torch.cat([rgb * inside[..., None], inside[..., None].to(torch.float32)], -1)
        return dict(images=images, fragments=types.SimpleNamespace(zbuf=zbuf))
    return render


def make_data_batch(seed: int, patch_nums=(2, 3), height: int = 48, width: int = 64, with_gt: bool = True) -> Dict:
    """A test-time data_batch in the reference's collated layout (per-image lists of per-patch tensors)."""
    g = torch.Generator().manual_seed(seed)
    img, rots, trs, labels, ks, tms, ori_k, gt_r, gt_t, masks, depths = [], [], [], [], [], [], [], [], [], [], []
    for n in patch_nums:
        img.append(torch.randn(n, 3, height, width, generator=g))
        q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
        rots.append(q)
        trs.append(torch.cat([30. * torch.randn(n, 2, generator=g), 800. + 100. * torch.rand(n, 1, generator=g)], 1))
        labels.append(torch.randint(0, 21, (n,), generator=g))
        k = torch.tensor([[572.4, 0., 325.3], [0., 573.6, 242.0], [0., 0., 1.]]).repeat(n, 1, 1)
        k[:, 0, 2] += 20. * torch.randn(n, generator=g)
        k[:, 1, 2] += 20. * torch.randn(n, generator=g)
        ks.append(k)
        tms.append(torch.eye(3).repeat(n, 1, 1) + 0.01 * torch.randn(n, 3, 3, generator=g))
        ori_k.append(torch.tensor([[1066.8, 0., 312.9], [0., 1067.5, 241.3], [0., 0., 1.]]) + torch.randn(3, 3, generator=g))
        gt_r.append(q.transpose(1, 2).contiguous())
        gt_t.append(trs[-1] + torch.randn(n, 3, generator=g))
        masks.append(torch.rand(n, height, width, generator=g) > 0.5)
        depths.append(torch.rand(n, height, width, generator=g) * 1000.)
    annots = dict(ref_rotations=rots, ref_translations=trs, labels=labels, k=ks, ori_k=ori_k, transform_matrix=tms)
    if with_gt:
        annots.update(gt_rotations=gt_r, gt_translations=gt_t, gt_masks=masks, depths=depths)
    metas = [dict(img_norm_cfg=dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)) for _ in patch_nums]
    return dict(img=img, annots=annots, img_metas=metas)


def format_data_test(data_batch: Dict, renderer) -> Dict:
    """base_refiner.py:79-133."""
    real_images, annots, meta_infos = data_batch['img'], data_batch['annots'], data_batch['img_metas']
    per_img_patch_num = [len(images) for images in real_images]
    ref_rotations, ref_translations = torch.cat(annots['ref_rotations'], 0), torch.cat(annots['ref_translations'], 0)
    labels, internel_k = torch.cat(annots['labels']), torch.cat(annots['k'])
    out = dict(real_images=torch.cat(real_images), labels=labels, internel_k=internel_k, ref_rotations=ref_rotations,
               ref_translations=ref_translations, transform_matrix=torch.cat(annots['transform_matrix']),
               ori_k=torch.cat([k[None].expand(n, 3, 3) for k, n in zip(annots['ori_k'], per_img_patch_num)]),
               per_img_patch_num=per_img_patch_num, meta_infos=meta_infos)
    ro = renderer(ref_rotations, ref_translations, internel_k, labels)                 # :91
    images = ro['images'][..., :3].permute(0, 3, 1, 2).contiguous()                    # :93
    depths = ro['fragments'].zbuf[..., 0]                                              # :94-95
    cfg = meta_infos[0]['img_norm_cfg']
    mean = torch.Tensor(cfg['mean']).view(1, 3, 1, 1) / 255.                           # :100-102
    std = torch.Tensor(cfg['std']).view(1, 3, 1, 1) / 255.
    out.update(rendered_images=(images - mean) / std, rendered_depths=depths, rendered_masks=(depths > 0).to(torch.float32))
    if 'depths' in annots:
        out.update(real_depths=torch.cat(annots['depths'], 0))
    if 'gt_rotations' in annots:
        out.update(gt_rotations=torch.cat(annots['gt_rotations'], 0), gt_translations=torch.cat(annots['gt_translations'], 0))
    if 'gt_masks' in annots:
        out.update(gt_masks=torch.cat([m.to_tensor(dtype=torch.bool, device=labels.device) if hasattr(m, 'to_tensor') else m
                                       for m in annots['gt_masks']], 0))
    return out


TENSOR_KEYS = ('real_images', 'rendered_images', 'labels', 'ori_k', 'transform_matrix', 'internel_k', 'ref_rotations',
               'ref_translations', 'rendered_masks', 'rendered_depths', 'real_depths', 'gt_rotations', 'gt_translations', 'gt_masks')


# ----------------------------------------------------------------------------------------------------------------------
# BaseRefiner.format_data_train_sup (base_refiner.py:136-191), pinned like format_data_test (fixture format_train_b5.npz)
# ----------------------------------------------------------------------------------------------------------------------
def make_train_batch(seed: int, patch_nums=(2, 3), height: int = 48, width: int = 64) -> Dict:
    batch = make_data_batch(seed, patch_nums, height, width, with_gt=True)
    g = torch.Generator().manual_seed(seed + 1)
    n = sum(patch_nums)
    batch['annots'].update(init_add_error=30. * torch.rand(n, generator=g), init_rot_error=20. * torch.rand(n, generator=g),
                           init_trans_error=40. * torch.rand(n, generator=g))
    return batch


def format_data_train_sup(data_batch: Dict, renderer) -> Dict:
    """base_refiner.py:136-191 with render_augmentation = None (every shipped config)."""
    real_images, annots, meta_infos = data_batch['img'], data_batch['annots'], data_batch['img_metas']
    out = {}
    for name in ('add', 'rot', 'trans'):
        std, mean = torch.std_mean(annots[f'init_{name}_error'], unbiased=False)
        out[f'init_{name}_error_mean'], out[f'init_{name}_error_std'] = mean, std
    ref_rotations, ref_translations = torch.cat(annots['ref_rotations'], 0), torch.cat(annots['ref_translations'], 0)
    labels, internel_k = torch.cat(annots['labels']), torch.cat(annots['k'])
    out.update(real_images=torch.cat(real_images), ref_rotations=ref_rotations, ref_translations=ref_translations,
               gt_rotations=torch.cat(annots['gt_rotations'], 0), gt_translations=torch.cat(annots['gt_translations'], 0),
               labels=labels, internel_k=internel_k)
    ro = renderer(ref_rotations, ref_translations, internel_k, labels)
    images = ro['images'][..., :3].permute(0, 3, 1, 2).contiguous()
    depths = ro['fragments'].zbuf[..., 0]
    cfg = meta_infos[0]['img_norm_cfg']
    mean = torch.Tensor(cfg['mean']).view(1, 3, 1, 1) / 255.
    std = torch.Tensor(cfg['std']).view(1, 3, 1, 1) / 255.
    out.update(rendered_images=(images - mean) / std, rendered_depths=depths, rendered_masks=(depths > 0).to(torch.float32))
    if 'gt_masks' in annots:
        out['gt_masks'] = torch.cat([m.to_tensor(dtype=torch.bool, device=labels.device) if hasattr(m, 'to_tensor') else m
                                     for m in annots['gt_masks']], 0)
    return out


TRAIN_KEYS = ('ref_rotations', 'ref_translations', 'gt_rotations', 'gt_translations', 'labels', 'internel_k', 'rendered_images',
              'real_images', 'rendered_masks', 'rendered_depths', 'init_add_error_mean', 'init_add_error_std', 'init_rot_error_mean',
              'init_rot_error_std', 'init_trans_error_mean', 'init_trans_error_std', 'gt_masks')
