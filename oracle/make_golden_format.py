"""Runs the reference's own BaseRefiner.format_data_test (through oracle/ref_shim.py, with the deterministic stand-in
renderer of oracle/format_oracle.py) on a seeded batch, checks the restatement in oracle/format_oracle.py against it
bit for bit, and writes tests/golden/format_test_b5.npz.   Usage (build container only): python -m oracle.make_golden_format"""
import os
import types

import numpy as np
import torch

from . import format_oracle as FO
from . import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Mask:
    """mmdet-style mask container: the reference calls .to_tensor(dtype=, device=) on each entry (base_refiner.py:129)."""
    def __init__(self, t):
        self.t = t

    def to_tensor(self, dtype, device):
        return self.t.to(dtype=dtype, device=device)


def main():
    ref_shim.install()
    from models.refiner import base_refiner
    batch = FO.make_data_batch(5)
    ref_batch = dict(batch, annots=dict(batch['annots'], gt_masks=[_Mask(m) for m in batch['annots']['gt_masks']]))
    fake_self = types.SimpleNamespace(renderer=FO.fake_renderer())
    ref = base_refiner.BaseRefiner.format_data_test(fake_self, ref_batch)
    mine = FO.format_data_test(batch, FO.fake_renderer())
    assert set(ref.keys()) == set(mine.keys()), set(ref.keys()) ^ set(mine.keys())
    out = {}
    for k in FO.TENSOR_KEYS:
        assert ref[k].shape == mine[k].shape and ref[k].dtype == mine[k].dtype, k
        assert torch.equal(ref[k], mine[k]), f'{k}: restatement differs from the reference'
        out[k] = ref[k].numpy()
    assert ref['per_img_patch_num'] == mine['per_img_patch_num']
    out['per_img_patch_num'] = np.asarray(ref['per_img_patch_num'])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'format_test_b5.npz'), **out)
    print('restatement == reference on', len(FO.TENSOR_KEYS), 'tensors; fixture written')
    # ---- format_data_train_sup
    tb = FO.make_train_batch(5)
    ref_tb = dict(tb, annots=dict(tb['annots'], gt_masks=[_Mask(m) for m in tb['annots']['gt_masks']]))
    fake_self = types.SimpleNamespace(renderer=FO.fake_renderer(), render_augmentation=None)
    ref = base_refiner.BaseRefiner.format_data_train_sup(fake_self, ref_tb)
    mine = FO.format_data_train_sup(tb, FO.fake_renderer())
    assert set(ref.keys()) == set(mine.keys()) == set(FO.TRAIN_KEYS), set(ref.keys()) ^ set(mine.keys())
    out = {}
    for k in FO.TRAIN_KEYS:
        assert torch.equal(ref[k], mine[k]), f'{k}: restatement differs from the reference'
        out[k] = ref[k].numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'format_train_b5.npz'), **out)
    print('format_data_train_sup: restatement == reference on', len(FO.TRAIN_KEYS), 'tensors; fixture written')


if __name__ == '__main__':
    main()
