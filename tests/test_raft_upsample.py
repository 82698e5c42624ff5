"""Convex x8 up-sampling of the RAFT baseline decoders: oracle vs the fixture generated from the
reference's own RAFTDecoder._upsample, and the CUDA kernel vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import raft_oracle as RO
from tests.util import load_golden


def test_upsample_oracle_matches_reference_golden():
    flow, mask = RO.make_upsample_case(4, 2, 6, 9)
    out = RO.convex_upsample(flow, mask)
    assert np.array_equal(out.numpy(), load_golden('convex_upsample_b2_6x9')['out'])
    # a one-hot mask on the centre tap reproduces 8 * nearest-neighbour up-sampling
    hot = torch.full((1, 576, 3, 4), -1e4)
    hot[:, 4 * 64:5 * 64] = 1e4
    f = torch.arange(24, dtype=torch.float32).view(1, 2, 3, 4)
    assert torch.equal(RO.convex_upsample(f, hot), 8. * f.repeat_interleave(8, 2).repeat_interleave(8, 3))


@pytest.mark.gpu
@pytest.mark.parametrize('b,h,w', [(2, 6, 9), (3, 32, 32), (1, 60, 80), (2, 5, 37)])
def test_convex_upsample_matches_oracle(b, h, w):
    import scflow_b200 as S
    flow, mask = RO.make_upsample_case(b * 100 + w, b, h, w)
    ref = RO.convex_upsample(flow, mask)
    got = S.ops.convex_upsample(flow.cuda(), mask.cuda()).cpu()
    assert got.shape == ref.shape
    # fp32 both sides; the reference's softmax / 9-term sum run in a different order: tolerance a few ulp of the x8 flow scale
    err = float((got - ref).abs().max())
    assert err < 2e-5 * max(1.0, float(ref.abs().max())), f'max err {err:.3e}'


@pytest.mark.gpu
def test_convex_upsample_rejects_bad_shapes_and_cpu_tensors():
    import scflow_b200 as S
    flow, mask = RO.make_upsample_case(1, 1, 4, 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        S.ops.convex_upsample(flow, mask)
    with pytest.raises(ValueError, match='576'):
        S.ops.convex_upsample(flow.cuda(), mask[:, :64].contiguous().cuda())
