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


# ---------------------------------------------------------------------------------------------------------------------
# the whole RAFT baseline decoder
# ---------------------------------------------------------------------------------------------------------------------
RAFT_CFG = dict(type='RAFTDecoder', net_type='Basic', num_levels=4, radius=4, iters=3, corr_lookup_cfg=dict(type='CorrLookup', align_corners=True),
                gru_type='SeqConv', act_cfg=dict(type='ReLU'))


def test_raft_decoder_oracle_matches_reference_golden():
    g = load_golden('raft_decoder_b2_16x16_it3')
    sd = RO.make_raft_decoder_weights(2)
    with torch.no_grad():
        preds = RO.raft_decoder_forward(sd, *RO.make_raft_inputs(2, 2, 16, 16), iters=3)
    assert len(preds) == 3
    for i, p in enumerate(preds):
        assert p.shape == (2, 2, 128, 128)
        assert float(np.abs(p.numpy() - g[f'upflow_{i}']).max()) < 1e-4


def test_raft_decoder_registers_with_reference_keys():
    import scflow_b200 as S
    from scflow_b200.builder import DECODERS, build_decoder
    assert DECODERS.get('RAFTDecoder') is S.RAFTDecoder
    dec = build_decoder(dict(RAFT_CFG))
    assert set(dec.state_dict().keys()) == set(RO.make_raft_decoder_weights(0).keys())
    assert dec.mask_channels == 576 and dec.iters == 3
    with pytest.raises(NotImplementedError):
        build_decoder(dict(RAFT_CFG, net_type='Small'))


@pytest.mark.gpu
def test_raft_decoder_matches_oracle():
    from scflow_b200.builder import build_decoder
    sd = RO.make_raft_decoder_weights(2)
    inputs = RO.make_raft_inputs(2, 2, 16, 16)
    with torch.no_grad():
        ref = RO.raft_decoder_forward(sd, *inputs, iters=3)
    dec = build_decoder(dict(RAFT_CFG))
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    with torch.no_grad():
        got = dec(*[t.cuda() for t in inputs])
    g = load_golden('raft_decoder_b2_16x16_it3')
    assert len(got) == 3
    for i, (a, b) in enumerate(zip(got, ref)):
        err = float((a.cpu() - b).abs().max())
        epe = float((a.cpu() - b).pow(2).sum(1).sqrt().mean())
        print(f'RAFTDecoder iteration {i}: max err {err:.3e}, EPE {epe:.3e}')
        assert err < 2e-3 and epe < 1e-4           # fp32 both sides; x8 flow scale; three recurrent iterations
        assert float(np.abs(a.cpu().numpy() - g[f'upflow_{i}']).max()) < 2e-3
    # without the mask head's convex combination the reference falls back to x8 bilinear up-sampling (raft_decoder.py:398-401)
    dec.convex_upsample_flow = False
    with torch.no_grad():
        plain = dec(*[t.cuda() for t in inputs])
    assert plain[0].shape == (2, 2, 128, 128) and not torch.equal(plain[0], got[0])


# ---------------------------------------------------------------------------------------------------------------------
# RAFTDecoderMask (flow + occlusion)
# ---------------------------------------------------------------------------------------------------------------------
RAFTM_CFG = dict(RAFT_CFG, type='RAFTDecoderMask', iters=2)


def test_raft_decoder_mask_oracle_matches_reference_golden():
    g = load_golden('raft_decoder_mask_b2_16x16_it2')
    sd = RO.make_raft_decoder_mask_weights(3)
    with torch.no_grad():
        flows, occs = RO.raft_decoder_mask_forward(sd, *RO.make_raft_inputs(3, 2, 16, 16), iters=2)
    for i in range(2):
        assert float(np.abs(flows[i].numpy() - g[f'upflow_{i}']).max()) < 1e-4
        assert float(np.abs(occs[i].numpy() - g[f'upocc_{i}']).max()) < 1e-5
        assert occs[i].shape == (2, 1, 128, 128) and float(occs[i].min()) >= 0. and float(occs[i].max()) <= 1.


def test_raft_decoder_mask_registers_with_reference_keys():
    from scflow_b200.builder import build_decoder
    dec = build_decoder(dict(RAFTM_CFG))
    assert set(dec.state_dict().keys()) == set(RO.make_raft_decoder_mask_weights(0).keys())


@pytest.mark.gpu
def test_raft_decoder_mask_matches_oracle():
    from scflow_b200.builder import build_decoder
    sd = RO.make_raft_decoder_mask_weights(3)
    inputs = RO.make_raft_inputs(3, 2, 16, 16)
    with torch.no_grad():
        ref_f, ref_o = RO.raft_decoder_mask_forward(sd, *inputs, iters=2)
    dec = build_decoder(dict(RAFTM_CFG))
    dec.load_state_dict(sd)
    dec = dec.cuda().eval()
    with torch.no_grad():
        got_f, got_o = dec(*[t.cuda() for t in inputs])
    for i in range(2):
        ef = float((got_f[i].cpu() - ref_f[i]).abs().max())
        eo = float((got_o[i].cpu() - ref_o[i]).abs().max())
        print(f'RAFTDecoderMask iteration {i}: flow max err {ef:.3e}, occlusion max err {eo:.3e}')
        assert ef < 2e-3 and eo < 2e-5
    # one-channel convex combination without the x8 factor, directly
    import scflow_b200 as S
    occ = torch.rand(2, 1, 5, 37)
    _, mask = RO.make_upsample_case(9, 2, 5, 37)
    err = float((S.ops.convex_upsample(occ.cuda(), mask.cuda(), mul=1.0).cpu() - RO.convex_upsample_mask(occ, mask)).abs().max())
    assert err < 2e-6
