"""Runs the reference's own RAFTDecoder._upsample (through oracle/ref_shim.py) on a seeded case, checks the restatement in
oracle/raft_oracle.py against it bit for bit and writes tests/golden/convex_upsample_b2_6x9.npz.
Usage (build container only): python -m oracle.make_golden_raft"""
import os
import types

import numpy as np
import torch

from . import raft_oracle as RO
from . import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ref_shim.install()
    from models.decoder.raft_decoder import RAFTDecoder
    flow, mask = RO.make_upsample_case(4, 2, 6, 9)
    fake_self = types.SimpleNamespace(num_levels=4, radius=4)
    ref = RAFTDecoder._upsample(fake_self, flow, mask)
    mine = RO.convex_upsample(flow, mask)
    assert torch.equal(ref, mine), 'restatement differs from the reference'
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'convex_upsample_b2_6x9.npz'), out=ref.numpy())
    print('restatement == reference; fixture written', tuple(ref.shape))
    # ---- the whole RAFTDecoder
    dec = RAFTDecoder(net_type='Basic', num_levels=4, radius=4, iters=3, corr_lookup_cfg=dict(align_corners=True),
                      gru_type='SeqConv', act_cfg=dict(type='ReLU'))
    sd = RO.make_raft_decoder_weights(2)
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    dec.eval()
    inputs = RO.make_raft_inputs(2, 2, 16, 16)
    with torch.no_grad():
        ref_preds = dec(*inputs)
        my_preds = RO.raft_decoder_forward(sd, *inputs, iters=3)
    out = {}
    for i, (a, b) in enumerate(zip(ref_preds, my_preds)):
        err = float((a - b).abs().max())
        assert err < 1e-4, f'iteration {i}: restatement differs from the reference by {err:.3e}'
        out[f'upflow_{i}'] = a.numpy()
        print(f'iteration {i}: |restatement - reference| max {err:.2e}, |flow| max {float(a.abs().max()):.2f}')
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'raft_decoder_b2_16x16_it3.npz'), **out)
    # ---- RAFTDecoderMask
    from models.decoder.raft_decoder_mask import RAFTDecoderMask
    decm = RAFTDecoderMask(net_type='Basic', num_levels=4, radius=4, iters=2, corr_lookup_cfg=dict(align_corners=True),
                           gru_type='SeqConv', act_cfg=dict(type='ReLU'))
    sdm = RO.make_raft_decoder_mask_weights(3)
    missing, unexpected = decm.load_state_dict(sdm, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    decm.eval()
    inputs = RO.make_raft_inputs(3, 2, 16, 16)
    with torch.no_grad():
        ref_f, ref_o = decm(*inputs)
        my_f, my_o = RO.raft_decoder_mask_forward(sdm, *inputs, iters=2)
    outm = {}
    for i in range(2):
        ef, eo = float((ref_f[i] - my_f[i]).abs().max()), float((ref_o[i] - my_o[i]).abs().max())
        assert ef < 1e-4 and eo < 1e-5, (i, ef, eo)
        outm[f'upflow_{i}'] = ref_f[i].numpy()
        outm[f'upocc_{i}'] = ref_o[i].numpy()
        print(f'RAFTDecoderMask iteration {i}: |restatement - reference| flow {ef:.2e}, occlusion {eo:.2e}')
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'raft_decoder_mask_b2_16x16_it2.npz'), **outm)


if __name__ == '__main__':
    main()
