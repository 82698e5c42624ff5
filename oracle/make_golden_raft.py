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


if __name__ == '__main__':
    main()
