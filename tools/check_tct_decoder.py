"""Decoder outputs with the transposed small-N kernel (SCFLOW_TC_T=1) against the pixels-as-rows kernel, per iteration."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests.test_gpu_decoder import NAMES, _call, _inputs  # noqa: E402
from tests.util import build_decoder_from_oracle_weights  # noqa: E402

seed, b, h, w, iters = 0, 2, 256, 256, 4
res = {}
for t in ('0', '2'):
    os.environ['SCFLOW_TC_T'] = t
    dec, _ = build_decoder_from_oracle_weights(seed, iters, precision=1)
    scene, f = _inputs(seed, b, h, w)
    res[t] = _call(dec, scene, f, b, h, w)
for nm, l0, l1 in zip(NAMES, res['0'], res['2']):
    print(nm, ' '.join(f'{(a - c).abs().max().item():.2e}' for a, c in zip(l0, l1)))
