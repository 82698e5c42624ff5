import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


# Hot-path rows run first, the widened rows (§8f: formatting, loss, RAFT baselines) after them, so that with
# `-x` a failure in a peripheral row can never hide the parity evidence of the path itself.
_ORDER = ('test_oracle_golden', 'test_boundary_cpu', 'test_gpu_kernels', 'test_gpu_tc', 'test_gpu_fused', 'test_gpu_decoder',
          'test_gpu_full_size', 'test_gpu_encoder', 'test_dist_gloo', 'test_bench_contract', 'test_train', 'test_loss',
          'test_raft_upsample', 'test_format')


def _rank(item):
    name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
    return _ORDER.index(name) if name in _ORDER else len(_ORDER)


def pytest_collection_modifyitems(config, items):
    items.sort(key=_rank)          # stable: keeps the order inside a file
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
