"""bench.py contract on a CPU-only machine: the reference arm prints ONE JSON line with the keys the driver reads, and the
product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-sample', '2')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'pairs/s' and d['higher_is_better'] is True and d['n_gpus'] == 1
    assert d['steps'] == 1 and d['warmup'] == 1 and d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'model' not in d['config']


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a machine without a GPU')
def test_product_arm_fails_loudly_without_a_gpu():
    r = _run('--steps', '1', '--warmup', '1', '--no-cpu-baseline', timeout=300)
    assert r.returncode != 0
    assert 'no CPU fallback' in (r.stderr + r.stdout)


def test_reference_arm_under_torchrun_prints_on_rank0_only():
    """N > 1: rank 0 alone measures and prints; the other ranks exit 0 without work."""
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29541', os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '1', '--cpu-sample', '2'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
