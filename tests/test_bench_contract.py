"""bench.py's reference arm (the CPU legs, runnable without a GPU): one JSON line with the keys the contract names, for the
headline workload and for the training step; the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(*args, timeout=300):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT)
    return r


def test_reference_arm_training_line():
    r = _run('--impl', 'reference', '--workload', 'train', '--steps', '1', '--warmup', '0')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d['impl'] == 'reference' and d['metric'] == 'train_samples_per_s'
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {"value": d['value'], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d['gpu_launches'] == 0 and d['config']['cpu_sample_only'] is True


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'train', '--gpus', '2'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    for extra in ([], ['--workload', 'train']):
        r = _run('--steps', '1', '--warmup', '0', *extra, timeout=120)
        assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
