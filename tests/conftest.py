import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def fingerprints():
    with open(os.path.join(GOLDEN, 'fingerprints.json')) as fh:
        return json.load(fh)


@pytest.fixture(scope='session')
def golden():
    return dict(np.load(os.path.join(GOLDEN, 'reference_golden.npz')))


def _scene(name):
    z = np.load(os.path.join(GOLDEN, 'scene_%s_u16.npz' % name))
    return tuple(z[k].astype(np.float32) for k in ('im10', 'im20', 'im60'))


@pytest.fixture(scope='session')
def malmo():
    return _scene('malmo')


@pytest.fixture(scope='session')
def shark():
    return _scene('shark')


@pytest.fixture(scope='session', params=['malmo', 'shark'])
def scene(request):
    """The two scenes of the reference's data/ directory that are present in the mount (name, im10, im20, im60)."""
    return (request.param,) + _scene(request.param)


def sha16(a):
    import hashlib
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
