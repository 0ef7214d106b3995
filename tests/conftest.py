import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope='session')
def golden():
    return load_golden


def assert_close(x, ref, rel=1e-6, scale=1.0, what=''):
    """|x - ref| <= rel * max(|ref|, scale)  (SURVEY 8c tolerance note)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    tol = rel * np.maximum(np.abs(ref), scale)
    bad = np.abs(x - ref) > tol
    assert not bad.any(), '%s: %d/%d out of tolerance, worst |d|=%.3e (tol %.3e)' % (
        what, bad.sum(), bad.size, np.abs(x - ref).max(), tol.min())


def wrap_pi(x):
    return (np.asarray(x) + np.pi) % (2 * np.pi) - np.pi
