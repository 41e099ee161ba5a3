import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden(object):
    """Read-only view of one .npz fixture: g.case('enc_f32') -> dict of tensors."""

    def __init__(self, path):
        self._npz = np.load(path)
        self.names = sorted({k.split('.')[0] for k in self._npz.files})

    def case(self, name):
        out = {}
        prefix = name + '.'
        for k in self._npz.files:
            if k.startswith(prefix):
                out[k[len(prefix):]] = torch.from_numpy(self._npz[k])
        if not out:
            raise KeyError(name)
        return out


@pytest.fixture(scope='session')
def op_golden():
    return Golden(os.path.join(GOLDEN_DIR, 'op_golden.npz'))


@pytest.fixture(scope='session')
def module_golden():
    return Golden(os.path.join(GOLDEN_DIR, 'module_golden.npz'))


def rel_err(a, b):
    """max |a-b| / max |b| — the 'relative' of the north-star tolerances."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))
