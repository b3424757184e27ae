import json
import os
import sys

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy view of one tests/golden/*.npz fixture (made by tools/make_golden.py from the reference)."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
        self.meta = json.loads(str(self.z['meta']))

    def __contains__(self, k):
        return k in self.z.files

    def t(self, k, device='cpu', dtype=None):
        v = torch.from_numpy(np.array(self.z[k]))
        if dtype is not None:
            v = v.to(dtype)
        return v.to(device)

    def keys(self):
        return self.z.files


_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = Golden(name)
    return _cache[name]


def rel_err(a, b):
    """max|a-b| / max|b| -- the parity metric north_star states."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


DT = {'float32': torch.float32, 'float64': torch.float64, 'float16': torch.float16}
