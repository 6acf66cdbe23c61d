import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


class Golden:
    """Lazy access to the fixtures recorded from the reference by oracle/make_golden.py."""

    def __init__(self):
        self._files = {}
        with open(os.path.join(GOLDEN_DIR, 'index.json')) as fd:
            self.index = json.load(fd)

    def __call__(self, key):
        group = key.split('/', 1)[0]
        if group not in self._files:
            self._files[group] = np.load(os.path.join(GOLDEN_DIR, f'{group}.npz'))
        return self._files[group][key]

    def has(self, key):
        group = key.split('/', 1)[0]
        self(f'{group}/__probe__') if False else None
        if group not in self._files:
            self._files[group] = np.load(os.path.join(GOLDEN_DIR, f'{group}.npz'))
        return key in self._files[group].files


@pytest.fixture(scope='session')
def golden():
    return Golden()
