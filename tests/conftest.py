"""Shared pytest configuration: registers the ``gpu`` marker and common fixtures."""

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name: str):
    data = np.load(os.path.join(GOLDEN, name))
    return {k: data[k] for k in data.files}


@pytest.fixture(scope="session")
def golden_small():
    return load_golden("small.npz")


@pytest.fixture(scope="session")
def golden_eviction():
    return load_golden("eviction.npz")


@pytest.fixture(scope="session")
def golden_c1():
    return load_golden("c1.npz")
