"""Shared test plumbing: import paths, the ``gpu`` marker, golden-fixture access."""

import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PACKAGE_DIR = os.path.join(REPO, "model-predictive-control-for-bipedal-locomotion_b200")
for path in (REPO, PACKAGE_DIR):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(REPO, "tests", "golden", name + ".npz"))
    return load


@pytest.fixture(scope="session")
def default_config():
    from mpc_bipedal.config import MPCConfig
    return MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
