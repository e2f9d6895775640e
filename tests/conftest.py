import json
import os
import sys

import numpy as np
import pytest

TESTS = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(TESTS)
GOLDEN = os.path.join(TESTS, "golden")
for p in (REPO, TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_small():
    with open(os.path.join(GOLDEN, "small_cases.json")) as fh:
        meta = json.load(fh)
    outputs = np.load(os.path.join(GOLDEN, "small_outputs.npz"))
    maps = np.load(os.path.join(GOLDEN, "small_maps.npz"))
    return meta, outputs, maps


@pytest.fixture(scope="session")
def golden_full():
    with open(os.path.join(GOLDEN, "full_configs.json")) as fh:
        return json.load(fh)


def mismatch_report(got, want):
    """(fraction of pixels bit-exact, max abs channel difference, number of differing pixels)"""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    diff = got != want
    px = diff.any(axis=-1) if diff.ndim == 3 else diff
    n_bad = int(px.sum())
    max_abs = int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) if n_bad else 0
    return 1.0 - n_bad / px.size, max_abs, n_bad
