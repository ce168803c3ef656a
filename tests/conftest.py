import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything compiled before any test runs (no GPU needed for the build)."""
    import __graft_entry__ as g
    g.build()


def random_csr(M, N, density_rows, seed, max_len=None, empty_frac=0.0):
    """Small seeded CSR with ascending, duplicate-free rows; some rows may be empty."""
    rng = np.random.default_rng(seed)
    row_ptr = [0]
    cols = []
    for u in range(M):
        if rng.random() < empty_frac:
            n = 0
        else:
            n = int(min(N, max(1, rng.geometric(1.0 / density_rows))))
            if max_len:
                n = min(n, max_len)
        c = np.sort(rng.choice(N, size=n, replace=False)).astype(np.int32)
        cols.append(c)
        row_ptr.append(row_ptr[-1] + n)
    return np.asarray(row_ptr, np.int64), (np.concatenate(cols) if cols else np.zeros(0, np.int32)).astype(np.int32)
