import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu via gpurun)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "pbmc3k50_hallmarks.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def fixture_mats(golden):
    """The reference's bundled pbmc3k-50cells X (7728 x 50 dgCMatrix) and hallmarks matG (4386 x 50)."""
    import scipy.sparse as sp
    g = golden
    X = sp.csc_matrix((g["X_data"], g["X_indices"], g["X_indptr"]), shape=tuple(g["X_shape"]))
    G = sp.csc_matrix((np.ones(g["G_indices"].size), g["G_indices"], g["G_indptr"]), shape=tuple(g["G_shape"]))
    return X, [str(s) for s in g["X_rownames"]], [str(s) for s in g["X_colnames"]], \
        G, [str(s) for s in g["G_rownames"]], [str(s) for s in g["G_colnames"]]


@pytest.fixture(scope="session")
def gpu_ctx():
    from plaid_b200 import Context
    return Context(0)


def rel_err(a, b):
    """max |a-b| / max(|b|, scale) with scale = median magnitude of b (guards near-zero entries)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    if a.size == 0:
        return 0.0
    m = ~nan_b
    if not m.any():
        return 0.0
    scale = max(np.median(np.abs(b[m])), 1e-300)
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), scale)))
