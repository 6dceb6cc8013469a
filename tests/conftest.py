"""Shared pytest plumbing: the `gpu` marker, golden loaders and the reference's tolerance rule."""
import lzma
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_wts(name):
    """Reference golden weight file (raw N^3 x N^3 doubles, /root/reference/src/weights.c:78-88)."""
    raw = lzma.decompress(open(os.path.join(GOLDEN, name + ".xz"), "rb").read())
    return np.frombuffer(raw, dtype=np.float64).copy()


def load_moments(name):
    return np.loadtxt(os.path.join(GOLDEN, name), comments="#")


def check_diff_two_sided(got, want):
    """tests/check_diff.py:19-29 made two-sided: |d| <= max(1e-14, 1e-6*|want|). Returns #violations."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    tol = np.maximum(1e-14, 1e-6 * np.abs(want))
    return int((np.abs(got - want) > tol).sum())


def relmax(a, b):
    """Normwise relative difference  max|a-b| / max|b|."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="session")
def ref_vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def W_bkw8():
    return load_wts("N8_isotropic_L_v5_lambda0.wts")


@pytest.fixture(scope="session")
def W_heat8():
    return load_wts("N8_isotropic_L_v9_lambda1.wts")


def seeded_f(v, seed, noise=0.05):
    """Same seeded distribution as tests/golden/make_golden.py::seeded_f."""
    rng = np.random.default_rng(seed)
    vx, vy, vz = np.meshgrid(v, v, v, indexing="ij")
    f = np.exp(-((vx - 0.3) ** 2 + (vy + 0.2) ** 2 + vz ** 2) / 1.7) / 7.0
    f *= 1.0 + noise * rng.standard_normal(f.shape)
    return np.ascontiguousarray(f.reshape(-1))
