import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_batch(which):
    b = dict(np.load(os.path.join(GOLDEN, "c1_%s_batch.npz" % which)))
    e = dict(np.load(os.path.join(GOLDEN, "c1_%s_expected.npz" % which)))
    return b, e


def load_lut():
    z = np.load(os.path.join(GOLDEN, "null_lut_fdr1e-06.npz"))
    return z["lut"].astype(np.int32), int(z["default"])


@pytest.fixture(scope="session")
def null_lut():
    return load_lut()


def sort_snv(rows):
    return rows[np.lexsort((rows["mm"], rows["pos"]))]


def sort_ld(rows):
    return rows[np.lexsort((rows["mm"], rows["pos_b"], rows["pos_a"]))]


def assert_snv_equal(a, b):
    """Bit-exact comparison of two SNV row arrays (any order)."""
    a, b = sort_snv(a), sort_snv(b)
    assert len(a) == len(b), (len(a), len(b))
    for f in ("pos", "mm", "cnt", "ref", "con", "var", "allele_count", "cls", "cryptic"):
        assert np.array_equal(a[f], b[f]), f


def assert_ld_equal(a, b, tol=1e-6):
    """Integer fields bit-exact; r2 / d_prime within `tol` (north_star: 1e-6), NaN pattern identical."""
    a, b = sort_ld(a), sort_ld(b)
    assert len(a) == len(b), (len(a), len(b))
    for f in ("pos_a", "pos_b", "mm", "c_AB", "c_Ab", "c_aB", "c_ab", "allele_A", "allele_a", "allele_B", "allele_b"):
        assert np.array_equal(a[f], b[f]), f
    for f in ("r2", "d_prime"):
        assert np.array_equal(np.isnan(a[f]), np.isnan(b[f])), f
        assert np.allclose(a[f], b[f], rtol=0, atol=tol, equal_nan=True), f
