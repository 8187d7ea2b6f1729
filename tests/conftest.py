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
    fields = ("r2", "d_prime")
    if "r2_normalized" in (a.dtype.names or ()) and "r2_normalized" in (b.dtype.names or ()):
        fields += ("r2_normalized", "d_prime_normalized")     # re-drawn columns: same counter-based draws on both sides
    for f in fields:
        assert np.array_equal(np.isnan(a[f]), np.isnan(b[f])), f
        assert np.allclose(a[f], b[f], rtol=0, atol=tol, equal_nan=True), f


def assert_clontr_equal(got, exp):
    """Rarefied clonality (float32, NaN = unset): bit-identical to the oracle's restatement of the same draws."""
    ok = ~np.isnan(exp)
    assert np.array_equal(np.isnan(got), ~ok)
    assert np.array_equal(got[ok].view(np.uint32), exp[ok].view(np.uint32))


def basewise_digest(values, index):
    """sha1 over the 2 x N array the reference stores for one (scaffold, mm) dataset of covT / clonT
    (`np.array([series.values, series.index])`, SNVprofile.py:717-733): int64 for coverage, float64 for clonality."""
    import hashlib
    values = np.asarray(values)
    dt = "<f8" if values.dtype.kind == "f" else "<i8"
    return np.frombuffer(hashlib.sha1(values.astype(dt).tobytes() + np.asarray(index).astype(dt).tobytes()).digest(), np.uint8)


def assert_basewise_matches_digest(which, names, offs, lens, covT, clonT, nmask):
    """Dense covT / clonT / nmask of the C1 batch -> per-scaffold {mm: Series} (instrain_b200.tables) -> compared with the
    digests of the reference's OWN stored covT.hd5 / clonT.hd5 (tests/golden/c1_<set>_hd5_digest.npz): same dataset
    names (levels, empty ones included), same lengths, same values and positions bit for bit."""
    from instrain_b200 import tables
    z = np.load(os.path.join(GOLDEN, "c1_%s_hd5_digest.npz" % which))
    gold = {k: (int(n1), bytes(d1), int(n2), bytes(d2)) for k, n1, d1, n2, d2 in
            zip(z["names"], z["cov_n"], z["cov_sha"], z["clon_n"], z["clon_sha"])}
    seen = set()
    for name, off, L in zip(names, offs, lens):
        sl = slice(int(off), int(off) + int(L))
        lv = tables.present_levels(covT[sl], nmask[sl])
        cov = tables.basewise(covT[sl], "coverage", lv)
        clon = tables.basewise(clonT[sl], "clonality", lv)
        for mm in lv:
            key = "%s::%d" % (name, mm)
            assert key in gold, "level not in the reference's store: " + key
            seen.add(key)
            n1, d1, n2, d2 = gold[key]
            assert len(cov[mm]) == n1 and bytes(basewise_digest(cov[mm].values, cov[mm].index.values)) == d1, "covT " + key
            assert len(clon[mm]) == n2 and bytes(basewise_digest(clon[mm].values, clon[mm].index.values)) == d2, "clonT " + key
    assert seen == set(gold), sorted(set(gold) - seen)[:5]
    return len(seen)
