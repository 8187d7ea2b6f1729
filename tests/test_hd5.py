"""the covT / clonT HDF5 stores (instrain_b200/hd5.py) -- CPU tests.

  * the oracle's basewise coverage / clonality equal the reference's OWN stored covT.hd5 / clonT.hd5 (committed digests,
    tests/golden/c1_<set>_hd5_digest.npz, made by tests/golden/make_golden.py with the repo's HDF5 reader);
  * the reader parses the reference's stored files (libhdf5 output) when /root/reference is there;
  * writer -> reader round trips: dtypes, empty datasets, edge chunks, group B-trees of 1 .. 3 levels, and the
    SNVprofile-style store_special / load_special pair (SNVprofile.py:717-786).
"""
import os
import struct

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN, assert_basewise_matches_digest, basewise_digest, load_batch, load_lut
from instrain_b200 import hd5
from oracle import restate

REF_RAW = "/root/reference/test/test_data/N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010%s.forRC.IS/raw_data"


@pytest.mark.parametrize("which", ["G1", "G2"])
def test_oracle_basewise_equals_reference_store(which):
    lut, dflt = load_lut()
    b, _ = load_batch(which)
    exp = restate.profile_events(b, b["ref_codes"], lut, dflt, b["splits"], do_linkage=False)
    n = assert_basewise_matches_digest(which, b["scaffold_names"], b["scaffold_off"], b["scaffold_len"],
                                       exp["covT"], exp["clonT"], exp["nmask"])
    assert n == {"G1": 1267, "G2": 1429}[which]


@pytest.mark.parametrize("which", ["G1", "G2"])
@pytest.mark.parametrize("name", ["covT", "clonT"])
def test_reader_on_reference_files(which, name):
    path = os.path.join(REF_RAW % which, name + ".hd5")
    if not os.path.exists(path):
        pytest.skip("reference test data not present on this machine")
    ds = hd5.read_hd5(path)
    z = np.load(os.path.join(GOLDEN, "c1_%s_hd5_digest.npz" % which))
    assert sorted(ds) == list(z["names"])
    pre = "cov" if name == "covT" else "clon"
    for k, n, sha in zip(z["names"], z[pre + "_n"], z[pre + "_sha"]):
        a = ds[str(k)]
        assert a.shape == (2, n) and a.dtype == (np.int64 if name == "covT" else np.float64)
        assert bytes(basewise_digest(a[0], a[1])) == bytes(sha), k


def _random_sets(rng, n_sets, max_n):
    out = {}
    for i in range(n_sets):
        n = int(rng.integers(0, max_n)) if i % 7 else 0
        idx = np.sort(rng.choice(max(4 * n, 1), n, replace=False))
        if i % 2:
            out["scaffold_%d::%d" % (i // 3, i % 3)] = np.array([rng.integers(1, 500, n), idx])
        else:
            out["scaffold_%d::%d" % (i // 3, i % 3)] = np.array([rng.random(n).astype(np.float32), idx])
    return out


@pytest.mark.parametrize("n_sets,max_n", [(1, 50), (9, 3000), (300, 400), (9000, 6)])
def test_write_read_round_trip(tmp_path, n_sets, max_n):
    rng = np.random.default_rng(n_sets)
    ds = _random_sets(rng, n_sets, max_n)
    path = str(tmp_path / "t.hd5")
    size = hd5.write_hd5(path, ds)
    assert size == os.path.getsize(path)
    back = hd5.read_hd5(path)
    assert set(back) == set(ds)
    for k, a in ds.items():
        assert back[k].shape == a.shape, k
        assert back[k].dtype == (np.float64 if a.dtype.kind == "f" else np.int64)
        assert np.array_equal(back[k], a), k
    # the file is self-consistent the way libhdf5 checks it on open: signature, end-of-file address, root symbol table
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and struct.unpack_from("<Q", raw, 40)[0] == len(raw)
    btree, heap = struct.unpack_from("<QQ", raw, 80)
    assert raw[btree:btree + 4] == b"TREE" and raw[heap:heap + 4] == b"HEAP"


def test_large_dataset_chunks(tmp_path):
    """A 1 Mb scaffold's worth of positions: 64 chunks of 1 x 31250, the last ones partial."""
    n = 1_000_003
    a = np.array([np.arange(n) % 97 + 1, np.arange(n)])
    path = str(tmp_path / "big.hd5")
    hd5.write_hd5(path, {"s::0": a})
    assert np.array_equal(hd5.read_hd5(path)["s::0"], a)


def test_store_special_round_trip(tmp_path):
    """scaffold -> mm -> Series, as SNVprofile.store(..., 'special') receives it."""
    rng = np.random.default_rng(5)
    obj = {}
    for s in ("scaffA", "scaffB|with odd:chars", "c"):
        obj[s] = {}
        for mm in (0, 2, 11):
            n = int(rng.integers(0, 200))
            idx = np.sort(rng.choice(1000, n, replace=False))
            obj[s][mm] = pd.Series(rng.integers(1, 90, n).astype("int32"), index=idx)
    path = str(tmp_path / "covT.hd5")
    hd5.store_special(path, obj)
    back = hd5.load_special(path)
    assert set(back) == set(obj)
    for s in obj:
        assert set(back[s]) == set(obj[s])
        for mm in obj[s]:
            assert np.array_equal(back[s][mm].values, obj[s][mm].values)
            assert np.array_equal(back[s][mm].index.values, obj[s][mm].index.values)
    only = hd5.load_special(path, scaffolds=["c"])
    assert list(only) == ["c"]


def test_reader_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.hd5"
    p.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(ValueError):
        hd5.read_hd5(str(p))


def test_threaded_writer_is_byte_identical(tmp_path):
    """write_hd5 deflates the datasets on a thread pool; the file must not depend on the thread count."""
    rng = np.random.default_rng(3)
    ds = {"s%d::%d" % (i, m): np.array([rng.integers(0, 200, n), np.arange(n)]) for i in range(12) for m in range(3)
          for n in [int(rng.integers(0, 5000))]}
    ds["f::0"] = np.array([rng.random(700), np.arange(700)], dtype=np.float64)
    sizes = [hd5.write_hd5(str(tmp_path / ("t%d.hd5" % t)), ds, threads=t) for t in (1, 3, 8)]
    blobs = [open(str(tmp_path / ("t%d.hd5" % t)), "rb").read() for t in (1, 3, 8)]
    assert sizes[0] == len(blobs[0]) and blobs[0] == blobs[1] == blobs[2]
    back = hd5.read_hd5(str(tmp_path / "t8.hd5"))
    assert set(back) == set(ds) and all(np.array_equal(back[k], ds[k].astype(back[k].dtype)) for k in ds)


def _plain(x):
    """describe_hd5 output -> JSON-comparable (tuples -> lists, bytes -> hex)."""
    if isinstance(x, dict):
        return {str(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    if isinstance(x, (bytes, bytearray)):
        return bytes(x).hex()
    if isinstance(x, (np.integer,)):
        return int(x)
    return x


@pytest.mark.parametrize("which", ["G1", "G2"])
@pytest.mark.parametrize("name", ["covT", "clonT"])
def test_writer_structure_equals_reference_files(tmp_path, which, name):
    """The writer against libhdf5's own output, structure by structure (h5py / libhdf5 are not in this image, so no
    library can open the written file here): the reference's stored covT.hd5 / clonT.hd5 (h5py, libver='earliest') are
    read, the same datasets are written by write_hd5, and the address-free descriptions of the two files
    (hd5.describe_hd5: superblock fields, root header messages, local heap, group B-tree depth and symbol-node fill,
    and per dataset every header message with its version byte and decoded fields -- dataspace, datatype bits and
    properties, fill value, filter pipeline with the deflate client data, chunked layout with h5py's chunk shape -- and
    the chunk B-tree's node type, depth, filter masks and chunk offsets) must be EQUAL for all 1267 / 1429 datasets."""
    path = os.path.join(REF_RAW % which, name + ".hd5")
    if not os.path.exists(path):
        pytest.skip("reference test data not present on this machine")
    ref = hd5.describe_hd5(path)
    ds = hd5.read_hd5(path)
    out = str(tmp_path / "w.hd5")
    hd5.write_hd5(out, ds, threads=2)
    mine = hd5.describe_hd5(out)
    assert mine["superblock"] == ref["superblock"] and mine["root_messages"] == ref["root_messages"]
    assert mine["heap"] == ref["heap"] and mine["group"] == ref["group"] and mine["n_datasets"] == ref["n_datasets"]
    bad = [k for k in ref["datasets"] if mine["datasets"].get(k) != ref["datasets"][k]]
    assert not bad, (bad[:3], ref["datasets"][bad[0]], mine["datasets"].get(bad[0]))
    back = hd5.read_hd5(out)
    assert all(np.array_equal(ds[k], back[k]) and ds[k].dtype == back[k].dtype for k in ds)


def test_writer_structure_equals_committed_reference_description(tmp_path):
    """The same comparison where /root/reference is absent: tests/golden/hd5_structure_G1.json holds describe_hd5 of the
    reference's stored G1 covT.hd5 / clonT.hd5 for the datasets of the first scaffolds (made by make_golden.py); the
    datasets are rebuilt from the oracle's basewise tables (equal to the stored ones: test_oracle_basewise_equals_
    reference_store) and written by write_hd5."""
    import json
    from instrain_b200 import tables
    gold = json.load(open(os.path.join(GOLDEN, "hd5_structure_G1.json")))
    lut, dflt = load_lut()
    b, _ = load_batch("G1")
    exp = restate.profile_events(b, b["ref_codes"], lut, dflt, b["splits"], do_linkage=False)
    cov_ds, clon_ds = {}, {}
    for sname, off, L in zip(b["scaffold_names"], b["scaffold_off"], b["scaffold_len"]):
        sl = slice(int(off), int(off) + int(L))
        lv = tables.present_levels(exp["covT"][sl], exp["nmask"][sl])
        cov = tables.basewise(exp["covT"][sl], "coverage", lv)
        clon = tables.basewise(exp["clonT"][sl], "clonality", lv)
        for mm in lv:
            cov_ds["%s::%d" % (sname, mm)] = np.array([cov[mm].values, cov[mm].index])
            clon_ds["%s::%d" % (sname, mm)] = np.array([clon[mm].values, clon[mm].index])
    for name, ds in (("covT", cov_ds), ("clonT", clon_ds)):
        out = str(tmp_path / (name + ".hd5"))
        hd5.write_hd5(out, ds)
        mine = _plain(hd5.describe_hd5(out))
        g = gold[name]
        for k in ("superblock", "root_messages", "heap", "group", "n_datasets"):
            assert mine[k] == g[k], (name, k)
        assert len(g["datasets"]) >= 40
        for k, v in g["datasets"].items():
            assert mine["datasets"][k] == v, (name, k)
