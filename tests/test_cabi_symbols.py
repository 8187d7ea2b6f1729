"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
include/instrain_b200.h declares; struct layouts agree between the header, the ctypes mirror and the oracle.
No compute call is made here (no GPU in the build container)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from instrain_b200 import build
    build.build()
    from instrain_b200 import _cabi
    return _cabi.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "instrain_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from instrain_b200 import _cabi
    syms = header_symbols()
    assert syms, "no prototypes parsed from the header"
    for s in syms:
        assert hasattr(lib, s), "symbol %s declared in include/instrain_b200.h is not exported" % s
    assert sorted(_cabi.EXPORTS) == syms


def test_abi_version_and_row_layouts(lib):
    from instrain_b200 import _cabi
    from oracle import restate
    assert lib.isb_abi_version() == 2                                     # 2: + clonTR, normalized LD columns, seed
    assert _cabi.SNV_DT == restate.SNV_DT and _cabi.LD_DT == restate.LD_DT_FULL and _cabi.LD_DT.itemsize == 64
    assert ctypes.sizeof(_cabi.IsbBatch) == 96 and ctypes.sizeof(_cabi.IsbParams) == 40
    assert ctypes.sizeof(_cabi.IsbResult) == 112


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from instrain_b200 import _cabi
    from instrain_b200.engine import Engine
    with pytest.raises(_cabi.IsbError):
        Engine(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "instrain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)
                assert "oracle/" not in src.replace("tests/golden", ""), os.path.join(dirpath, f)


def test_null_model_lut_matches_fixture(null_lut):
    from instrain_b200.null_model import load_lut
    lut, dflt = load_lut()
    assert dflt == null_lut[1] and np.array_equal(lut, null_lut[0])
    # any other fdr resolves from the packaged probability table, as the reference resolves its bundled NullModel.txt
    # whatever --fdr says (profile_controller.py:72-73); P2's fixture holds the reference's own dict for fdr 1e-3
    z = np.load(os.path.join(ROOT, "tests", "golden", "c1_G1_params_P2.npz"))
    lut3, dflt3 = load_lut(fdr=1e-3)
    assert dflt3 == int(z["lut_default"]) and np.array_equal(lut3, z["lut"].astype(np.int32))
    lut9, dflt9 = load_lut(fdr=1e-9)
    assert dflt9 >= dflt3 and (lut9[lut9 >= 0] >= 0).all()
