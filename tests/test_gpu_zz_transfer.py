"""profile_bam with the opt-in host -> device formats (kwargs["b200_transfer"]): the reference-delta transfer format
(C++ host encoder -> K0d -> K1r) and column words laid out on the host (C++ host conversion -> K1c) give exactly the
tables of the default read-major path on a real BAM; so does packing the scaffolds on several host threads.  (Named to sort last; the pieces are covered by test_gpu_reads.py / test_gpu_cols.py / the CPU suite.  The chunk
pipeline tests ran green on a B200 in round 2: gpurun_out/r2l_pytest_exp.log.)"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
@pytest.mark.parametrize("transfer,threads", [("delta", 1), ("cols", 1), ("segments", 3)])
def test_profile_bam_transfer_formats(transfer, threads):
    from instrain_b200.profile import profile_bam
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    kw = dict(s2s=seqs, min_cov=5, min_freq=0.05, min_snp=20, window_length=10000, store=False)
    a = profile_bam(bam, None, rdic, None, **kw)
    b = profile_bam(bam, None, rdic, None, b200_transfer=transfer, packer_threads=threads, **kw)
    assert a.scaffold_list == b.scaffold_list and len(a.raw_snp_table) > 1000 and len(a.raw_linkage_table) > 1000
    for name, key in (("raw_snp_table", ["scaffold", "position", "mm"]),
                      ("raw_linkage_table", ["scaffold", "position_A", "position_B", "mm"]),
                      ("cumulative_scaffold_table", ["scaffold", "mm"])):
        x = getattr(a, name).sort_values(key).reset_index(drop=True)
        y = getattr(b, name).sort_values(key).reset_index(drop=True)
        assert list(x.columns) == list(y.columns) and len(x) == len(y), name
        for c in x.columns:
            if x[c].dtype.kind == "f":
                assert np.array_equal(x[c].values, y[c].values, equal_nan=True), (name, c)
            else:
                assert (x[c].values == y[c].values).all(), (name, c)
    for s in a.scaffold_list:
        for m in a.scaffolds[s].covT:
            assert a.scaffolds[s].covT[m].equals(b.scaffolds[s].covT[m])
            assert a.scaffolds[s].clonT[m].equals(b.scaffolds[s].clonT[m])


@pytest.mark.parametrize("skip_mm,lean", [(True, True), (True, False), (False, False)])
def test_cols_chunk_pipeline_equals_single_pass(skip_mm, lean):
    """ISB_PIPELINE on the column-word path (opt-in): K1c of all chunks back to back on the main stream, K3 (+ K2 when not
    fused) of chunk c on a second stream underneath; the tables must equal the single-pass ones exactly."""
    from conftest import assert_ld_equal, assert_snv_equal, load_lut
    from instrain_b200 import synth as dsynth
    from instrain_b200.engine import Engine
    lut = load_lut()
    eng = Engine(0, lut[0], lut[1])
    try:
        d = dsynth.generate(0, 300000, 15, 12, 0.01, 99, skip_mm=skip_mm, events=False, reads=True)   # 4.5e6 positions
        cd = dsynth.reads_to_cols_device(eng, d)
        ev = dict(pair_mm=d["pair_mm"].cpu().numpy())
        ref, spl = d["ref_codes"].cpu().numpy(), d["splits"].cpu().numpy()
        M = int(ev["pair_mm"].max()) + 1 if len(ev["pair_mm"]) else 1
        want = ("covT", "clonT", "site_flags", "snv", "ld") if lean else ("counts", "nmask", "covT", "clonT", "site_flags", "snv", "ld")
        a = eng.profile_batch(ev, ref, spl, M=M, want=want, min_cov=3, min_snp=3, cols=cd)
        b = eng.profile_batch(ev, ref, spl, M=M, want=want, min_cov=3, min_snp=3, cols=cd, pipeline=True)
        assert len(a["snv"]) > 1000 and len(a["ld"]) > 100
        for k in ("counts", "nmask", "covT", "site_flags"):
            if k in want:
                assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(a["clonT"].view(np.uint32), b["clonT"].view(np.uint32))
        assert_snv_equal(a["snv"], b["snv"])
        assert_ld_equal(a["ld"], b["ld"], tol=0)
        assert (a["n_sites"], a["n_site_pairs"]) == (b["n_sites"], b["n_site_pairs"])
    finally:
        eng.close()
