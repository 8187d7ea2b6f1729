"""The seam from the REFERENCE's side: inStrain's own ProfileController.run_profile() -> write_output()
(/root/reference/inStrain/controller.py:324-360) with `inStrain.profile.profile_bam` re-pointed at this package's shim the
way INTEGRATION.md section 3 tells a maintainer to, the CUDA engine answered by the oracle (test stub; no GPU here).

What it pins: the object profile_bam hands back is one the controller can keep using as `self.ISP` -- `.generate(t)` for
every table write_output asks for ('SNVs', 'scaffold_info', 'linkage', 'gene_info', 'mapping_info'), `.get`, `.store`,
`.get_location` -- both as this package's ProfileStore and as the reference's own SNVprofile class opened on the directory
the shim wrote.  Build container only (needs /root/reference)."""
import json
import os
import types

import pandas as pd
import pytest

from conftest import GOLDEN
from oracle import ref_harness
from test_profile_host_cpu import OracleEngine

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present")


class _Engine(OracleEngine):
    def __init__(self, *a, **k):
        super().__init__()

    def close(self):
        pass


def _controller(tmp_path, monkeypatch, native):
    ref_harness.load_reference()                                   # stub finder for pysam / h5py / Bio / matplotlib ...
    import inStrain.controller
    import inStrain.profile
    import inStrain.profile.fasta
    import inStrain.SNVprofile
    import instrain_b200.profile as P
    monkeypatch.setattr(P, "Engine", _Engine)
    monkeypatch.setenv("ISB_NATIVE_STORE", "1" if native else "0")
    monkeypatch.setattr(inStrain.profile, "profile_bam", P.profile_bam)
    rdic = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_r2m.json")))
    seqs = json.load(open(os.path.join(GOLDEN, "c1_G1_subset_seqs.json")))
    out = str(tmp_path / ("seam_native.IS" if native else "seam_ref.IS"))
    args = types.SimpleNamespace(output=out, min_cov=5, min_freq=0.05, fdr=1e-6, min_snp=20, window_length=10000,
                                 skip_mm_profiling=False, processes=1, debug=False, min_read_ani=0.95, min_mapq=-1,
                                 max_insert_relative=3, min_insert=50, pairing_filter="paired_only")
    pc = inStrain.controller.ProfileController(args)
    # what validate_arguments / profile_filter_reads leave behind (controller.py:171-322), without the BAM / FASTA tools
    pc.ISP = inStrain.SNVprofile.SNVprofile(out)                   # the reference's own object, as validate_arguments makes it
    pc.kwargs = vars(args)
    pc.bam = os.path.join(GOLDEN, "c1_G1_subset.bam")
    pc.scaff2sequence = seqs
    pc.scaffold2pairs = {s: len(v) for s, v in rdic.items()}
    pc.Rdic = rdic
    rows = []
    for s, seq in seqs.items():                                    # the split table load_fasta builds (fasta.py:30-52)
        for i, (a, b) in enumerate(inStrain.profile.fasta.iterate_splits(len(seq), args.window_length)):
            rows.append({"scaffold": s, "split_number": i, "start": a, "end": b})
    pc.fasta_db = pd.DataFrame(rows)
    return pc, out, rdic


@pytest.mark.parametrize("native", [True, False])
def test_reference_controller_runs_through_the_shim(tmp_path, monkeypatch, native):
    pc, out, rdic = _controller(tmp_path, monkeypatch, native)
    pc.run_profile()                                               # profile_bam + write_output: must not raise
    isp = pc.ISP
    import inStrain.SNVprofile
    from instrain_b200.store import ProfileStore
    assert isinstance(isp, ProfileStore if native else inStrain.SNVprofile.SNVprofile)
    for meth in ("generate", "get", "store", "get_location"):
        assert callable(getattr(isp, meth))
    assert isp.get("object_type") == "profile" and sorted(isp.get("scaffold_list")) == sorted(rdic)
    assert os.path.isdir(isp.get_location("output"))
    base = os.path.join(out, "output", os.path.basename(out) + "_")
    for t in ("SNVs", "scaffold_info", "linkage"):
        db = pd.read_csv(base + t + ".tsv", sep="\t")
        assert len(db) == len(rdic) if t == "scaffold_info" else len(db) > 30, t
    snvs = pd.read_csv(base + "SNVs.tsv", sep="\t")
    assert not snvs.duplicated(["scaffold", "position"]).any()
    assert list(snvs.columns[:4]) == ["scaffold", "position", "position_coverage", "allele_count"]
    assert not os.path.exists(base + "gene_info.tsv")              # no genes profiled: generate('gene_info') is a logged no-op
    # the controller's later steps keep storing into the same object (profile_genome_wide does, controller.py:381-398)
    isp.store("note", {"a": 1}, "dictionary", "test attribute")
    assert isp.get("note") == {"a": 1}
    # in-memory result of the run rides along
    assert len(isp.result.raw_snp_table) == len(isp.get("raw_snp_table")) > 1000


def test_reference_generate_equals_native_generate(tmp_path, monkeypatch):
    """The reference's SNVprofile.generate, run on the directory the shim wrote, gives the tables this package's own
    generate writes (same rows, same column order)."""
    pc, out, _ = _controller(tmp_path, monkeypatch, native=True)
    pc.run_profile()
    import inStrain.SNVprofile
    from instrain_b200.store import SNVprofileStore
    ref_isp = inStrain.SNVprofile.SNVprofile(out)
    own = SNVprofileStore(out)
    for t, key in (("SNVs", ["scaffold", "position"]), ("scaffold_info", ["scaffold"]), ("linkage", ["scaffold", "position_A", "position_B"])):
        a = ref_isp.generate(t, store=False, return_table=True).sort_values(key).reset_index(drop=True)
        b = own.generate(t, store=False, return_table=True).sort_values(key).reset_index(drop=True)
        # the reference appends the columns outside its order list in SET order (SNVprofile.py:1162-1163: hash order, it
        # changes from run to run); the ordered head must agree exactly, the tail as a set
        n_head = len([c for c in own._OUTPUTS[t][2] if c in set(b.columns)])
        assert list(a.columns[:n_head]) == list(b.columns[:n_head]) and set(a.columns) == set(b.columns), t
        assert len(a) == len(b) > 5
        pd.testing.assert_frame_equal(a[list(b.columns)], b, check_dtype=False)
