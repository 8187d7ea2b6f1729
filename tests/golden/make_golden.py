"""Generate the committed golden fixtures (build container only; needs /root/reference).

    python tests/golden/make_golden.py

For each of the reference's two primary known-answer sets:
  * c1_<set>_batch.npz     the hot path's INPUT as one scaffold batch: position-major event arrays produced by
                           oracle/pileup_emul.py from the bundled BAM + the stored Rdic.json (sR2M), reference
                           base codes, split table (inStrain/profile/fasta.py:56-73, window 10000)
  * c1_<set>_expected.npz  the reference's OWN stored outputs: raw_snp_table.csv.gz / raw_linkage_table.csv.gz of
                           `...forRC.IS/raw_data/` (inStrain v1.7.0, pysam 0.16.0.1), re-encoded as arrays
                           (random columns r2_normalized / d_prime_normalized dropped, as the reference's tests do)
and null_lut_fdr1e-06.npz = generate_snp_model(NullModel.txt, 1e-6) (snv_utilities.py:14-38) as an int LUT.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bamio, pileup_emul, ref_harness, restate  # noqa: E402
from oracle.validate_against_reference import load_set  # noqa: E402

CLS = {n: i for i, n in enumerate(restate.CLASS_NAMES)}
B = {b: i for i, b in enumerate("ACTG")}


def make_batch(which, window=10000):
    refs, by_tid, seqs, rdic, raw = load_set(which)
    names, lens, offs = [], [], []
    ref_codes, ev_pos, ev_base, ev_qual, ev_rid, pair_mm, splits = [], [], [], [], [], [], []
    off, pair_off = 0, 0
    for tid, (name, length) in enumerate(refs):
        if name not in rdic:
            continue
        seq = seqs[name]
        ev = restate.sort_events(pileup_emul.scaffold_events(by_tid.get(tid, []), rdic[name]))
        names.append(name)
        lens.append(len(seq))
        offs.append(off)
        ref_codes.append(restate.encode_ref(seq))
        ev_pos.append(ev["ref_pos"] + off)
        ev_base.append(ev["base"])
        ev_qual.append(ev["qual"])
        ev_rid.append(ev["read_id"] + pair_off)
        pair_mm.append(ev["pair_mm"])
        splits.extend([(s + off, e + off) for s, e in bamio.iterate_splits(len(seq), window)])
        off += len(seq)
        pair_off += len(ev["pair_mm"])
    cat = np.concatenate
    batch = dict(scaffold_names=np.array(names), scaffold_len=np.array(lens, np.int32),
                 scaffold_off=np.array(offs, np.int32), ref_codes=cat(ref_codes),
                 ref_pos=cat(ev_pos).astype(np.int32), base=cat(ev_base), qual=cat(ev_qual),
                 read_id=cat(ev_rid).astype(np.int32), pair_mm=cat(pair_mm).astype(np.int32),
                 splits=np.array(splits, np.int32))
    name2off = dict(zip(names, offs))
    snp = pd.read_csv(os.path.join(raw, "raw_snp_table.csv.gz"))
    ld = pd.read_csv(os.path.join(raw, "raw_linkage_table.csv.gz"))
    soff = snp["scaffold"].map(name2off).values
    loff = ld["scaffold"].map(name2off).values
    exp = dict(
        snv_pos=(snp["position"].values + soff).astype(np.int32), snv_mm=snp["mm"].values.astype(np.int32),
        snv_cnt=snp[["A", "C", "T", "G"]].values.astype(np.int32),
        snv_con=snp["con_base"].map(B).values.astype(np.uint8), snv_var=snp["var_base"].map(B).values.astype(np.uint8),
        snv_allele_count=snp["allele_count"].values.astype(np.uint8),
        snv_cls=snp["class"].map(CLS).values.astype(np.uint8), snv_cryptic=snp["cryptic"].values.astype(np.uint8),
        ld_pos_a=(ld["position_A"].values + loff).astype(np.int32),
        ld_pos_b=(ld["position_B"].values + loff).astype(np.int32), ld_mm=ld["mm"].values.astype(np.int32),
        ld_counts=ld[["countAB", "countAb", "countaB", "countab"]].values.astype(np.int32),
        ld_alleles=np.stack([ld[c].map(B).values for c in ("allele_A", "allele_a", "allele_B", "allele_b")], 1).astype(np.uint8),
        ld_r2=ld["r2"].values.astype(np.float64), ld_d_prime=ld["d_prime"].values.astype(np.float64))
    # merge-stage summary: the reference's stored cumulative_scaffold_table, non-random columns
    from oracle.summary import COLUMNS
    cst = pd.read_csv(os.path.join(raw, "cumulative_scaffold_table.csv.gz"))
    name2idx = {n: i for i, n in enumerate(names)}
    cst = cst.assign(sidx=cst["scaffold"].map(name2idx)).sort_values(["sidx", "mm"])
    exp["sum_scaffold"] = cst["sidx"].values.astype(np.int32)
    exp["sum_columns"] = np.array(COLUMNS)
    exp["sum_values"] = cst[COLUMNS].values.astype(np.float64)
    return batch, exp


def make_subset_bam(which="G1", n_scaffolds=6):
    """A small coordinate-sorted BAM + its sR2M for the host-packer test: scaffolds whose mates hit htslib's
    overlap-walker quirk plus the most SNV-rich ones; ALL reads of those scaffolds are kept
    (also the ones the read filter dropped), re-numbered to a compact reference list."""
    import json
    refs, by_tid, seqs, rdic, raw = load_set(which)
    snp = pd.read_csv(os.path.join(raw, "raw_snp_table.csv.gz"))
    rich = snp.groupby("scaffold").size().sort_values(ascending=False).index.tolist()
    want = ["N5_271_010G1_scaffold_62", "N5_271_010G1_scaffold_649"] + rich
    names = []
    for nme in want:
        if nme in rdic and nme not in names:
            names.append(nme)
        if len(names) == n_scaffolds:
            break
    tids = sorted(t for t, (nme, _) in enumerate(refs) if nme in names)
    new_tid = {t: i for i, t in enumerate(tids)}
    new_refs = [refs[t] for t in tids]
    reads = []
    for t in tids:
        for r in by_tid.get(t, []):
            mt = new_tid.get(r.mtid, -1) if r.mtid >= 0 else -1
            reads.append(r._replace(tid=new_tid[t], mtid=mt))
    bamio.write_bam(os.path.join(HERE, "c1_%s_subset.bam" % which), new_refs, reads)
    with open(os.path.join(HERE, "c1_%s_subset_r2m.json" % which), "w") as fh:
        json.dump({refs[t][0]: rdic[refs[t][0]] for t in tids}, fh)
    with open(os.path.join(HERE, "c1_%s_subset_seqs.json" % which), "w") as fh:
        json.dump({refs[t][0]: seqs[refs[t][0]] for t in tids}, fh)
    print(which, "subset BAM:", len(new_refs), "scaffolds", len(reads), "reads")


def make_small_scaffold():
    """Tiny-scaffold case (the reference's test_profile_18 input: SmallScaffold.fa + its BAM, one 126 bp scaffold).
    No stored answer exists for it, so the expected tables come from the reference's OWN functions
    (oracle/ref_harness.py) on the emulated pileup; R2M = every name seen exactly twice with mm = summed NM
    (what get_paired_reads + paired_only yield before the ANI filter, inStrain/filter_reads.py:885-956,503-505)."""
    import json
    import shutil
    src_bam = os.path.join(TD_DIR, "SmallScaffold.fa.sorted.bam")
    refs, reads = bamio.read_bam(src_bam)
    seqs = bamio.read_fasta(os.path.join(TD_DIR, "SmallScaffold.fa"))
    name = refs[0][0]
    cnt, nm = {}, {}
    for r in reads:
        if r.tid == 0 and not (r.flag & 4):
            cnt[r.name] = cnt.get(r.name, 0) + 1
            nm[r.name] = nm.get(r.name, 0) + int(r.nm or 0)
    r2m = {k: int(nm[k]) for k, v in cnt.items() if v == 2 and nm[k] < 12}
    ev = pileup_emul.scaffold_events([r for r in reads if r.tid == 0], r2m)
    model = ref_harness.null_model(1e-6)
    out = ref_harness.run_split(ev, seqs[name], 0, len(seqs[name]) - 1, r2m, model, scaffold=name)
    shutil.copyfile(src_bam, os.path.join(HERE, "small_scaffold.bam"))
    snp = pd.DataFrame(out["snp"])
    ld = pd.DataFrame(out["ld"])
    with open(os.path.join(HERE, "small_scaffold.json"), "w") as fh:
        json.dump(dict(scaffold=name, seq=seqs[name], r2m=r2m,
                       snp=snp.drop(columns=["scaffold"]).to_dict("list") if len(snp) else {},
                       ld=ld[["position_A", "position_B", "mm", "countAB", "countAb", "countaB", "countab", "r2", "d_prime",
                              "allele_A", "allele_a", "allele_B", "allele_b"]].to_dict("list") if len(ld) else {},
                       covT={str(k): np.asarray(v).tolist() for k, v in out["covT"].items()}), fh)
    print("small scaffold:", name, len(seqs[name]), "bp,", len(r2m), "pairs,", len(snp), "snv rows,", len(ld), "ld rows")


def make_hd5_digests(which):
    """c1_<set>_hd5_digest.npz: per dataset of the reference's stored covT.hd5 / clonT.hd5 (read with the repo's own HDF5
    reader, instrain_b200/hd5.py -- h5py is not installable here) its length and a sha1 of the 2 x N array.  Pins the
    basewise coverage / clonality of the hot path at every position and the set of mm levels per scaffold."""
    from instrain_b200 import hd5
    sys.path.insert(0, os.path.dirname(HERE))
    from conftest import basewise_digest
    raw = load_set(which)[4]
    cov = hd5.read_hd5(os.path.join(raw, "covT.hd5"))
    clon = hd5.read_hd5(os.path.join(raw, "clonT.hd5"))
    assert set(cov) == set(clon)
    names = sorted(cov)
    np.savez_compressed(
        os.path.join(HERE, "c1_%s_hd5_digest.npz" % which), names=np.array(names),
        cov_n=np.array([cov[k].shape[1] for k in names], np.int32),
        cov_sha=np.stack([basewise_digest(cov[k][0], cov[k][1]) for k in names]),
        clon_n=np.array([clon[k].shape[1] for k in names], np.int32),
        clon_sha=np.stack([basewise_digest(clon[k][0], clon[k][1]) for k in names]))
    print(which, "hd5 digests:", len(names), "datasets,", sum(clon[k].shape[1] == 0 for k in names), "empty clonT levels")


TD_DIR = os.path.join(ref_harness.REFERENCE_ROOT, "test", "test_data")


if __name__ == "__main__":
    if sys.argv[1:] == ["hd5"]:                   # only the covT / clonT digests (fast)
        for which in ("G1", "G2"):
            make_hd5_digests(which)
        sys.exit(0)
    make_small_scaffold()
    make_subset_bam("G1")
    model = ref_harness.null_model(1e-6)
    lut, dflt = restate.lut_from_model(model)
    np.savez_compressed(os.path.join(HERE, "null_lut_fdr1e-06.npz"), lut=lut, default=np.int32(dflt))
    for which in ("G1", "G2"):
        batch, exp = make_batch(which)
        np.savez_compressed(os.path.join(HERE, "c1_%s_batch.npz" % which), **batch)
        np.savez_compressed(os.path.join(HERE, "c1_%s_expected.npz" % which), **exp)
        make_hd5_digests(which)
        print(which, "events", len(batch["ref_pos"]), "pairs", len(batch["pair_mm"]), "L", len(batch["ref_codes"]),
              "snv rows", len(exp["snv_pos"]), "ld rows", len(exp["ld_pos_a"]))


def make_hd5_structure_fixture():
    """tests/golden/hd5_structure_G1.json: hd5.describe_hd5 (address-free structure) of the reference's stored G1
    covT.hd5 / clonT.hd5, first 60 datasets -- what tests/test_hd5.py holds the native writer against where
    /root/reference is absent.  Run: python -c "import make_golden as m; m.make_hd5_structure_fixture()" in tests/golden."""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from instrain_b200 import hd5
    from test_hd5 import REF_RAW, _plain
    out = {}
    for name in ("covT", "clonT"):
        d = hd5.describe_hd5(os.path.join(REF_RAW % "G1", name + ".hd5"))
        d["datasets"] = {k: d["datasets"][k] for k in sorted(d["datasets"])[:60]}
        out[name] = _plain(d)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hd5_structure_G1.json"), "w"), indent=0, sort_keys=True)
