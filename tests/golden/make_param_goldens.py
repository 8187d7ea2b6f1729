"""Known-answer vectors of the REFERENCE ITSELF for non-default parameters (build container only; needs /root/reference).

    python tests/golden/make_param_goldens.py

The stored goldens of the reference (c1_G{1,2}_expected.npz) were produced with inStrain's default thresholds.  This
script runs the reference's OWN functions (process_bam_sites / update_snp_table / calc_mm_SNV_linkage_network /
calculate_ld, imported from /root/reference through oracle/ref_harness.py, fed with the htslib-faithful columns of
oracle/pileup_emul.py -- the combination that reproduces the stored goldens row for row) on the 30 best-covered
scaffolds of the bundled G1 BAM with OTHER settings, and stores what they return:

  params_P1  min_cov 2,  min_freq 0.10, min_snp 5,  fdr 1e-6
  params_P2  min_cov 10, min_freq 0.02, min_snp 40, fdr 1e-3  (its own null-model LUT is stored with it)
  params_P3  defaults, R2M given as a SET of read names (--skip_mm_profiling: every read counts at mm 0)

-> tests/golden/c1_G1_<name>.npz: SNV rows, linkage rows (batch coordinates of c1_G1_batch.npz), and sha1 digests of the
dense covT (int32 [L][M]) / clonT (float32 [L][M], NaN = unset) of every chosen scaffold.  tests/test_oracle_golden.py
checks the oracle against them, which pins its parameter handling on the reference, not only on the defaults.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bamio, pileup_emul, ref_harness, restate  # noqa: E402
from oracle.validate_against_reference import load_set  # noqa: E402

CLS = {n: i for i, n in enumerate(restate.CLASS_NAMES)}
B = {b: i for i, b in enumerate("ACTG")}
PARAMS = {
    "params_P1": dict(min_cov=2, min_freq=0.10, min_snp=5, fdr=1e-6, set_mode=False),
    "params_P2": dict(min_cov=10, min_freq=0.02, min_snp=40, fdr=1e-3, set_mode=False),
    "params_P3": dict(min_cov=5, min_freq=0.05, min_snp=20, fdr=1e-6, set_mode=True),
}
N_SCAFFOLDS = 30


def digest(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def main():
    refs, by_tid, seqs, rdic, _ = load_set("G1")
    batch = dict(np.load(os.path.join(HERE, "c1_G1_batch.npz")))
    name2off = dict(zip(batch["scaffold_names"], batch["scaffold_off"].astype(np.int64)))
    chosen = sorted(rdic, key=lambda s: -len(rdic[s]))[:N_SCAFFOLDS]
    tid_of = {name: tid for tid, (name, _) in enumerate(refs)}
    for pname, prm in PARAMS.items():
        model = ref_harness.null_model(prm["fdr"])
        lut, dflt = restate.lut_from_model(model)
        snp, ld, names, cov_sha, clon_sha, Ms = [], [], [], [], [], []
        for name in chosen:
            r2m = rdic[name]
            ev = pileup_emul.scaffold_events(by_tid.get(tid_of[name], []), r2m)
            r2m_arg = set(r2m) if prm["set_mode"] else r2m
            seq, off = seqs[name], int(name2off[name])
            M = 1 if prm["set_mode"] else (max(r2m.values()) + 1)
            cov = np.zeros((len(seq), M), dtype=np.int32)
            clon = np.full((len(seq), M), np.nan, dtype=np.float32)
            for (start, end) in bamio.iterate_splits(len(seq), 10000):
                out = ref_harness.run_split(ev, seq, start, end, r2m_arg, model, scaffold=name, min_cov=prm["min_cov"],
                                            min_freq=prm["min_freq"], min_snp=prm["min_snp"])
                for r in out["snp"]:
                    r["position"] += off
                    snp.append(r)
                for r in out["ld"]:
                    r["position_A"] += off
                    r["position_B"] += off
                    ld.append(r)
                for mm, arr in out["covT"].items():
                    cov[start:end + 1, mm] = arr
                for mm, arr in out["clonT"].items():
                    clon[start:end + 1, mm] = arr
            names.append(name)
            cov_sha.append(digest(cov))
            clon_sha.append(digest(clon))
            Ms.append(M)
        exp = dict(
            scaffolds=np.array(names), M=np.array(Ms, np.int32), cov_sha=np.stack(cov_sha), clon_sha=np.stack(clon_sha),
            lut=np.asarray(lut, np.int32), lut_default=np.int32(dflt),
            min_cov=np.int32(prm["min_cov"]), min_freq=np.float64(prm["min_freq"]), min_snp=np.int32(prm["min_snp"]),
            set_mode=np.bool_(prm["set_mode"]),
            snv_pos=np.array([r["position"] for r in snp], np.int32), snv_mm=np.array([r["mm"] for r in snp], np.int32),
            snv_cnt=np.array([[r["A"], r["C"], r["T"], r["G"]] for r in snp], np.int32).reshape(-1, 4),
            snv_con=np.array([B[r["con_base"]] for r in snp], np.uint8), snv_var=np.array([B[r["var_base"]] for r in snp], np.uint8),
            snv_allele_count=np.array([r["allele_count"] for r in snp], np.uint8),
            snv_cls=np.array([CLS[r["class"]] for r in snp], np.uint8), snv_cryptic=np.array([r["cryptic"] for r in snp], np.uint8),
            ld_pos_a=np.array([r["position_A"] for r in ld], np.int32), ld_pos_b=np.array([r["position_B"] for r in ld], np.int32),
            ld_mm=np.array([r["mm"] for r in ld], np.int32),
            ld_counts=np.array([[r["countAB"], r["countAb"], r["countaB"], r["countab"]] for r in ld], np.int32).reshape(-1, 4),
            ld_alleles=np.array([[B[r[c]] for c in ("allele_A", "allele_a", "allele_B", "allele_b")] for r in ld], np.uint8).reshape(-1, 4),
            ld_r2=np.array([r["r2"] for r in ld], np.float64), ld_d_prime=np.array([r["d_prime"] for r in ld], np.float64))
        np.savez_compressed(os.path.join(HERE, "c1_G1_%s.npz" % pname), **exp)
        print(pname, "scaffolds", len(names), "snv rows", len(snp), "ld rows", len(ld))


def rows_to_arrays(snp, ld):
    return dict(
        snv_pos=np.array([r["position"] for r in snp], np.int32), snv_mm=np.array([r["mm"] for r in snp], np.int32),
        snv_cnt=np.array([[r["A"], r["C"], r["T"], r["G"]] for r in snp], np.int32).reshape(-1, 4),
        snv_con=np.array([B[r["con_base"]] for r in snp], np.uint8), snv_var=np.array([B[r["var_base"]] for r in snp], np.uint8),
        snv_allele_count=np.array([r["allele_count"] for r in snp], np.uint8),
        snv_cls=np.array([CLS[r["class"]] for r in snp], np.uint8), snv_cryptic=np.array([r["cryptic"] for r in snp], np.uint8),
        ld_pos_a=np.array([r["position_A"] for r in ld], np.int32), ld_pos_b=np.array([r["position_B"] for r in ld], np.int32),
        ld_mm=np.array([r["mm"] for r in ld], np.int32),
        ld_counts=np.array([[r["countAB"], r["countAb"], r["countaB"], r["countab"]] for r in ld], np.int32).reshape(-1, 4),
        ld_alleles=np.array([[B[r[c]] for c in ("allele_A", "allele_a", "allele_B", "allele_b")] for r in ld], np.uint8).reshape(-1, 4),
        ld_r2=np.array([r["r2"] for r in ld], np.float64), ld_d_prime=np.array([r["d_prime"] for r in ld], np.float64))


def make_ns_case():
    """The reference's edge-case input `N5_271_010G1_scaffold_963_Ns` (a reference with an N, ~185x coverage, 26395 reads of
    which 970 pairs pass the default read filter; the reference's tests only check that it runs): INPUT batch (events of
    the pairs the pinned read filter keeps) + what the reference's functions return for it with default settings."""
    from oracle import read_filter
    td = os.path.join(ref_harness.REFERENCE_ROOT, "test", "test_data")
    refs, reads = bamio.read_bam(os.path.join(td, "N5_271_010G1_scaffold_963_Ns.fasta.sorted.bam"))
    seqs = bamio.read_fasta(os.path.join(td, "N5_271_010G1_scaffold_963_Ns.fasta"))
    name = refs[0][0]
    rs = [r for r in reads if r.tid == 0]
    kept, _, _ = read_filter.filter_pairs({name: read_filter.pair2info(rs)})
    r2m = kept[name]
    ev = pileup_emul.scaffold_events(rs, r2m)
    seq = seqs[name]
    model = ref_harness.null_model(1e-6)
    out = ref_harness.run_split(ev, seq, 0, len(seq) - 1, r2m, model, scaffold=name)
    M = max(r2m.values()) + 1
    cov = np.zeros((len(seq), M), dtype=np.int32)
    clon = np.full((len(seq), M), np.nan, dtype=np.float32)
    for mm, arr in out["covT"].items():
        cov[:, mm] = arr
    for mm, arr in out["clonT"].items():
        clon[:, mm] = arr
    sev = restate.sort_events(ev)
    exp = rows_to_arrays(out["snp"], out["ld"])
    exp.update(ref_codes=restate.encode_ref(seq), ref_pos=sev["ref_pos"].astype(np.int32), base=sev["base"], qual=sev["qual"],
               read_id=sev["read_id"].astype(np.int32), pair_mm=sev["pair_mm"].astype(np.int32),
               splits=np.array([[0, len(seq) - 1]], np.int32), covT=cov, clonT=clon)
    np.savez_compressed(os.path.join(HERE, "c963_Ns.npz"), **exp)
    print("c963_Ns: events", len(sev["ref_pos"]), "pairs", len(sev["pair_mm"]), "N in reference", int((exp["ref_codes"] > 3).sum()),
          "snv rows", len(out["snp"]), "ld rows", len(out["ld"]))


if __name__ == "__main__":
    if sys.argv[1:] != ["ns"]:
        main()
    make_ns_case()
