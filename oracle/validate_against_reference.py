"""Pin the pileup emulation against the reference's golden tables (build container only).

    python -m oracle.validate_against_reference [G1|G2]

Feeds oracle/pileup_emul.py columns to the reference's own functions (oracle/ref_harness.py) for
every scaffold/split of the bundled BAM and compares with the stored
`...forRC.IS/raw_data/raw_snp_table.csv.gz` and `raw_linkage_table.csv.gz`
(produced by inStrain v1.7.0 / pysam 0.16.0.1).  Random columns (r2_normalized,
d_prime_normalized) are excluded, exactly as the reference's own tests do
(test/tests/test_profile.py:896-900,922-927).
"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

from . import bamio, pileup_emul, ref_harness

TD = os.path.join(ref_harness.REFERENCE_ROOT, "test", "test_data")
FASTA = os.path.join(TD, "N5_271_010G1_scaffold_min1000.fa")
SETS = {
    "G1": ("N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.sorted.bam",
           "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G1.forRC.IS"),
    "G2": ("N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G2.sorted.bam",
           "N5_271_010G1_scaffold_min1000.fa-vs-N5_271_010G2.forRC.IS"),
}
SNP_COLS = ["scaffold", "position", "mm", "ref_base", "A", "C", "T", "G", "con_base", "var_base",
            "allele_count", "class", "cryptic"]
LD_INT = ["scaffold", "position_A", "position_B", "mm", "distance", "total", "countAB", "countAb", "countaB",
          "countab", "allele_A", "allele_a", "allele_B", "allele_b"]


def load_set(which):
    bam, isdir = SETS[which]
    refs, reads = bamio.read_bam(os.path.join(TD, bam))
    seqs = bamio.read_fasta(FASTA)
    rdic = json.load(open(os.path.join(TD, isdir, "raw_data", "Rdic.json")))
    by_tid = {}
    for r in reads:
        if r.tid >= 0:
            by_tid.setdefault(r.tid, []).append(r)
    return refs, by_tid, seqs, rdic, os.path.join(TD, isdir, "raw_data")


def run_reference_on_set(which, window=10000, scaffolds=None):
    refs, by_tid, seqs, rdic, _ = load_set(which)
    model = ref_harness.null_model(1e-6)
    snp, ld = [], []
    for tid, (name, length) in enumerate(refs):
        if name not in rdic or (scaffolds is not None and name not in scaffolds):
            continue
        r2m = rdic[name]
        ev = pileup_emul.scaffold_events(by_tid.get(tid, []), r2m)
        for (start, end) in bamio.iterate_splits(len(seqs[name]), window):
            out = ref_harness.run_split(ev, seqs[name], start, end, r2m, model, scaffold=name)
            snp.extend(out["snp"])
            ld.extend(out["ld"])
    return pd.DataFrame(snp), pd.DataFrame(ld)


def compare(which):
    t0 = time.time()
    snp, ld = run_reference_on_set(which)
    raw = load_set(which)[4]
    g_snp = pd.read_csv(os.path.join(raw, "raw_snp_table.csv.gz"))
    g_ld = pd.read_csv(os.path.join(raw, "raw_linkage_table.csv.gz"))
    key = ["scaffold", "position", "mm"]
    a = snp[SNP_COLS].sort_values(key).reset_index(drop=True)
    b = g_snp[SNP_COLS].sort_values(key).reset_index(drop=True)
    ok_snp = len(a) == len(b) and all((a[c].values == b[c].values).all() for c in SNP_COLS)
    key = ["scaffold", "position_A", "position_B", "mm"]
    a = ld.sort_values(key).reset_index(drop=True)
    b = g_ld.sort_values(key).reset_index(drop=True)
    ok_ld = len(a) == len(b) and all((a[c].values == b[c].values).all() for c in LD_INT)
    if ok_ld:
        for c in ("r2", "d_prime"):
            ok_ld &= bool(np.allclose(a[c].values.astype(float), b[c].values.astype(float), rtol=0, atol=1e-9,
                                      equal_nan=True))
    print("%s: raw_snp_table %d/%d rows %s ; raw_linkage_table %d/%d rows %s ; %.1fs" % (
        which, len(snp), len(g_snp), "IDENTICAL" if ok_snp else "MISMATCH",
        len(ld), len(g_ld), "IDENTICAL" if ok_ld else "MISMATCH", time.time() - t0))
    return ok_snp and ok_ld


if __name__ == "__main__":
    which = sys.argv[1:] or ["G1", "G2"]
    ok = all([compare(w) for w in which])
    sys.exit(0 if ok else 1)
