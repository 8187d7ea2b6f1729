"""numpy restatement of the merge-stage summary (test infrastructure only).

    make_coverage_table          inStrain/profile/profile_utilities.py:425-506
    mm_counts_to_counts_shrunk   inStrain/profile/profile_utilities.py:508-532
    get_basewise_clons           inStrain/profile/profile_utilities.py:534-546
    estimate_breadth             inStrain/profile/profile_utilities.py:548-556
    calc_snps                    inStrain/profile/snv_utilities.py:249-272

One row per (scaffold, mm level present in the scaffold's covT).  Random columns (nucl_diversity_rarefied*,
breadth_rarefied: clonTR, unseeded RNG) are not produced.  Pinned against the reference's stored
cumulative_scaffold_table (tests/test_oracle_golden.py).
"""
import numpy as np

COLUMNS = ["length", "breadth", "coverage", "coverage_median", "coverage_std", "coverage_SEM", "nucl_diversity",
           "nucl_diversity_median", "breadth_minCov", "breadth_expected", "divergent_site_count", "SNS_count",
           "SNV_count", "consensus_divergent_sites", "population_divergent_sites", "conANI_reference",
           "popANI_reference", "mm"]


def scaffold_summary(covT, clonT, nmask, snv, pos_lo):
    """covT int32[L,M], clonT float32[L,M] (NaN = unset), nmask uint64[L], snv = SNV rows of this scaffold (structured
    array with absolute `pos`), pos_lo = batch coordinate of the scaffold's first position.  Returns list of dict rows."""
    L, M = covT.shape
    rows = []
    cum = np.zeros(L, dtype=np.int64)
    last = np.full(L, np.nan, dtype=np.float64)             # clonality of the highest level <= mm at which it is set
    snv = snv[np.lexsort((snv["mm"], snv["pos"]))]
    for mm in range(M):
        present = bool((covT[:, mm] > 0).any()) or bool(((nmask >> np.uint64(mm)) & np.uint64(1)).any())
        cum += covT[:, mm]
        c = clonT[:, mm]
        setm = ~np.isnan(c)
        last[setm] = c[setm].astype(np.float64)
        if not present:
            continue
        covs = cum.astype(np.float64)
        clons = last[~np.isnan(last)]
        counted = len(clons)
        # calc_snps: rows with mm' <= mm, last (highest mm') row per position
        d = snv[snv["mm"] <= mm]
        if len(d):
            keep = np.ones(len(d), dtype=bool)
            keep[:-1] = d["pos"][1:] != d["pos"][:-1]
            d = d[keep]
        sns = int((d["allele_count"] == 1).sum())
        snvc = int((d["allele_count"] > 1).sum())
        con = int(np.isin(d["cls"], (2, 4, 5)).sum())          # SNS, con_SNV, pop_SNV
        pop = int(np.isin(d["cls"], (2, 5)).sum())             # SNS, pop_SNV
        cov_mean = float(np.mean(covs))
        rows.append(dict(
            length=L, breadth=np.count_nonzero(covs) / L, coverage=cov_mean, coverage_median=int(np.median(covs)),
            coverage_std=float(np.std(covs)), coverage_SEM=float(np.std(covs, ddof=1) / np.sqrt(L)) if L > 1 else np.nan,
            nucl_diversity=1 - float(np.mean(clons)) if counted else np.nan,
            nucl_diversity_median=1 - float(np.median(clons)) if counted else np.nan,
            breadth_minCov=counted / L, breadth_expected=float(-np.exp(-0.883 * cov_mean) + 1),
            divergent_site_count=len(d), SNS_count=sns, SNV_count=snvc, consensus_divergent_sites=con,
            population_divergent_sites=pop,
            conANI_reference=(counted - con) / counted if counted else 0,
            popANI_reference=(counted - pop) / counted if counted else 0, mm=mm))
    return rows
