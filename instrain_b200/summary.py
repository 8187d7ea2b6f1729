"""cumulative_scaffold_table from K4's exact reductions (host glue; replaces the pandas side of make_coverage_table,
inStrain/profile/profile_utilities.py:425-506, and calc_snps, inStrain/profile/snv_utilities.py:249-272).

The rarefied columns (nucl_diversity_rarefied*, breadth_rarefied) come from a second K4 pass over clonTR.
"""
import math

import numpy as np
import pandas as pd

COLUMNS = ["scaffold", "length", "breadth", "coverage", "coverage_median", "coverage_std", "coverage_SEM", "nucl_diversity",
           "nucl_diversity_median", "nucl_diversity_rarefied", "nucl_diversity_rarefied_median", "breadth_minCov",
           "breadth_rarefied", "breadth_expected", "divergent_site_count", "SNS_count", "SNV_count",
           "consensus_divergent_sites", "population_divergent_sites", "conANI_reference", "popANI_reference", "mm"]


def snv_level_counts(snv, scaffold_off, n_scaffolds, M):
    """calc_snps for every (scaffold, mm): among rows with mm' <= mm keep the last row per position, then count.
    Returns int64[n_scaffolds, M, 5] = (SNS_count, SNV_count, divergent_site_count, con_snps, pop_snps)."""
    out = np.zeros((n_scaffolds, M + 1, 5), dtype=np.int64)
    if len(snv):
        snv = snv[np.lexsort((snv["mm"], snv["pos"]))]
        sidx = np.searchsorted(scaffold_off, snv["pos"], side="right") - 1
        nxt = np.full(len(snv), M, dtype=np.int64)                # level at which the next row of the position takes over
        same = snv["pos"][1:] == snv["pos"][:-1]
        nxt[:-1][same] = snv["mm"][1:][same]
        cat = np.stack([snv["allele_count"] == 1, snv["allele_count"] > 1, np.ones(len(snv), bool),
                        np.isin(snv["cls"], (2, 4, 5)), np.isin(snv["cls"], (2, 5))], 1).astype(np.int64)
        np.add.at(out, (sidx, snv["mm"].astype(np.int64)), cat)   # difference array over mm ...
        np.add.at(out, (sidx, nxt), -cat)
    return np.cumsum(out, axis=1)[:, :M]                          # ... integrated


def summary_table(rows, snv, scaffold_names, scaffold_off, M, rows_rarefied=None):
    """K4 rows (SUMMARY_DT[n_scaffolds*M]) + SNV rows -> DataFrame shaped like the reference's cumulative_scaffold_table.
    rows_rarefied = the K4 rows of the same scaffolds computed on clonTR instead of clonT: nucl_diversity_rarefied,
    nucl_diversity_rarefied_median, breadth_rarefied (make_coverage_table, profile_utilities.py:474-487); NaN / 0 without."""
    n_sc = len(scaffold_names)
    rows = rows.reshape(n_sc, M)
    rare = rows_rarefied.reshape(n_sc, M) if rows_rarefied is not None else None
    counts = snv_level_counts(snv, np.asarray(scaffold_off), n_sc, M)
    table = []
    for s in range(n_sc):
        for m in range(M):
            r = rows[s, m]
            if not r["present"]:
                continue
            L = int(r["length"])
            sm, sq = int(r["sum_cov"]), int(r["sum_cov2"])
            mean = sm / L
            var_num = L * sq - sm * sm                             # exact integer: L^2 * population variance
            std = math.sqrt(var_num) / L
            sem = math.sqrt(var_num / (L * (L - 1))) / math.sqrt(L) if L > 1 else float("nan")
            counted = int(r["counted"])
            sns, snvc, div, con, pop = (int(x) for x in counts[s, m])
            n_r = int(rare[s, m]["counted"]) if rare is not None else 0
            nd_r = 1 - float(rare[s, m]["sum_clon"]) / n_r if n_r else float("nan")
            nd_rm = 1 - (float(rare[s, m]["clon_med_lo"]) + float(rare[s, m]["clon_med_hi"])) / 2 if n_r else float("nan")
            table.append((
                scaffold_names[s], L, int(r["nonzero"]) / L, mean, int((int(r["cov_med_lo"]) + int(r["cov_med_hi"])) / 2.0),
                std, sem,
                1 - float(r["sum_clon"]) / counted if counted else float("nan"),
                1 - (float(r["clon_med_lo"]) + float(r["clon_med_hi"])) / 2 if counted else float("nan"),
                nd_r, nd_rm, counted / L, n_r / L, -math.exp(-0.883 * mean) + 1, div, sns, snvc, con, pop,
                (counted - con) / counted if counted else 0, (counted - pop) / counted if counted else 0, m))
    return pd.DataFrame(table, columns=COLUMNS)
