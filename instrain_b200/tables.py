"""Row arrays from the kernels -> the reference's pandas tables (host glue, no compute).

Column names / order follow what the reference stores:
  raw_snp_table      generate_snp_table              inStrain/profile/snv_utilities.py:274-290 (+ position shift :181)
  raw_linkage_table  _calc_ld_single / _update_r2    inStrain/profile/linkage.py:138-252
  covT / clonT       shrink_basewise                 inStrain/profile/profile_utilities.py:337-350
The reference's resampled columns (r2_normalized, d_prime_normalized, clonTR) come from the device's counter-based
generator (seeded: reproducible; same distribution as the reference's unseeded np.random.choice).
"""
import numpy as np
import pandas as pd

from ._cabi import CLASS_NAMES

BASES = np.array(list("ACTG"))
SNV_COLUMNS = ["scaffold", "position", "ref_base", "A", "C", "T", "G", "con_base", "var_base", "mm", "allele_count",
               "class", "cryptic", "position_coverage"]
LD_COLUMNS = ["r2", "d_prime", "r2_normalized", "d_prime_normalized", "total", "countAB", "countAb", "countaB", "countab",
              "allele_A", "allele_a", "allele_B", "allele_b", "distance", "position_A", "position_B", "mm", "scaffold"]


def _locate(pos, scaffold_off):
    """batch coordinate -> (scaffold index, scaffold-relative position)."""
    idx = np.searchsorted(scaffold_off, pos, side="right") - 1
    return idx, pos - np.asarray(scaffold_off)[idx]


_REF_CHARS = np.array(list("ACTGN"), dtype=object)


def snv_ref_chars(rows, sidx, rel, names, seqs, ref_codes=None):
    """Reference character of every SNV row: from the batch's reference codes when given (one gather for the whole batch;
    only positions whose code is "not A/C/G/T" look their letter up in the sequence), else per scaffold from `seqs`."""
    if ref_codes is not None:
        codes = np.asarray(ref_codes)[rows["pos"].astype(np.int64)]
        ref_base = _REF_CHARS[np.minimum(codes, 4)]
        for k in np.nonzero(codes > 3)[0]:                               # rare: N or another IUPAC letter
            ref_base[k] = seqs[names[sidx[k]]][rel[k]]
        return ref_base
    ref_base = np.empty(len(rows), dtype=object)
    order = np.argsort(sidx, kind="stable")
    bounds = np.searchsorted(sidx[order], np.arange(len(names) + 1))
    for i in np.nonzero(np.diff(bounds))[0]:
        m = order[bounds[i]:bounds[i + 1]]
        seq = np.frombuffer(seqs[names[i]].encode(), dtype="S1")
        ref_base[m] = seq[rel[m]].astype(str)
    return ref_base


def snv_frame(rows, scaffold_of_row, rel, ref_base):
    """isb_snv_row[] + (scaffold name, scaffold-relative position, reference character) per row -> raw_snp_table."""
    cnt = rows["cnt"].astype(np.int64)
    return pd.DataFrame({
        "scaffold": scaffold_of_row, "position": np.asarray(rel, dtype=np.int64), "ref_base": ref_base,
        "A": cnt[:, 0], "C": cnt[:, 1], "T": cnt[:, 2], "G": cnt[:, 3],
        "con_base": BASES[rows["con"]], "var_base": BASES[rows["var"]], "mm": rows["mm"].astype(np.int64),
        "allele_count": rows["allele_count"].astype(np.int64), "class": np.array(CLASS_NAMES, dtype=object)[rows["cls"]],
        "cryptic": rows["cryptic"].astype(bool), "position_coverage": cnt.sum(1)}, columns=SNV_COLUMNS)


def snv_table(rows, scaffold_names, scaffold_off, seqs, ref_codes=None):
    """isb_snv_row[] (batch coordinates) -> raw_snp_table DataFrame (sorted by scaffold order, position, mm)."""
    rows = rows[np.lexsort((rows["mm"], rows["pos"]))]
    names = np.asarray(scaffold_names, dtype=object)
    sidx, rel = _locate(rows["pos"].astype(np.int64), scaffold_off)
    return snv_frame(rows, names[sidx], rel, snv_ref_chars(rows, sidx, rel, names, seqs, ref_codes))


def cumulative_snv_table(raw):
    """raw_snp_table -> cumulative_snv_table: _parse_Sdb (profile_utilities.py:600-616) adds var_freq / con_freq / ref_freq
    = count of that base / position_coverage (ref_freq NaN when the reference base is not A/C/T/G).  The reference applies
    it in place, so its stored raw_snp_table carries the same three columns."""
    out = raw.copy()
    if len(out) == 0:
        return out
    cnt = out[["A", "C", "T", "G"]].values.astype(np.float64)
    cov = out["position_coverage"].values.astype(np.float64)
    lut = {b: i for i, b in enumerate("ACTG")}
    rows = np.arange(len(out))
    for col, src in (("var_freq", "var_base"), ("con_freq", "con_base"), ("ref_freq", "ref_base")):
        idx = np.array([lut.get(b, -1) for b in out[src]], dtype=np.int64)
        val = cnt[rows, np.maximum(idx, 0)] / cov
        out[col] = np.where(idx >= 0, val, np.nan)
    return out


def linkage_frame(rows, scaffold_of_row, rel_a, rel_b):
    c = [rows[k].astype(np.int64) for k in ("c_AB", "c_Ab", "c_aB", "c_ab")]
    return pd.DataFrame({
        "r2": rows["r2"], "d_prime": rows["d_prime"], "r2_normalized": rows["r2_normalized"],
        "d_prime_normalized": rows["d_prime_normalized"],
        "total": c[0] + c[1] + c[2] + c[3], "countAB": c[0], "countAb": c[1], "countaB": c[2], "countab": c[3],
        "allele_A": BASES[rows["allele_A"]], "allele_a": BASES[rows["allele_a"]], "allele_B": BASES[rows["allele_B"]],
        "allele_b": BASES[rows["allele_b"]], "distance": rel_b - rel_a, "position_A": rel_a, "position_B": rel_b,
        "mm": rows["mm"].astype(np.int64), "scaffold": scaffold_of_row}, columns=LD_COLUMNS)


def linkage_table(rows, scaffold_names, scaffold_off):
    rows = rows[np.lexsort((rows["mm"], rows["pos_b"], rows["pos_a"]))]
    names = np.asarray(scaffold_names, dtype=object)
    sidx, rel_a = _locate(rows["pos_a"].astype(np.int64), scaffold_off)
    rel_b = rows["pos_b"].astype(np.int64) - np.asarray(scaffold_off)[sidx]
    return linkage_frame(rows, names[sidx], rel_a, rel_b)


def present_levels(covT, nmask):
    """mm levels that are keys of a scaffold's covT / clonT dicts in the reference: every level that is a key of some
    column's MMcounts (update_covT profile_utilities.py:288-295 and update_snp_table snv_utilities.py:87-96 create
    covT[mm] / clonT[mm] on first sight of mm) -- a level with coverage somewhere, or one made present only by a
    non-ACGT read base (the defaultdict side effect, profile_utilities.py:280-281 -> nmask)."""
    M = covT.shape[1]
    bits = int(np.bitwise_or.reduce(nmask)) if len(nmask) else 0
    return [mm for mm in range(M) if (bits >> mm) & 1 or (covT[:, mm] > 0).any()]


def basewise(dense, kind, levels=None):
    """Dense [L, M] array of one scaffold -> {mm: pd.Series} like shrink_basewise: coverage keeps values > 0 (int32),
    clonality drops NaN (float32).  `levels` = the dict keys (present_levels); a present level with no retained position
    is an EMPTY Series, as in the reference's stored covT / clonT (e.g. clonT of a level that never reaches min_cov).
    Without `levels`, a level appears iff it has at least one retained position."""
    out = {}
    for mm in (range(dense.shape[1]) if levels is None else levels):
        col = dense[:, mm]
        keep = np.nonzero(col > 0)[0] if kind == "coverage" else np.nonzero(~np.isnan(col) & (col > 0))[0]
        if len(keep) or levels is not None:
            out[mm] = pd.Series(col[keep].astype("int32" if kind == "coverage" else "float32"), index=keep)
    return out
