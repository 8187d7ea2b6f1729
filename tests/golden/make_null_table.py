"""Build instrain_b200/data/null_model_probs.npz from the reference's bundled error-model table
(/root/reference/inStrain/helper_files/NullModel.txt: coverage x P(>= k alternate bases by sequencing error), k = 1..18,
made by the reference's helper_scripts/calculate_null.py).  The packaged form holds the parsed numbers only (float64, the
values `float(field)` gives), so instrain_b200.null_model.load_lut can derive the LUT for ANY fdr the way
generate_snp_model does (inStrain/profile/snv_utilities.py:14-38) without the text file.  Build container only.

    python tests/golden/make_null_table.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.environ.get("INSTRAIN_REFERENCE", "/root/reference") + "/inStrain/helper_files/NullModel.txt"


def main():
    cov, rows = [], []
    with open(SRC) as fh:
        for line in fh:
            if "coverage" in line:
                continue
            f = line.split()
            if not f:
                continue
            cov.append(int(f[0]))
            rows.append([float(x) for x in f[1:]])
    width = max(len(r) for r in rows)
    probs = np.full((len(rows), width), np.nan)
    for i, r in enumerate(rows):
        probs[i, :len(r)] = r
    out = os.path.join(ROOT, "instrain_b200", "data", "null_model_probs.npz")
    np.savez_compressed(out, coverage=np.asarray(cov, np.int32), probs=probs)
    print(out, probs.shape, os.path.getsize(out))


if __name__ == "__main__":
    sys.exit(main())
