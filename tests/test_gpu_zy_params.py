"""CUDA path against the REFERENCE's own outputs for non-default settings (tests/golden/c1_G1_params_*.npz, produced by
tests/golden/make_param_goldens.py from the reference's functions): min_cov / min_freq / min_snp / fdr away from the
defaults, and set-mode R2M (--skip_mm_profiling), in all three input layouts.  SNV rows, linkage rows and the dense covT /
clonT of 30 scaffolds.  (The same fixtures pin the oracle in tests/test_oracle_golden.py.)"""
import pytest

from test_oracle_golden import _param_case, check_against_param_golden, check_ns_case, load_ns_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("name", ["params_P1", "params_P2", "params_P3"])
def test_cuda_reproduces_reference_with_other_settings(name):
    from instrain_b200 import cols, reads
    from instrain_b200.engine import Engine
    z, batch, off, ln, keep, sha = _param_case(name)
    M = int(batch["pair_mm"].max()) + 1
    rd = reads.events_to_reads(batch)
    cd = cols.reads_to_cols(rd, len(batch["ref_codes"]))
    eng = Engine(0, z["lut"], int(z["lut_default"]))             # the fixture's own null model (fdr 1e-3 for P2)
    try:
        kw = dict(M=M, min_cov=int(z["min_cov"]), min_freq=float(z["min_freq"]), min_snp=int(z["min_snp"]),
                  want=("covT", "clonT", "snv", "ld"))
        for layout in (dict(cols=cd), dict(reads=rd), {}):
            out = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], **kw, **layout)
            check_against_param_golden(out, z, batch, off, ln, keep, sha)
    finally:
        eng.close()


def test_cuda_reproduces_reference_on_n_reference_case(null_lut):
    """scaffold_963_Ns (an N in the reference, ~185x coverage): the three layouts against the reference's own outputs."""
    from instrain_b200 import cols, reads
    from instrain_b200.engine import Engine
    z, batch = load_ns_case()
    M = int(batch["pair_mm"].max()) + 1
    rd = reads.events_to_reads(batch)
    cd = cols.reads_to_cols(rd, len(batch["ref_codes"]))
    eng = Engine(0, null_lut[0], null_lut[1])
    try:
        for layout in (dict(cols=cd), dict(reads=rd), {}):
            out = eng.profile_batch(batch, batch["ref_codes"], batch["splits"], M=M, want=("covT", "clonT", "snv", "ld"), **layout)
            check_ns_case(out, z, batch)
    finally:
        eng.close()
