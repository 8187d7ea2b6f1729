"""Oracle = TEST INFRASTRUCTURE ONLY.

CPU restatement of the inStrain `profile` hot path (pileup -> SNV call -> linkage).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import anything from this package.  The product (`instrain_b200/`) never does:
it fails loudly when the CUDA library is missing.

Parity pin: see oracle/README.md -- the restatement is pinned against the reference's own
golden tables (raw_snp_table / raw_linkage_table of the two `forRC.IS` profile directories)
and against the reference's own functions driven through `oracle/ref_harness.py`.
"""
