#!/bin/bash
# K1f split (pileup / general sites / site rows as three kernels): GPU tests, then A/B bench lines + ncu of the new kernels
tag=${1:-r2l}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
ISB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_transfer.py -m gpu -q > $out/${tag}_pytest_exp.log 2>&1
echo "pytest experimental exit $?" >> $out/${tag}_pytest_exp.log
tail -8 $out/${tag}_pytest_exp.log
SPECS="${SPECS:---no-e2e}" NCU=${NCU:-1} NCU_ARGS="--no-e2e" NCU_K="k1f_pileup|k2q_sites|k3f_site_rows" bash tools/gpu_ab2.sh $tag
