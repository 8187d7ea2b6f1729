#!/bin/bash
# Full GPU parity suite, then A/B bench runs (20 scaffolds) of build variants / switches.
#   gpurun --timeout 1500 -- 'bash tools/gpu_cols.sh r1w'
tag=${1:-cols}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
bash tools/gpu_ab.sh $tag "ISB_LIB_PATH=instrain_b200/lib/libisbv_mb8.so -- " "ISB_LIB_PATH=instrain_b200/lib/libisbv_enum2.so -- "
