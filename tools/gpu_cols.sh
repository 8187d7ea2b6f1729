#!/bin/bash
# Column-word path: its GPU parity tests, A/B bench runs (20 scaffolds) of build variants, then the ncu passes.
#   gpurun --timeout 1500 -- 'bash tools/gpu_cols.sh r1s'
tag=${1:-cols}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 700 python -m pytest tests/test_gpu_cols.py -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
bash tools/gpu_ab.sh $tag "ISB_LIB_PATH=instrain_b200/lib/libisb_rows8.so -- " "ISB_LIB_PATH=instrain_b200/lib/libisb_rows8b5.so -- " \
    "ISB_LIB_PATH=instrain_b200/lib/libisb_minb8.so -- " "ISB_LIB_PATH=instrain_b200/lib/libisb_rows8.so ISB_K1C_FUSE=0 -- "
SKIP_TESTS=1 SKIP_BENCH=1 bash tools/gpu_round.sh $tag
