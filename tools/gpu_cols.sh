#!/bin/bash
# Column-word path: its GPU parity tests, then A/B bench runs (20 scaffolds) of build variants / switches.
#   gpurun --timeout 1500 -- 'bash tools/gpu_cols.sh r1t'
tag=${1:-cols}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 700 python -m pytest tests/test_gpu_cols.py -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
bash tools/gpu_ab.sh $tag "ISB_K1C_FUSE=0 -- " "ISB_LIB_PATH=instrain_b200/lib/libisb_rows4.so -- " \
    "ISB_LIB_PATH=instrain_b200/lib/libisb_rows1.so -- " "ISB_LIB_PATH=instrain_b200/lib/libisb_mb8.so -- " \
    "ISB_LIB_PATH=instrain_b200/lib/libisb_mb8.so ISB_K1C_FUSE=0 -- "
