#!/bin/bash
# First GPU call of round 2: the paths written after the last GPU session of round 1 (gated tests incl. the column-word
# chunk pipeline, the parameter / N-reference pins on the CUDA path), then the pipeline A/B at 20 x 1 Mb and at the full
# configuration.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r2a'
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
ISB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_zy_params.py tests/test_gpu_zz_transfer.py -m gpu -q \
    > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -8 $out/${tag}_pytest.log
bash tools/gpu_ab.sh $tag " -- --pipeline" " -- --scaffolds 100 --steps 10" " -- --scaffolds 100 --steps 10 --pipeline"
