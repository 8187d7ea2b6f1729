#!/bin/bash
# fused-path tests + A/B bench lines by environment (SPECS) + optional ncu
tag=${1:-r2n}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
if [ "${TESTS:-1}" = "1" ]; then
timeout 1200 python -m pytest tests/test_gpu_reads.py tests/test_gpu_profile_bam.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
fi
NCU=${NCU:-0} NCU_ARGS="--no-e2e" NCU_K="${NCU_K:-k1f_pileup|k2q_sites|k3f_site_rows|k3_enum_pairs_tiles|k3_pair_stats_dev}" NCU_C=${NCU_C:-5} bash tools/gpu_ab2.sh $tag
