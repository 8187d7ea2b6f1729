#!/bin/bash
# A/B of register budgets (variant libraries) + fused-path tests + launch list of the default step
tag=${1:-r2m}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_reads.py tests/test_gpu_zz_transfer.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
L=$PWD/instrain_b200/lib
SPECS="--no-e2e;ISB_LIB_PATH=$L/libisbv_k1f8.so --no-e2e;ISB_LIB_PATH=$L/libisbv_k3f5.so --no-e2e;ISB_LIB_PATH=$L/libisbv_k3f6.so --no-e2e" NCU=1 NCU_ARGS="--no-e2e" NCU_K="k1f_pileup|k2q_sites|k3f_site_rows|k3_enum_pairs_tiles|k3_pair_stats_dev" NCU_C=5 bash tools/gpu_ab2.sh $tag
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --scaffolds 10 --steps 2 --warmup 1 --also-layouts 0 --no-cpu-baseline --no-e2e --sustain-s 0 --from-bam-scaffolds 0 > $out/${tag}_launches.log 2>&1
echo "launch list exit $?"
