#!/bin/bash
# Final evidence of a round: full GPU test suite, smoke, the default bench line, launch list + ncu --set full of one step.
tag=${1:-final}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
bash tools/gpu_r2d.sh $tag
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --scaffolds 10 --steps 2 --warmup 1 --also-layouts 0 --no-cpu-baseline --no-e2e --sustain-s 0 --from-bam-scaffolds 0 > $out/${tag}_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k1f_pileup|k2q_sites|k3f_site_rows|k3_enum_pairs_tiles|k3_pair_stats_dev" -s 5 -c 5 \
    -f -o $out/${tag}_full python bench.py --scaffolds 10 --steps 1 --warmup 1 --also-layouts 0 --no-cpu-baseline --no-e2e --sustain-s 0 --from-bam-scaffolds 0 > $out/${tag}_full.log 2>&1
echo "ncu full exit $?"
