#!/bin/bash
# ncu --set full of the read-major kernels on the 10-scaffold bench: bash tools/gpu_ncu.sh <tag>
tag=${1:-ncu}
out=gpurun_out
mkdir -p $out
SMALL="--scaffolds 10 --steps 1 --warmup 1 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k1r_pileup|k2_call|k3r_site_rows|k3r_site_cand|k3_enum_pairs|k3_pair_stats|k3_suffix' -s 12 -c 7 \
    -f -o $out/${tag}_full python bench.py $SMALL > $out/${tag}_full.log 2>&1
echo "ncu full exit $?"
