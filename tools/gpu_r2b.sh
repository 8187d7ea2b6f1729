#!/bin/bash
# Round 2: parity of the read-major kernels, A/B bench lines, ncu --set full of K1f (fused and plain) with source counters.
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_reads.py -m gpu -x -q -k "fused or synthetic_parity or pileup_reads_stage or many_mm" > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
ARGS="--scaffolds 20 --steps 5 --warmup 3 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
i=0
IFS=';' read -ra SPECS <<< "--layout reads;ISB_K1F=0 --layout reads${EXTRA_SPECS:+;$EXTRA_SPECS}"
for spec in "${SPECS[@]}"; do
  envs=""; extra=""
  for tok in $spec; do case $tok in *=*) envs="$envs $tok";; *) extra="$extra $tok";; esac; done
  env $envs timeout 600 python bench.py $ARGS $extra > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_$i.json"))
    print("[$spec]", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "e2e %.3e" % d["e2e"]["value"])
except Exception as ex:
    print("[$spec] failed", ex)
PY
  tail -3 $out/${tag}_$i.err
  i=$((i+1))
done
if [ "${SKIP_NCU:-0}" != "1" ]; then
SMALL="--scaffolds 10 --steps 1 --warmup 1 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1 --layout reads"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1f_pileup|k3_enum_pairs_tiles|k3_pair_stats_dev' -s 3 -c 3 \
    -f -o $out/${tag}_full python bench.py $SMALL > $out/${tag}_full.log 2>&1
echo "ncu full exit $?"
ISB_K1F=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1f_pileup' -s 1 -c 1 \
    -f -o $out/${tag}_full_unfused python bench.py $SMALL > $out/${tag}_full_unfused.log 2>&1
echo "ncu full unfused exit $?"
fi
