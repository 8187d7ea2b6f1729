#!/bin/bash
# Round 2, first GPU call: parity of the new read-major kernels (K1f), then A/B bench lines and a launch list.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_reads.py tests/test_gpu_zy_params.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
ARGS="--scaffolds 20 --steps 5 --warmup 3 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
i=0
for spec in "--layout reads" "ISB_K1F=0 --layout reads" "ISB_K1F=0 ISB_K1R_LEGACY=1 --layout reads" "--layout cols"; do
  envs=""; extra=""
  for tok in $spec; do case $tok in *=*) envs="$envs $tok";; *) extra="$extra $tok";; esac; done
  env $envs timeout 600 python bench.py $ARGS $extra > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_$i.json"))
    print("[$spec]", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "e2e %.3e" % d["e2e"]["value"], d["rows"])
except Exception as ex:
    print("[$spec] failed", ex)
PY
  tail -3 $out/${tag}_$i.err
  i=$((i+1))
done
SMALL="--scaffolds 10 --steps 1 --warmup 1 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1 --layout reads"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file $out/${tag}_launches.csv python bench.py $SMALL > $out/${tag}_launches.log 2>&1
echo "ncu launches exit $?"
