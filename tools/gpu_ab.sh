#!/bin/bash
# Short bench runs (20 scaffolds) for A/B comparisons of kernel changes:
#   bash tools/gpu_ab.sh <tag> ["ENV=.. ENV2=.. -- extra bench args" ...]   (each variant: optional env assignments, then --, then args)
tag=${1:-ab}; shift
out=gpurun_out
mkdir -p $out
ARGS="--scaffolds 20 --steps 5 --warmup 3 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
i=0
for spec in " -- " "$@"; do
  envs="${spec%%--*}"; extra="${spec#*--}"
  env $envs timeout 600 python bench.py $ARGS $extra > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_$i.json"))
    print("[$spec]", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "e2e %.3e" % d["e2e"]["value"])
except Exception as ex:
    print("[$spec] failed", ex)
PY
  i=$((i+1))
done
