#!/bin/bash
# A/B of the K1r variants on the bench workload (20 scaffolds keep it short): bash tools/gpu_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
ARGS="--scaffolds 20 --steps 5 --warmup 3 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
ISB_K1R_VARIANT=1 timeout 600 python bench.py $ARGS > $out/${tag}_v1.json 2> $out/${tag}_v1.err
timeout 600 python bench.py $ARGS > $out/${tag}_v2.json 2> $out/${tag}_v2.err
python - <<PY
import json
for v in ("v1", "v2"):
    try:
        d = json.load(open("$out/${tag}_%s.json" % v))
        print(v, "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], d["roofline"]["stage_ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as ex:
        print(v, "failed", ex)
PY
