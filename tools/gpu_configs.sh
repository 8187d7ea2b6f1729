#!/bin/bash
# The BASELINE configurations besides the default: C2 (1 Mb x 50x, pileup + SNV only), the reference's default mode
# (mm profiling on, M ~ 15) and C5 (LD stress: 10 Mb x 500x, 5 % SNVs); one bench line each + ncu launch lists.
tag=${1:-cfg}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
COMMON="--also-layouts 0 --from-bam-scaffolds 0 --sustain-s 1"
run() { name=$1; shift
  timeout 900 python bench.py $COMMON "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_$name.json"))
    print("[$name]", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()},
          "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % (d["e2e"]["value"] if d["e2e"] else 0), d["rows"])
except Exception as ex:
    print("[$name] failed", ex)
PY
  tail -2 $out/${tag}_$name.err
}
run C2 --scaffolds 1 --L 1000000 --cov 50 --skip-linkage --e2e-scaffolds 1 --steps 50
run mm --mm --scaffolds ${MM_SC:-40} --e2e-scaffolds 2 --steps 10
run C5 --scaffolds 1 --L 10000000 --cov 500 --dens 0.05 --no-e2e --no-cpu-baseline --steps 5
if [ "${NCU:-1}" = "1" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_mm.csv \
      python bench.py $COMMON --mm --scaffolds 5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --sustain-s 0 > $out/${tag}_launches_mm.log 2>&1
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_C5.csv \
      python bench.py $COMMON --scaffolds 1 --L 2000000 --cov 500 --dens 0.05 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --sustain-s 0 > $out/${tag}_launches_C5.log 2>&1
  echo "ncu exit $?"
fi
