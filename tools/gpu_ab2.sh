#!/bin/bash
# A/B bench lines (20 scaffolds) for semicolon-separated specs "ENV=.. --args"; optional fused-parity tests first and an
# ncu --set full capture of the K1f kernels afterwards.
#   gpurun -- 'SPECS="--layout reads;ISB_LIB_PATH=... --layout reads" TESTS=1 NCU=1 bash tools/gpu_ab2.sh tag'
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
if [ "${TESTS:-0}" = "1" ]; then
  timeout 1200 python -m pytest tests/test_gpu_reads.py -m gpu -x -q -k "fused or synthetic_parity or pileup_reads_stage or many_mm or layout_violations" > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi
ARGS="--scaffolds ${NSC:-20} --steps 5 --warmup 3 --also-layouts 0 --no-cpu-baseline --e2e-scaffolds 1 --sustain-s 0"
IFS=';' read -ra SP <<< "${SPECS:---layout reads}"
i=0
for spec in "${SP[@]}"; do
  envs=""; extra=""
  for tok in $spec; do case $tok in *=*) envs="$envs $tok";; *) extra="$extra $tok";; esac; done
  env $envs timeout 600 python bench.py $ARGS $extra > $out/${tag}_$i.json 2> $out/${tag}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_$i.json"))
    print("[$spec]", "value %.3e" % d["value"], "ms/step %.3f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["roofline"]["stage_ms_per_step"].items()}, "frac %.3f" % d["roofline"]["frac"], "e2e %.3e" % (d["e2e"]["value"] if d["e2e"] else 0))
except Exception as ex:
    print("[$spec] failed", ex)
PY
  tail -2 $out/${tag}_$i.err
  i=$((i+1))
done
if [ "${NCU:-0}" = "1" ]; then
  SMALL="--scaffolds 10 --steps 1 --warmup 1 --also-layouts 0 --no-cpu-baseline --e2e-scaffolds 1 --sustain-s 0 ${NCU_ARGS}"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k1f_pileup|k3_enum_pairs_tiles|k3_pair_stats_dev}" -s ${NCU_S:-3} -c ${NCU_C:-3} \
      -f -o $out/${tag}_full python bench.py $SMALL > $out/${tag}_full.log 2>&1
  echo "ncu full exit $?"
fi
