#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, the ncu launch list of the same command at reduced size and
# one `ncu --set full` capture of the dominant kernels.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1h'
tag=${1:-run}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1

if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -3 $out/${tag}_pytest.log
fi

if [ "${SKIP_BENCH:-0}" != "1" ]; then
  timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
  echo "bench exit $?"
  tail -c 1500 $out/${tag}_bench.json
fi

if [ "${SKIP_NCU:-0}" != "1" ]; then
  SMALL="--scaffolds 10 --steps 1 --warmup 1 --also-events 0 --no-cpu-baseline --e2e-scaffolds 1"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
      --log-file $out/${tag}_launches.csv python bench.py $SMALL > $out/${tag}_launches.log 2>&1
  echo "ncu launches exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'k1c_pileup|k1r_pileup|k2_call|k3r_site_rows|k3_enum_pairs|k3_pair_stats' -s 8 -c 8 \
      -f -o $out/${tag}_full python bench.py $SMALL > $out/${tag}_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la $out | tail -20
