#!/bin/bash
# Round 2: full GPU test suite + the default bench line (new bench.py) + smoke; optional 2-GPU line when visible.
tag=${1:-r2d}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
  timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?"; tail -4 $out/${tag}_smoke.log
fi
timeout 900 python bench.py ${BENCH_ARGS} > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench exit $?"; tail -c 3000 $out/${tag}_bench.json; tail -5 $out/${tag}_bench.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG ${BENCH_ARGS} > $out/${tag}_bench_n$NG.json 2> $out/${tag}_bench_n$NG.err
  echo "bench N=$NG exit $?"; tail -c 2500 $out/${tag}_bench_n$NG.json; tail -5 $out/${tag}_bench_n$NG.err
fi
