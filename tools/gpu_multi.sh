#!/bin/bash
# Multi-GPU bench line(s) under torchrun on however many GPUs the box shows: strong scaling (+ the weak line beside).
tag=${1:-multi}
out=gpurun_out
mkdir -p $out
export PYTHONUNBUFFERED=1
NG=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $out/${tag}_gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $NG ${BENCH_ARGS} > $out/${tag}_n$NG.json 2> $out/${tag}_n$NG.err
echo "bench N=$NG exit $?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("$out/${tag}_n$NG.json") if l.startswith("{")][-1])
    print("N=%d value %.3e ms/step %.3f scaling %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["scaling"]), d["roofline"]["stage_ms_per_step"])
    print("weak", d["weak_scaling"]); print("e2e", d["e2e"] and ("%.3e" % d["e2e"]["value"], d["e2e"]["single_context"], d["e2e"]["two_contexts_pipelined"]))
    print("gather bytes/step", d["gather_bytes_per_step"], "sustained", d["sustained"])
except Exception as ex:
    print("failed", ex)
PY
tail -5 $out/${tag}_n$NG.err
