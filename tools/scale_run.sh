#!/bin/bash
# Scaling record of the BASELINE configs on ONE box (run under `gpurun --gpus 8`): c1 (data parallel, config 2 at N=8),
# c3 (hierarchical sampling, sharded table), c4 (posterior extraction) at N = 1, 2, 4, 8 -> gpurun_out/scale/*.json
OUT=gpurun_out/scale; mkdir -p $OUT
run() {  # n config extra...
  n=$1; c=$2; shift 2
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --config $c --steps 40 --warmup 5 --no-cpu-baseline "$@" 2>$OUT/${c}_$n$TAG.err | tail -1 > $OUT/${c}_$n$TAG.json
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $n --config $c --steps 40 --warmup 5 --no-cpu-baseline "$@" 2>$OUT/${c}_$n$TAG.err | tail -1 > $OUT/${c}_$n$TAG.json; fi
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${c}_$n$TAG.json")); print("$c N=$n $TAG", round(d["value"]), "seg/s", round(d["ms_per_step"],4), "ms/step", "e2e", round(d.get("e2e",{}).get("value",0)))
except Exception as e: print("$c N=$n $TAG FAILED", e)
PY
}
# optional argument: a list of "config:N[:tag:extra flags]" items instead of the full set, e.g. "c1:8 c1:1 c3:8"
if [ -n "$1" ]; then
  for item in $1; do
    IFS=: read c n tag extra <<< "$item"
    TAG="${tag:+_$tag}" run $n $c $extra
  done
  exit 0
fi
for n in 8 4 2 1; do TAG="" run $n c1; done
TAG="_overlap" run 8 c1 --overlap
for n in 8 4 2 1; do TAG="" run $n c3; done
for n in 8 4 2 1; do TAG="" run $n c4; done
TAG="_bf16" run 1 c1 --mode bf16
TAG="_f32" run 1 c1 --mode f32
