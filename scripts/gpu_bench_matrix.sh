#!/bin/bash
# Round-2 measurement matrix on ONE B200: BASELINE configs 2-5 through bench.py, the strong-scaling per-GPU batches of
# config 2 with and without the whole-loop CUDA graph, and the torch-library comparator.  Lines land in gpurun_out/matrix.jsonl.
mkdir -p gpurun_out
: > gpurun_out/matrix.jsonl
run() { echo "== bench.py $*"; timeout 900 python bench.py --no-cpu-baseline "$@" 2>> gpurun_out/matrix.err | tee -a gpurun_out/matrix.jsonl | cut -c1-420; }
for b in 1 2 4; do
  run --config c2 --batch $b --steps 3 --warmup 3 --no-profile
  run --config c2 --batch $b --steps 3 --warmup 3 --no-profile --cuda-graph
done
run --config c2 --steps 3 --warmup 3 --cuda-graph --no-profile
run --config c3 --steps 2 --warmup 2
run --config c3 --steps 2 --warmup 2 --hires-fix --no-profile          # config 3 with the reference's default option
run --config c4 --steps 2 --warmup 2
run --config c5 --steps 2 --warmup 2
run --config c2 --impl torchlib --steps 2 --warmup 2
tail -n 5 gpurun_out/matrix.err
