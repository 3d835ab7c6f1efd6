#!/bin/bash
# Multi-GPU measurement matrix on ONE node with N GPUs (launched like the driver launches bench.py): strong scaling of the
# BASELINE configurations (global batch split over the ranks) next to the weak-scaling default.  usage: gpu_scale_matrix.sh N
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/scale_$N.jsonl
: > $OUT
PORT=29510
run() {
  echo "== N=$N bench.py $*"
  PORT=$((PORT + 1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --no-cpu-baseline --no-profile "$@" 2>> gpurun_out/scale_$N.err | tee -a $OUT | cut -c1-330
}
run --config c2 --steps 3 --warmup 3                                   # weak: 8 images per GPU
run --config c2 --scaling strong --steps 3 --warmup 3                  # global batch 8
run --config c2 --scaling strong --steps 3 --warmup 3 --cuda-graph
run --config c4 --scaling strong --steps 2 --warmup 2                  # global batch 16 (BASELINE configs[3]: sharded 2 / 4 / 8 ways)
if [ "$N" -ge 8 ]; then
  run --config c5 --scaling strong --steps 2 --warmup 2                # global batch 64 over 8 GPUs (BASELINE configs[4])
fi
tail -n 3 gpurun_out/scale_$N.err
