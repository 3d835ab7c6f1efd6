#!/bin/bash
# Trimmed round-end session on ONE B200 (the GPU budget left did not cover scripts/gpu_final.sh): all GPU tests, smoke, the
# default bench line (with the CPU baseline leg), the reference arm, configs c3-c5, the per-shape kernel table, the ncu passes.
mkdir -p gpurun_out
bash scripts/gpu_session.sh tests
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench -> $?"; cut -c1-600 gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm -> $?"; cut -c1-400 gpurun_out/bench_reference.json
: > gpurun_out/matrix_short.jsonl
for c in c3 c4 c5; do
  timeout 900 python bench.py --no-cpu-baseline --config $c --steps 2 --warmup 2 2>> gpurun_out/matrix.err | tee -a gpurun_out/matrix_short.jsonl | cut -c1-300
done
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernel_shapes.log 2>&1; echo "kernel shapes -> $?"; tail -4 gpurun_out/kernel_shapes.log
bash scripts/gpu_profile.sh 2>&1 | tail -12
