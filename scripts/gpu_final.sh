#!/bin/bash
# Round-end measurement session on ONE B200: all GPU tests, smoke, the default bench line (with the CPU baseline leg), the
# reference arm, the configuration matrix, the per-shape kernel table and the ncu passes.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
bash scripts/gpu_session.sh tests
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench -> $?"; cut -c1-600 gpurun_out/bench_final.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm -> $?"; cut -c1-400 gpurun_out/bench_reference.json
bash scripts/gpu_bench_matrix.sh
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernel_shapes.log 2>&1; echo "kernel shapes -> $?"; tail -4 gpurun_out/kernel_shapes.log
bash scripts/gpu_profile.sh 2>&1 | tail -12
