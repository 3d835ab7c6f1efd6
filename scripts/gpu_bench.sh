#!/bin/bash
# bench + launch list + one full ncu capture of the top kernels; everything lands in gpurun_out/
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
