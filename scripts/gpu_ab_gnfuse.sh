#!/bin/bash
# Same-box A/B of GN_FUSE (GroupNorm statistics from the producing conv's epilogue): alternating bench.py runs.
mkdir -p gpurun_out
: > gpurun_out/ab_gnfuse.txt
for v in 1 0 1 0; do
  GYRE_B200_GN_FUSE=$v timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_gnfuse.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
f={k:round(v['ms'],1) for k,v in d['families'].items() if v['ms']>1}
print('GN_FUSE=$v', round(d['value'],3),'img/s', round(d['ms_per_step'],1),'ms', f, d['gpu_launches'])
" >> gpurun_out/ab_gnfuse.txt
done
cat gpurun_out/ab_gnfuse.txt
