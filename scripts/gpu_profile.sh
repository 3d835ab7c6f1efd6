#!/bin/bash
# ncu evidence for one UNet forward (CFG batch 16) + one VAE decode (batch 8):
#   launches.csv            every launch with its device time (cold-cache, serialised: compare SHARES)
#   prof_*.ncu-rep          --set full captures of the top kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof1.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
# conv / gemm: skip the first launches (conv_in, time-embedding GEMMs), take a level-0 conv and the GEMMs that follow
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 4 -c 14 -o gpurun_out/prof_gemm -f python scripts/profile_step.py --no-vae > gpurun_out/prof3.log 2>&1
echo "gemm exit $?"
$NCU --set full --import-source on -k regex:attention -c 2 -o gpurun_out/prof_attention -f python scripts/profile_step.py --no-vae > gpurun_out/prof2.log 2>&1
echo "attention exit $?"
$NCU --set full --import-source on -k regex:"gn_|layernorm" -c 6 -o gpurun_out/prof_norm -f python scripts/profile_step.py --no-vae > gpurun_out/prof4.log 2>&1
echo "norm exit $?"
ls -la gpurun_out/*.ncu-rep
