#!/bin/bash
# ncu evidence for one UNet forward + VAE decode: launch list (durations) + full captures of the top kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py > gpurun_out/prof1.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
# full sets: attention (level 0 self-attn is the 1st attention launch), conv/gemm kernel, groupnorm
$NCU --set full --import-source on -k regex:attention_kernel -c 2 -o gpurun_out/prof_attention -f python scripts/profile_step.py --no-vae > gpurun_out/prof2.log 2>&1
echo "attention exit $?"
$NCU --set full --import-source on -k regex:gemm_tc_kernel -s 6 -c 6 -o gpurun_out/prof_gemm -f python scripts/profile_step.py --no-vae > gpurun_out/prof3.log 2>&1
echo "gemm exit $?"
$NCU --set full --import-source on -k regex:gn_ -c 3 -o gpurun_out/prof_gn -f python scripts/profile_step.py --no-vae > gpurun_out/prof4.log 2>&1
echo "gn exit $?"
ls -la gpurun_out/
