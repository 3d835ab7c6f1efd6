#!/bin/bash
# ncu evidence for one UNet forward (CFG batch 16) and one VAE decode (batch 8); everything lands in gpurun_out/:
#   launches_unet.csv / launches_vae.csv   every launch with device time + DRAM bytes (cold-cache, serialised:
#                                          compare SHARES with the live numbers, not absolutes)
#   prof_*.ncu-rep                         --set full captures of the top kernels
mkdir -p gpurun_out
# ncu cannot replay the cooperative cluster launches of the stream-K convs (launch list stops with exit 9 at the first
# one): the profiling passes run with stream-K off (whole-tile scheduling of the 8x8-level convs; everything else unchanged)
export GYRE_B200_STREAMK=0
NCU="ncu --clock-control none --profile-from-start off"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
$NCU --metrics $M --csv --log-file gpurun_out/launches_unet.csv python scripts/profile_step.py --no-vae > gpurun_out/prof1.log 2>&1
echo "launch list (unet) exit $?"; wc -l gpurun_out/launches_unet.csv
$NCU --metrics $M --csv --log-file gpurun_out/launches_vae.csv python scripts/profile_step.py --no-unet > gpurun_out/prof1b.log 2>&1
echo "launch list (vae) exit $?"; wc -l gpurun_out/launches_vae.csv
# conv: skip conv_in, take level-0 / level-1 convs
$NCU --set full --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.[0-9]+, .bool.1" -s 1 -c 6 -o gpurun_out/prof_conv -f python scripts/profile_step.py --no-vae > gpurun_out/prof2.log 2>&1
echo "conv exit $?"
$NCU --set full --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.[0-9]+, .bool.0" -s 3 -c 8 -o gpurun_out/prof_gemm -f python scripts/profile_step.py --no-vae > gpurun_out/prof3.log 2>&1
echo "gemm exit $?"
$NCU --set full -k regex:attention -c 4 -o gpurun_out/prof_attention -f python scripts/profile_step.py --no-vae > gpurun_out/prof4.log 2>&1
echo "attention exit $?"
$NCU --set full -k regex:"gn_|layernorm" -c 6 -o gpurun_out/prof_norm -f python scripts/profile_step.py --no-vae > gpurun_out/prof5.log 2>&1
echo "norm exit $?"
# the reports are tens of MB each and gpurun_out/ travels back with a 64 MiB cap: keep the raw-metric pages only
for n in conv gemm attention norm; do
  ncu -i gpurun_out/prof_$n.ncu-rep --page raw --csv > gpurun_out/prof_${n}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$n.ncu-rep
done
ls -la gpurun_out/prof_*_raw.csv
