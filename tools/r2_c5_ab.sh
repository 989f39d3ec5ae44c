#!/bin/bash
# A/B of the two-level traversal kernels on C5 (1080p @ 16 spp, timing only), run under gpurun.
P=gpurun_out
mkdir -p $P
for cfg in "2 1" "2 4" "2 0"; do
  set -- $cfg
  echo "== B200PT_2L_KERNEL=$1 B200PT_2L_TUNE=$2" | tee -a $P/r2_c5_ab.txt
  B200PT_2L_KERNEL=$1 B200PT_2L_TUNE=$2 python tools/run_config.py c5 --spp 16 --li 0 --reps 3 2>&1 | grep -E "^render" | tee -a $P/r2_c5_ab.txt
done
