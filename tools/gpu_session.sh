#!/bin/bash
# usage: tools/gpu_session.sh <logname> group1[:args] group2 ...   (each group in its own process with a timeout)
mkdir -p gpurun_out
LOG=gpurun_out/$1.log; shift
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv >> $LOG 2>&1
for g in "$@"; do
  echo "===== $g =====" >> $LOG
  IFS=: read -ra parts <<< "$g"
  timeout 600 python tests/diag/gpu_diag.py "${parts[@]}" >> $LOG 2>&1
  echo "exit=$?" >> $LOG
done
tail -c 6000 $LOG
