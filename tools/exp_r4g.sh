#!/bin/bash
# fps against frames in flight (device-resident, 4K16 grain)
mkdir -p gpurun_out
L=gpurun_out/r4g.log
: > $L
for B in 16 32 64 96; do
  (PROBE_KERNELS=1 timeout 300 python tools/probe_content.py $B grain 2>&1 | grep "^B=\|kernel" | tail -2) >> $L
done
cat $L
