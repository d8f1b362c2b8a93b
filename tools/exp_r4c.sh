#!/bin/bash
# frames in flight: 128 / 192 / 256 on the 4K16 workload (device-resident), kernel times beside
mkdir -p gpurun_out
L=gpurun_out/r4c.log
: > $L
for B in 128 192 256; do
  (PROBE_KERNELS=1 timeout 600 python tools/probe_content.py $B grain 2>&1 | grep "^B=\|kernel" | tail -2) >> $L
done
for RS in 24 32; do
  echo "## B200_RANGE_SMS=$RS B=256" >> $L
  (B200_RANGE_SMS=$RS PROBE_KERNELS=1 timeout 600 python tools/probe_content.py 256 grain 2>&1 | grep "^B=\|kernel" | tail -2) >> $L
done
cat $L
