#!/bin/bash
# decoder: registers per thread (CTAs per SM) x frames in flight x slices per warp
mkdir -p gpurun_out
L=gpurun_out/dec_occ.log
: > $L
for M in 4 6 8; do
  echo "#### min CTAs per SM $M" >> $L
  B200_EXTRA_NVCC="-DB200_DEC_MIN_CTAS=$M" python __graft_entry__.py -f -v 2>&1 | grep -A3 "k_decodeILb1" | grep "Used" >> $L
  (timeout 600 python tools/probe_decode.py 128 grain 1 2 4 2>&1 | grep "^decode") >> $L
  (timeout 600 python tools/probe_decode.py 256 grain 2 4 2>&1 | grep "^decode") >> $L
done
cat $L
