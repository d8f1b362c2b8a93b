#!/bin/bash
# k_range code size (unrolling of the 16-record groups) vs its time alone and inside the pipeline
mkdir -p gpurun_out
L=gpurun_out/r3a.log
: > $L
for U in 1 2 8; do
  echo "#### unroll $U" >> $L
  B200_EXTRA_NVCC="-DB200_RANGE_UNROLL=$U" python __graft_entry__.py -f > /dev/null 2>&1
  (timeout 600 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -1) >> $L
  PROBE_KERNELS=1 python tools/probe_content.py 128 grain 2>&1 | grep "^B=\|kernel" | tail -2 >> $L
  B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A6 "^band" | tail -3 >> $L
done
cat $L
