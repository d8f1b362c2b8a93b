#!/bin/bash
# k_range: depth of the record ring (blocks asked for ahead) vs its time per band inside the pipeline
mkdir -p gpurun_out
L=gpurun_out/r2x.log
: > $L
for SL in 4 8; do
  echo "#### ring slots $SL" >> $L
  B200_EXTRA_NVCC="-DB200_RANGE_SLOTS=$SL" python __graft_entry__.py -f > /dev/null 2>&1
  (timeout 600 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -1) >> $L
  PROBE_KERNELS=1 python tools/probe_content.py 128 grain 2>&1 | grep "^B=\|kernel" | tail -2 >> $L
  B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A6 "^band" | head -8 >> $L
done
cat $L
