#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2g.log
: > $L
(timeout 900 python -m pytest tests/test_ffv1_gpu.py -m gpu -x -q 2>&1 | tail -4) >> $L
run() { echo "== B=${B:-128} $*" >> $L; env "$@" python tools/probe_content.py ${B:-128} grain 2>&1 | grep "^B=" | tail -1 >> $L; }
run B200_RANGE_SMS=24
B=64 run B200_RANGE_SMS=24
B=64 run B200_RANGE_SMS=16
B=96 run B200_RANGE_SMS=24
echo "== kernels serial" >> $L
PROBE_KERNELS=1 python tools/probe_content.py 128 grain 2>&1 | grep kernel >> $L
echo "== trace R=24" >> $L
B200_TRACE=1 python tools/probe_content.py 128 grain 2>&1 | grep -A8 "^band" | head -10 >> $L
cat $L
